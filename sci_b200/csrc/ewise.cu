// ewise.cu -- HBM-bound elementwise kernels: the pointwise macros (.+ .- .* ./, flomat.rkt:3021-3146), the
// daxpy/dscal/dcopy based matrix sum, scale and copy (flomat.rkt:805-842, 1020-1036, 497-545), fills and eye.
//
// Layout: column-major (m, n, lda).  When every operand is contiguous (ld == m) the matrix is flattened to one
// long column so the grid covers m*n elements with 128-bit accesses; views (ld > m) run column by column.
// Each thread keeps 4 independent 16-byte loads per operand in flight (64 B/operand/thread) and the grid is a
// multiple of the SM count, which is what it takes to pull HBM3e bandwidth on 148 SMs.
// Roofline: HBM.  Algorithmic bytes/element: 24 (binary matrix-matrix), 16 (matrix-scalar, scale, copy), 8 (fill).
#include "common.cuh"

namespace fm {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ double apply_op(int op, double x, double y) {
    switch (op) {
        case 0: return x + y;
        case 1: return x - y;
        case 2: return x * y;
        case 3: return x / y;
        // the rest of the pointwise family (flomat.rkt:3019 `define-pointwise-unaries sin cos tan sqr sqrt log exp`, :3076 `expt`);
        // unary maps ignore y.  sqr and sqrt are correctly rounded; the transcendental ones are CUDA libdevice (<= 2 ulp).
        case 4: return pow(x, y);
        case 5: return sin(x);
        case 6: return cos(x);
        case 7: return tan(x);
        case 8: return x * x;
        case 9: return sqrt(x);
        case 10: return log(x);
        default: return exp(x);
    }
}

// generic 2-D driver: calls body(i, j) for i in [0, mv) (vector index along the column) and j in [0, n)
template <typename Body>
__device__ __forceinline__ void for_each_2d(long long mv, int n, Body body) {
    const long long stride = (long long)gridDim.x * blockDim.x * kUnroll;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        for (long long i0 = ((long long)blockIdx.x * blockDim.x) * kUnroll + threadIdx.x; i0 < mv; i0 += stride) {
            body(i0, j);
        }
    }
}

// OP: 0 add 1 sub 2 mul 3 div; MODE: 0 = A op B, 1 = A op sb, 2 = sa op B; V: doubles per access (1 or 2)
template <int OP, int MODE, int V>
__global__ void __launch_bounds__(kThreads) ewise_kernel(long long m, int n, const double *a, long long lda, double sa,
                                                          const double *b, long long ldb, double sb, double *c, long long ldc) {
    const long long mv = m / V;
    for_each_2d(mv, n, [&](long long i0, int j) {
        double x[kUnroll][V], y[kUnroll][V];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                if (MODE != 2) {
                    const double *p = a + j * lda + i * V;
                    if (V == 2) { double2 t = *reinterpret_cast<const double2 *>(p); x[u][0] = t.x; x[u][V - 1] = t.y; }
                    else x[u][0] = *p;
                }
                if (MODE != 1) {
                    const double *p = b + j * ldb + i * V;
                    if (V == 2) { double2 t = *reinterpret_cast<const double2 *>(p); y[u][0] = t.x; y[u][V - 1] = t.y; }
                    else y[u][0] = *p;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                double r[V];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const double xx = (MODE == 2) ? sa : x[u][v];
                    const double yy = (MODE == 1) ? sb : y[u][v];
                    r[v] = apply_op(OP, xx, yy);
                }
                double *q = c + j * ldc + i * V;
                if (V == 2) *reinterpret_cast<double2 *>(q) = make_double2(r[0], r[V - 1]);
                else *q = r[0];
            }
        }
    });
}

// B := alpha*A + B (alpha = +-1 in the reference => one rounding per element, same as daxpy)
template <int V>
__global__ void __launch_bounds__(kThreads) axpy2d_kernel(long long m, int n, double alpha, const double *a, long long lda,
                                                           double *b, long long ldb) {
    const long long mv = m / V;
    for_each_2d(mv, n, [&](long long i0, int j) {
        double x[kUnroll][V], y[kUnroll][V];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                const double *p = a + j * lda + i * V;
                const double *q = b + j * ldb + i * V;
                if (V == 2) {
                    double2 t = *reinterpret_cast<const double2 *>(p); x[u][0] = t.x; x[u][V - 1] = t.y;
                    double2 s = *reinterpret_cast<const double2 *>(q); y[u][0] = s.x; y[u][V - 1] = s.y;
                } else { x[u][0] = *p; y[u][0] = *q; }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                double *q = b + j * ldb + i * V;
                // daxpy semantics: y += alpha*x as a multiply followed by an add (no FMA contraction), so that
                // alpha = +-1 gives exactly x + y / y - x and other alphas match reference BLAS rounding.
                double r0 = __dadd_rn(y[u][0], __dmul_rn(alpha, x[u][0]));
                if (V == 2) {
                    double r1 = __dadd_rn(y[u][V - 1], __dmul_rn(alpha, x[u][V - 1]));
                    *reinterpret_cast<double2 *>(q) = make_double2(r0, r1);
                } else *q = r0;
            }
        }
    });
}

// num := E + O ; den := E - O  (expm.rkt:197-198 fused: 2 reads + 2 writes instead of 2 copies + 2n daxpys)
template <int V>
__global__ void __launch_bounds__(kThreads) addsub_kernel(long long m, int n, const double *e, long long lde, const double *o,
                                                           long long ldo, double *num, long long ldn, double *den, long long ldd) {
    const long long mv = m / V;
    for_each_2d(mv, n, [&](long long i0, int j) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                if (V == 2) {
                    double2 x = *reinterpret_cast<const double2 *>(e + j * lde + i * 2);
                    double2 y = *reinterpret_cast<const double2 *>(o + j * ldo + i * 2);
                    *reinterpret_cast<double2 *>(num + j * ldn + i * 2) = make_double2(x.x + y.x, x.y + y.y);
                    *reinterpret_cast<double2 *>(den + j * ldd + i * 2) = make_double2(x.x - y.x, x.y - y.y);
                } else {
                    double x = e[j * lde + i], y = o[j * ldo + i];
                    num[j * ldn + i] = x + y;
                    den[j * ldd + i] = x - y;
                }
            }
        }
    });
}

template <int V>
__global__ void __launch_bounds__(kThreads) fill_kernel(long long m, int n, double *a, long long lda, double x) {
    const long long mv = m / V;
    for_each_2d(mv, n, [&](long long i0, int j) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                if (V == 2) *reinterpret_cast<double2 *>(a + j * lda + i * 2) = make_double2(x, x);
                else a[j * lda + i] = x;
            }
        }
    });
}

template <int V>
__global__ void __launch_bounds__(kThreads) copy2d_kernel(long long m, int n, double *d, long long ldd, const double *s, long long lds) {
    const long long mv = m / V;
    for_each_2d(mv, n, [&](long long i0, int j) {
        double x[kUnroll][V];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                if (V == 2) { double2 t = *reinterpret_cast<const double2 *>(s + j * lds + i * 2); x[u][0] = t.x; x[u][V - 1] = t.y; }
                else x[u][0] = s[j * lds + i];
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < mv) {
                if (V == 2) *reinterpret_cast<double2 *>(d + j * ldd + i * 2) = make_double2(x[u][0], x[u][V - 1]);
                else d[j * ldd + i] = x[u][0];
            }
        }
    });
}

__global__ void eye_diag_kernel(double *a, long long lda, int m, int n, int k) {
    // ones on the k-th diagonal: (i, i+k)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = i + k;
    if (i < m && i >= 0 && j >= 0 && j < n) a[(long long)j * lda + i] = 1.0;
}

// out := ci*I + c1*A1 + c2*A2  (a2 may be null); each term rounded like the reference's separate passes would be
// when only one matrix term is present.  Used by the expm polynomial assembly (batched over blockIdx.z).
__global__ void __launch_bounds__(kThreads) lincomb_kernel(int m, int n, double ci, double c1, const double *a1, long long ld1, double c2,
                                                            const double *a2, long long ld2, double *out, long long ldo, long long s1,
                                                            long long s2, long long so) {
    const int z = blockIdx.z;
    a1 += z * s1;
    if (a2) a2 += z * s2;
    out += z * so;
    for (int j = blockIdx.y; j < n; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
            double v = __dmul_rn(c1, a1[j * ld1 + i]);
            if (a2) v = __dadd_rn(v, __dmul_rn(c2, a2[j * ld2 + i]));
            if (i == j) v = __dadd_rn(ci, v);
            out[j * ldo + i] = v;
        }
}

// X_z := A_z / scale[z]  (expm.rkt:206 `(./ A 2^s)` with a per-item divisor read from device memory)
__global__ void __launch_bounds__(kThreads) div_scale_kernel(int n, const double *a, long long lda, long long sa, double *x, long long ldx,
                                                              long long sx, const double *scale) {
    const int z = blockIdx.z;
    const double d = scale[z];
    a += z * sa;
    x += z * sx;
    for (int j = blockIdx.y; j < n; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[j * ldx + i] = a[j * lda + i] / d;
}

// The expm workspaces of a batch are one contiguous ld x (n*batch) array (item stride = ld*n, ld even), so the polynomial
// terms are assembled over the flattened array with 16-byte accesses, four in flight per operand per thread, and up to two
// results per pass over the inputs:  out_k := ci_k*I + c1_k*A1 + c2_k*A2  (k < NOUT; A2 may be null).
// Rounding per element is that of lincomb_kernel (separate multiply and add, flomat.rkt:1020-1036 + 805-842).
struct TermCoef { double ci, c1, c2; };
template <int NOUT>
__global__ void __launch_bounds__(kThreads) pade_terms_kernel(long long total2, int ld, int n, const double *__restrict__ a1,
                                                               const double *__restrict__ a2, TermCoef k0, TermCoef k1,
                                                               double *__restrict__ o0, double *__restrict__ o1) {
    const long long stride = (long long)gridDim.x * kThreads * kUnroll;
    for (long long i0 = (long long)blockIdx.x * kThreads * kUnroll + threadIdx.x; i0 < total2; i0 += stride) {
        double2 x[kUnroll], y[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < total2) {
                x[u] = reinterpret_cast<const double2 *>(a1)[i];
                y[u] = a2 ? reinterpret_cast<const double2 *>(a2)[i] : make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < total2) {
                const long long col = (2 * i) / ld;
                const int row = (int)(2 * i - col * ld), j = (int)(col % n);  // ld is even: both elements are in column j
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
                    const TermCoef c = k ? k1 : k0;
                    double v0 = __dmul_rn(c.c1, x[u].x), v1 = __dmul_rn(c.c1, x[u].y);
                    if (a2) {
                        v0 = __dadd_rn(v0, __dmul_rn(c.c2, y[u].x));
                        v1 = __dadd_rn(v1, __dmul_rn(c.c2, y[u].y));
                    }
                    if (row == j) v0 = __dadd_rn(c.ci, v0);
                    if (row + 1 == j) v1 = __dadd_rn(c.ci, v1);
                    reinterpret_cast<double2 *>(k ? o1 : o0)[i] = make_double2(v0, v1);
                }
            }
        }
    }
}

// flattened form of div_scale_kernel for contiguous even-sized items: x[e] = a[e] / scale[e / item]
__global__ void __launch_bounds__(kThreads) div_scale_flat_kernel(long long total2, long long item2, const double *__restrict__ a,
                                                                   double *__restrict__ x, const double *__restrict__ scale) {
    const long long stride = (long long)gridDim.x * kThreads * kUnroll;
    for (long long i0 = (long long)blockIdx.x * kThreads * kUnroll + threadIdx.x; i0 < total2; i0 += stride) {
        double2 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < total2) v[u] = reinterpret_cast<const double2 *>(a)[i];
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + (long long)u * kThreads;
            if (i < total2) {
                const double d = scale[i / item2];
                reinterpret_cast<double2 *>(x)[i] = make_double2(v[u].x / d, v[u].y / d);
            }
        }
    }
}

__global__ void axpy_strided_kernel(int n, double alpha, const double *x, long long incx, double *y, long long incy) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[i * incy] = __dadd_rn(y[i * incy], __dmul_rn(alpha, x[i * incx]));
}
__global__ void copy_strided_kernel(int n, const double *x, long long incx, double *y, long long incy) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i * incy] = x[i * incx];
}
__global__ void scal_strided_kernel(int n, double alpha, double *x, long long incx) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i * incx] = alpha * x[i * incx];
}

struct Shape {
    long long m;  // rows after flattening
    int n;        // columns after flattening
    bool vec;     // 128-bit path legal
    dim3 grid;
};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Decide flattening / vector width for up to four (pointer, ld) operands.
Shape plan(int m, int n, const void *const *ptrs, const int *lds, int count) {
    Shape s;
    bool contiguous = true, vec = (m % 2 == 0);
    for (int i = 0; i < count; ++i) {
        if (!ptrs[i]) continue;
        if (lds[i] != m) contiguous = false;
        if (!aligned16(ptrs[i]) || (lds[i] % 2) != 0) vec = false;
    }
    if (contiguous) {
        s.m = (long long)m * n;
        s.n = 1;
        vec = ((s.m % 2) == 0);
        for (int i = 0; i < count; ++i)
            if (ptrs[i] && !aligned16(ptrs[i])) vec = false;
    } else {
        s.m = m;
        s.n = n;
    }
    s.vec = vec;
    const long long mv = s.m / (vec ? 2 : 1);
    const int sms = ctx().sm_count;
    long long bx = (mv + (long long)kThreads * kUnroll - 1) / ((long long)kThreads * kUnroll);
    if (bx < 1) bx = 1;
    long long by = s.n;
    const long long cap = (long long)sms * 16;  // 16 CTAs of 256 threads per SM = full occupancy, grid-stride beyond
    if (bx > cap) bx = cap;
    if (by > 65535) by = 65535;
    if (bx * by > cap && s.n > 1) { by = cap / bx; if (by < 1) by = 1; }
    s.grid = dim3((unsigned)bx, (unsigned)by, 1);
    return s;
}

template <int OP, int MODE>
void launch_ewise(const Shape &s, const double *a, int lda, double sa, const double *b, int ldb, double sb, double *c, int ldc) {
    if (s.vec)
        ewise_kernel<OP, MODE, 2><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, a, lda, sa, b, ldb, sb, c, ldc);
    else
        ewise_kernel<OP, MODE, 1><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, a, lda, sa, b, ldb, sb, c, ldc);
}

template <int OP>
void dispatch_mode(int mode, const Shape &s, const double *a, int lda, double sa, const double *b, int ldb, double sb, double *c, int ldc) {
    if (mode == 0) launch_ewise<OP, 0>(s, a, lda, sa, b, ldb, sb, c, ldc);
    else if (mode == 1) launch_ewise<OP, 1>(s, a, lda, sa, b, ldb, sb, c, ldc);
    else launch_ewise<OP, 2>(s, a, lda, sa, b, ldb, sb, c, ldc);
}

}  // namespace

int k_ewise(int op, int m, int n, const double *a, int lda, double sa, const double *b, int ldb, double sb, double *c, int ldc) {
    if (op < 0 || op > 11) return set_error(FM_ERR_ARG, "fm_ewise: op %d not in 0..11", op);
    if (op >= 5 && !a) return set_error(FM_ERR_ARG, "fm_ewise: a unary map needs a matrix operand A");
    if (m < 0 || n < 0) return set_error(FM_ERR_ARG, "fm_ewise: negative dimension");
    if (!a && !b) return set_error(FM_ERR_ARG, "fm_ewise: at least one operand must be a matrix");
    if (!c) return set_error(FM_ERR_ARG, "fm_ewise: C is NULL");
    if ((a && lda < m) || (b && ldb < m) || ldc < m) return set_error(FM_ERR_ARG, "fm_ewise: leading dimension < m");
    if (m == 0 || n == 0) return FM_OK;
    const void *ptrs[3] = {a, b, c};
    const int lds[3] = {lda, ldb, ldc};
    Shape s = plan(m, n, ptrs, lds, 3);
    const int mode = (a && b && op < 5) ? 0 : (a ? 1 : 2);
    switch (op) {
        case 0: dispatch_mode<0>(mode, s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 1: dispatch_mode<1>(mode, s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 2: dispatch_mode<2>(mode, s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 3: dispatch_mode<3>(mode, s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 4: dispatch_mode<4>(mode, s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 5: launch_ewise<5, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 6: launch_ewise<6, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 7: launch_ewise<7, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 8: launch_ewise<8, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 9: launch_ewise<9, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
        case 10: launch_ewise<10, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
        default: launch_ewise<11, 1>(s, a, lda, sa, b, ldb, sb, c, ldc); break;
    }
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_axpy2d(int m, int n, double alpha, const double *a, int lda, double *b, int ldb) {
    if (m < 0 || n < 0 || lda < m || ldb < m) return set_error(FM_ERR_ARG, "fm_axpy2d: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    const void *ptrs[2] = {a, b};
    const int lds[2] = {lda, ldb};
    Shape s = plan(m, n, ptrs, lds, 2);
    if (s.vec) axpy2d_kernel<2><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, alpha, a, lda, b, ldb);
    else axpy2d_kernel<1><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, alpha, a, lda, b, ldb);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_scal2d(int m, int n, double sc, double *a, int lda) {
    return k_ewise(2, m, n, nullptr, 1, sc, a, lda, 0.0, a, lda);  // s * a_ij, same operand order as dscal's alpha*x
}

int k_addsub(int m, int n, const double *e, int lde, const double *o, int ldo, double *num, int ldn, double *den, int ldd) {
    if (m < 0 || n < 0 || lde < m || ldo < m || ldn < m || ldd < m) return set_error(FM_ERR_ARG, "fm_addsub: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    const void *ptrs[4] = {e, o, num, den};
    const int lds[4] = {lde, ldo, ldn, ldd};
    Shape s = plan(m, n, ptrs, lds, 4);
    if (s.vec) addsub_kernel<2><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, e, lde, o, ldo, num, ldn, den, ldd);
    else addsub_kernel<1><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, e, lde, o, ldo, num, ldn, den, ldd);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_fill(double *a, int lda, int m, int n, double x) {
    if (m < 0 || n < 0 || lda < m) return set_error(FM_ERR_ARG, "fm_fill: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    const void *ptrs[1] = {a};
    const int lds[1] = {lda};
    Shape s = plan(m, n, ptrs, lds, 1);
    if (s.vec) fill_kernel<2><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, a, lda, x);
    else fill_kernel<1><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, a, lda, x);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_eye(double *a, int lda, int m, int n, int k) {
    FM_TRY(k_fill(a, lda, m, n, 0.0));
    if (m == 0 || n == 0) return FM_OK;
    eye_diag_kernel<<<ceil_div(m, 256), 256, 0, ctx().stream>>>(a, lda, m, n, k);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_copy2d(double *dst, int ldd, const double *src, int lds_, int m, int n) {
    if (m < 0 || n < 0 || ldd < m || lds_ < m) return set_error(FM_ERR_ARG, "fm_copy2d: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    const void *ptrs[2] = {dst, src};
    const int lds[2] = {ldd, lds_};
    Shape s = plan(m, n, ptrs, lds, 2);
    if (s.vec) copy2d_kernel<2><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, dst, ldd, src, lds_);
    else copy2d_kernel<1><<<s.grid, kThreads, 0, ctx().stream>>>(s.m, s.n, dst, ldd, src, lds_);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_lincomb(int m, int n, double ci, double c1, const double *a1, int ld1, double c2, const double *a2, int ld2, double *out, int ldo,
              long long s1, long long s2, long long so, int batch) {
    if (m <= 0 || n <= 0 || batch <= 0) return FM_OK;
    dim3 grid(ceil_div(m, kThreads), n > 1024 ? 1024 : n, batch);
    lincomb_kernel<<<grid, kThreads, 0, ctx().stream>>>(m, n, ci, c1, a1, ld1, c2, a2, ld2, out, ldo, s1, s2, so);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_pade_terms(int n, int ld, int batch, const double *a1, const double *a2, double ci0, double c10, double c20, double *o0, double ci1,
                 double c11, double c21, double *o1) {
    if (n <= 0 || batch <= 0) return FM_OK;
    const long long total2 = (long long)ld * n * batch / 2;
    long long blocks = (total2 + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    const long long cap = (long long)ctx().sm_count * 16;
    if (blocks > cap) blocks = cap;
    const TermCoef k0 = {ci0, c10, c20}, k1 = {ci1, c11, c21};
    if (o1) pade_terms_kernel<2><<<(int)blocks, kThreads, 0, ctx().stream>>>(total2, ld, n, a1, a2, k0, k1, o0, o1);
    else pade_terms_kernel<1><<<(int)blocks, kThreads, 0, ctx().stream>>>(total2, ld, n, a1, a2, k0, k1, o0, o1);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_div_pow2_batched(int n, const double *a, int lda, long long sa, double *x, int ldx, long long sx, const double *scale_dev, int batch) {
    if (n <= 0 || batch <= 0) return FM_OK;
    if (lda == ldx && (batch == 1 || sa == sx) && sx == (long long)ldx * n && (ldx % 2) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0) {
        const long long total2 = sx * batch / 2;
        long long blocks = (total2 + kThreads * kUnroll - 1) / (kThreads * kUnroll);
        const long long cap = (long long)ctx().sm_count * 16;
        if (blocks > cap) blocks = cap;
        div_scale_flat_kernel<<<(int)blocks, kThreads, 0, ctx().stream>>>(total2, sx / 2, a, x, scale_dev);
        FM_LAUNCH_CHECK();
        return FM_OK;
    }
    dim3 grid(ceil_div(n, kThreads), n > 1024 ? 1024 : n, batch);
    div_scale_kernel<<<grid, kThreads, 0, ctx().stream>>>(n, a, lda, sa, x, ldx, sx, scale_dev);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_axpy_strided(int n, double alpha, const double *x, long long incx, double *y, long long incy) {
    if (n <= 0) return FM_OK;
    int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
    axpy_strided_kernel<<<blocks, 256, 0, ctx().stream>>>(n, alpha, x, incx, y, incy);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
int k_copy_strided(int n, const double *x, long long incx, double *y, long long incy) {
    if (n <= 0) return FM_OK;
    int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
    copy_strided_kernel<<<blocks, 256, 0, ctx().stream>>>(n, x, incx, y, incy);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
int k_scal_strided(int n, double alpha, double *x, long long incx) {
    if (n <= 0 || incx <= 0) return FM_OK;
    int blocks = ceil_div(n, 256); if (blocks > 148 * 8) blocks = 148 * 8;
    scal_strided_kernel<<<blocks, 256, 0, ctx().stream>>>(n, alpha, x, incx);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

}  // namespace fm
