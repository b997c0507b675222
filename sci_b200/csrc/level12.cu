// level12.cu -- the callers on either side of the dense hot path that the reference runs as BLAS level 1/2 calls or as
// interpreted Racket loops, kept on the device so that no intermediate has to be downloaded:
//   flomat-transpose (flomat.rkt:1360, O(mn) Racket loop)            -> transpose_kernel
//   flomat-extract-upper / -lower / -lower/ones (:1540-1572)           -> tri_kernel
//   pivots->flomat (:1471)                                             -> pivots_src_kernel + perm_matrix_kernel
//   cblas_dgemv (:934, flomat*vector :957-978)                         -> gemv_n_partial/finish, gemv_t
//   cblas_ddot (:1916; with incY = 0 it is `unsafe-sum` :2066)         -> dot_partial + reduce_final
//   cblas_dnrm2 (:1299), cblas_idamax (:1156), cblas_dswap (:1074)     -> nrm2_partial, iamax_partial, swap_kernel
//   flomat-column-sums / -row-sums (:2094, :2101)                      -> colsums_kernel, gemv_n with x = 1
//   .sin .cos .tan .sqr .sqrt .log .exp and .expt (:3019, :3076)       -> ewise.cu (ops 4..11)
//   dlarnv_ (:2798; `rand` / `randn` :2814-2852)                       -> larnv_kernel (LAPACK's 48-bit multiplicative generator)
// All of it is HBM-bound streaming; reductions are two-pass with a fixed combination order (deterministic).
#include "common.cuh"

namespace fm {
namespace {

constexpr int RT = 256;  // threads of the reduction kernels

// ---------------------------------------------------------------------------------------------
// transpose: 32 x 32 tiles through padded shared memory, both sides coalesced (16 B/element of HBM traffic)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(int m, int n, const double *__restrict__ a, long long lda, double *__restrict__ b,
                                                         long long ldb) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (long long t = blockIdx.x; t < (long long)((m + 31) / 32) * ((n + 31) / 32); t += gridDim.x) {
        const int ti = (int)(t % ((m + 31) / 32)), tj = (int)(t / ((m + 31) / 32));
        const int i0 = ti * 32, j0 = tj * 32;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = i0 + tx, j = j0 + ty + 8 * r;
            tile[ty + 8 * r][tx] = (i < m && j < n) ? a[j * lda + i] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = j0 + tx, i = i0 + ty + 8 * r;  // B(j, i) = A(i, j)
            if (i < m && j < n) b[i * ldb + j] = tile[tx][ty + 8 * r];
        }
        __syncthreads();
    }
}

// B(i,j) = A(i,j) inside the triangle (upper: i <= j, lower: i >= j), 0 outside, optionally 1 on the diagonal
__global__ void __launch_bounds__(256) tri_kernel(int lower, int unit, int m, int n, const double *__restrict__ a, long long lda,
                                                   double *__restrict__ b, long long ldb) {
    for (int j = blockIdx.y; j < n; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
            const bool in = lower ? (i >= j) : (i <= j);
            double v = in ? a[j * lda + i] : 0.0;
            if (unit && i == j) v = 1.0;
            b[j * ldb + i] = v;
        }
}

// pivots->flomat (flomat.rkt:1471-1484): start from I, for i = k-1 .. 0 swap rows i and ipiv[i]; row r of the result is
// e_src[r] where src follows the same swaps.  k sequential dependent swaps: one thread.
__global__ void pivots_src_kernel(int k, const int32_t *ipiv, int32_t *src) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int i = 0; i < k; ++i) src[i] = i;
    for (int i = k - 1; i >= 0; --i) {
        const int p = ipiv[i] - 1;
        if (p != i && p >= 0 && p < k) { const int t = src[i]; src[i] = src[p]; src[p] = t; }
    }
}
__global__ void __launch_bounds__(256) perm_matrix_kernel(int k, const int32_t *src, double *p, long long ldp) {
    for (int j = blockIdx.y; j < k; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) p[j * ldp + i] = (src[i] == j) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

__global__ void __launch_bounds__(RT) dot_partial_kernel(long long n, const double *__restrict__ x, long long incx, const double *__restrict__ y,
                                                          long long incy, double *part) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * RT + threadIdx.x; i < n; i += (long long)gridDim.x * RT) acc += x[i * incx] * y[i * incy];
    const double r = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
}
// out[0] = sum of part[0..cnt) in a fixed order (one CTA)
__global__ void __launch_bounds__(RT) sum_final_kernel(int cnt, const double *part, double *out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < cnt; i += RT) acc += part[i];
    const double r = block_sum(acc, sh);
    if (threadIdx.x == 0) out[0] = r;
}

// dnrm2 without overflow or a division per element: Blue's three-accumulator sum of squares (the algorithm of LAPACK >= 3.10's
// DNRM2): values above 2^486 are accumulated scaled by 2^-538, values below 2^-511 scaled by 2^537, the rest plainly.
constexpr double NRM_TSML = 0x1p-511, NRM_TBIG = 0x1p486, NRM_SSML = 0x1p537, NRM_SBIG = 0x1p-538;
__global__ void __launch_bounds__(RT) nrm2_partial_kernel(long long n, const double *__restrict__ x, long long incx, double *part /*3 per CTA*/) {
    __shared__ double sh[32];
    double amed = 0.0, asml = 0.0, abig = 0.0;
    for (long long i = (long long)blockIdx.x * RT + threadIdx.x; i < n; i += (long long)gridDim.x * RT) {
        const double v = fabs(x[i * incx]);
        if (v > NRM_TBIG) { const double t = v * NRM_SBIG; abig += t * t; }
        else if (v < NRM_TSML) { const double t = v * NRM_SSML; asml += t * t; }
        else amed += v * v;
    }
    const double m = block_sum(amed, sh), sm = block_sum(asml, sh), bg = block_sum(abig, sh);
    if (threadIdx.x == 0) { part[3 * blockIdx.x] = m; part[3 * blockIdx.x + 1] = sm; part[3 * blockIdx.x + 2] = bg; }
}
__global__ void nrm2_final_kernel(int cnt, const double *part, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double amed = 0.0, asml = 0.0, abig = 0.0;
    for (int i = 0; i < cnt; ++i) { amed += part[3 * i]; asml += part[3 * i + 1]; abig += part[3 * i + 2]; }
    double scl = 1.0, sumsq;
    if (abig > 0.0) {  // combine abig and amed
        if (amed > 0.0 || amed != amed) abig += (amed * NRM_SBIG) * NRM_SBIG;
        scl = 1.0 / NRM_SBIG;
        sumsq = abig;
    } else if (asml > 0.0) {  // combine amed and asml
        if (amed > 0.0 || amed != amed) {
            const double ym = sqrt(amed), ys = sqrt(asml) / NRM_SSML;
            const double ymin = ym < ys ? ym : ys, ymax = ym < ys ? ys : ym;
            scl = 1.0;
            sumsq = ymax * ymax * (1.0 + (ymin / ymax) * (ymin / ymax));
        } else {
            scl = 1.0 / NRM_SSML;
            sumsq = asml;
        }
    } else {
        sumsq = amed;
    }
    out[0] = scl * sqrt(sumsq);
}

// idamax: first index of the largest |x| (0-based, CBLAS).  key = (bits(|x|), ~index) so that max picks the lowest index.
__global__ void __launch_bounds__(RT) iamax_partial_kernel(long long n, const double *__restrict__ x, long long incx, unsigned long long *pk,
                                                            long long *pi) {
    __shared__ unsigned long long sk[RT];
    __shared__ long long si[RT];
    unsigned long long bk = 0ull;
    long long bi = -1;
    for (long long i = (long long)blockIdx.x * RT + threadIdx.x; i < n; i += (long long)gridDim.x * RT) {
        const double v = fabs(x[i * incx]);
        const unsigned long long k = (v == v) ? (unsigned long long)__double_as_longlong(v) : 0ull;  // a NaN never wins (IDAMAX compares with >)
        if (bi < 0 || k > bk) { bk = k; bi = i; }
    }
    sk[threadIdx.x] = bk; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = RT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const unsigned long long k2 = sk[threadIdx.x + o];
            const long long i2 = si[threadIdx.x + o];
            const long long i1 = si[threadIdx.x];
            if (i2 >= 0 && (i1 < 0 || k2 > sk[threadIdx.x] || (k2 == sk[threadIdx.x] && i2 < i1))) { sk[threadIdx.x] = k2; si[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pk[blockIdx.x] = sk[0]; pi[blockIdx.x] = si[0]; }
}
__global__ void iamax_final_kernel(int cnt, const unsigned long long *pk, const long long *pi, long long *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long bk = 0ull;
    long long bi = -1;
    for (int i = 0; i < cnt; ++i)
        if (pi[i] >= 0 && (bi < 0 || pk[i] > bk || (pk[i] == bk && pi[i] < bi))) { bk = pk[i]; bi = pi[i]; }
    out[0] = bi < 0 ? 0 : bi;
}

__global__ void __launch_bounds__(256) swap_kernel(long long n, double *x, long long incx, double *y, long long incy) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double t = x[i * incx];
        x[i * incx] = y[i * incy];
        y[i * incy] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// gemv.  N: y_i = sum_j A(i,j) x_j -- thread per row, columns split over gridDim.y chunks (partials), then a finishing pass.
//        T: y_j = sum_i A(i,j) x_i -- one CTA per column, coalesced down the column.
// ---------------------------------------------------------------------------------------------
template <int V>  // V = rows per thread (2: 16-byte loads, needs an aligned A with even lda)
__global__ void __launch_bounds__(256) gemv_n_partial_kernel(int m, int n, const double *__restrict__ a, long long lda, const double *__restrict__ x,
                                                              long long incx, int chunk, double *part /* gridDim.y x m */) {
    const int j0 = blockIdx.y * chunk, j1 = min(n, j0 + chunk);
    constexpr int U = 8;  // columns in flight per thread
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * V; i < m; i += gridDim.x * blockDim.x * V) {
        double acc[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[u][v] = 0.0;
        for (int j = j0; j < j1; j += U) {
            double t[U][V], xs[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool ok = j + u < j1;
                const double *p = a + (long long)(ok ? j + u : j) * lda + i;
                if (V == 2) {
                    if (i + 1 < m) { const double2 d = *reinterpret_cast<const double2 *>(p); t[u][0] = d.x; t[u][V - 1] = d.y; }
                    else { t[u][0] = *p; t[u][V - 1] = 0.0; }
                } else t[u][0] = *p;
                xs[u] = ok ? (x ? x[(j + u) * incx] : 1.0) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < V; ++v) acc[u][v] += t[u][v] * xs[u];
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            double r = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) r += acc[u][v];
            if (i + v < m) part[(long long)blockIdx.y * m + i + v] = r;
        }
    }
}
__global__ void __launch_bounds__(256) gemv_finish_kernel(int len, int nparts, const double *part, double alpha, double beta, double *y, long long incy) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < nparts; ++p) acc += part[(long long)p * len + i];
        const double r = alpha * acc;
        y[i * incy] = (beta == 0.0) ? r : r + beta * y[i * incy];
    }
}
__global__ void __launch_bounds__(RT) gemv_t_kernel(int m, int n, const double *__restrict__ a, long long lda, const double *__restrict__ x, long long incx,
                                                     double alpha, double beta, double *y, long long incy) {
    __shared__ double sh[32];
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        const double *col = a + (long long)j * lda;
        double acc = 0.0;
        for (int i = threadIdx.x; i < m; i += RT) acc += col[i] * (x ? x[i * incx] : 1.0);
        const double r = block_sum(acc, sh);
        if (threadIdx.x == 0) {
            const double v = alpha * r;
            y[j * incy] = (beta == 0.0) ? v : v + beta * y[j * incy];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dlarnv: LAPACK's DLARUV generator is x_k = (s0 * a^k mod 2^48) / 2^48 with a = 33952834046453 (the 128-row table MM of
// dlaruv.f is a^1..a^128 in base 4096), so the k-th uniform is available directly: every thread jumps to its position by
// modular exponentiation and strides by a^T.  idist 1: U(0,1); 2: U(-1,1); 3: N(0,1) by Box-Muller on consecutive
// pairs, sqrt(-2 ln u_{2k-1}) cos(2 pi u_{2k}) exactly as DLARNV pairs them.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long LCG_A = 33952834046453ull, LCG_MASK = (1ull << 48) - 1;
__host__ __device__ inline unsigned long long lcg_pow(unsigned long long e) {
    unsigned long long r = 1, b = LCG_A;
    while (e) {
        if (e & 1) r = (r * b) & LCG_MASK;
        b = (b * b) & LCG_MASK;
        e >>= 1;
    }
    return r;
}
__global__ void __launch_bounds__(256) larnv_kernel(int idist, unsigned long long seed, long long n, double *x) {
    const long long T = (long long)gridDim.x * blockDim.x;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double R48 = 1.0 / 281474976710656.0;  // 2^-48
    if (idist != 3) {
        unsigned long long cur = (seed * lcg_pow((unsigned long long)(t + 1))) & LCG_MASK;
        const unsigned long long step = lcg_pow((unsigned long long)T);
        for (long long k = t; k < n; k += T) {
            const double u = (double)cur * R48;
            x[k] = (idist == 1) ? u : 2.0 * u - 1.0;
            cur = (cur * step) & LCG_MASK;
        }
    } else {
        const double TWOPI = 6.28318530717958647692528676655900576839;
        unsigned long long cur = (seed * lcg_pow((unsigned long long)(2 * t + 1))) & LCG_MASK;
        const unsigned long long step = lcg_pow((unsigned long long)(2 * T - 1));
        for (long long k = t; k < n; k += T) {
            const double u1 = (double)cur * R48;
            cur = (cur * LCG_A) & LCG_MASK;
            const double u2 = (double)cur * R48;
            x[k] = sqrt(-2.0 * log(u1)) * cos(TWOPI * u2);
            cur = (cur * step) & LCG_MASK;
        }
    }
}

int blocks_for(long long n, int per_cta) {
    long long b = (n + per_cta - 1) / per_cta;
    const long long cap = (long long)ctx().sm_count * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

int k_transpose(int m, int n, const double *a, int lda, double *b, int ldb) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1) || ldb < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "transpose: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    const long long tiles = (long long)ceil_div(m, 32) * ceil_div(n, 32);
    const long long cap = (long long)ctx().sm_count * 16;
    transpose_kernel<<<(int)(tiles < cap ? tiles : cap), 256, 0, ctx().stream>>>(m, n, a, lda, b, ldb);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_tri(int lower, int unit, int m, int n, const double *a, int lda, double *b, int ldb) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1) || ldb < (m > 1 ? m : 1)) return set_error(FM_ERR_ARG, "tri: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    dim3 grid(ceil_div(m, 256) < 64 ? ceil_div(m, 256) : 64, n < 4096 ? n : 4096);
    tri_kernel<<<grid, 256, 0, ctx().stream>>>(lower ? 1 : 0, unit ? 1 : 0, m, n, a, lda, b, ldb);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_pivots_to_matrix(int k, const int32_t *ipiv, double *p, int ldp) {
    if (k < 0 || ldp < (k > 1 ? k : 1)) return set_error(FM_ERR_ARG, "pivots->flomat: bad dimensions");
    if (k == 0) return FM_OK;
    int32_t *src = static_cast<int32_t *>(scratch((size_t)k * sizeof(int32_t)));
    if (!src) return FM_ERR_NOMEM;
    pivots_src_kernel<<<1, 32, 0, ctx().stream>>>(k, ipiv, src);
    FM_LAUNCH_CHECK();
    dim3 grid(ceil_div(k, 256) < 64 ? ceil_div(k, 256) : 64, k < 4096 ? k : 4096);
    perm_matrix_kernel<<<grid, 256, 0, ctx().stream>>>(k, src, p, ldp);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_dot_device(long long n, const double *x, long long incx, const double *y, long long incy, double *d_out) {
    if (n < 0) return set_error(FM_ERR_ARG, "dot: negative length");
    const int nb = blocks_for(n, RT * 8);
    double *part = static_cast<double *>(scratch((size_t)nb * sizeof(double)));
    if (!part) return FM_ERR_NOMEM;
    dot_partial_kernel<<<nb, RT, 0, ctx().stream>>>(n, x, incx, y, incy, part);
    FM_LAUNCH_CHECK();
    sum_final_kernel<<<1, RT, 0, ctx().stream>>>(nb, part, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_nrm2_device(long long n, const double *x, long long incx, double *d_out) {
    if (n < 0) return set_error(FM_ERR_ARG, "nrm2: negative length");
    const int nb = blocks_for(n, RT * 8);
    double *part = static_cast<double *>(scratch((size_t)nb * 3 * sizeof(double)));
    if (!part) return FM_ERR_NOMEM;
    nrm2_partial_kernel<<<nb, RT, 0, ctx().stream>>>(n, x, incx, part);
    FM_LAUNCH_CHECK();
    nrm2_final_kernel<<<1, 32, 0, ctx().stream>>>(nb, part, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_iamax_device(long long n, const double *x, long long incx, long long *d_out) {
    if (n < 0) return set_error(FM_ERR_ARG, "iamax: negative length");
    const int nb = blocks_for(n, RT * 8);
    char *buf = static_cast<char *>(scratch((size_t)nb * 16));
    if (!buf) return FM_ERR_NOMEM;
    unsigned long long *pk = reinterpret_cast<unsigned long long *>(buf);
    long long *pi = reinterpret_cast<long long *>(buf + (size_t)nb * 8);
    iamax_partial_kernel<<<nb, RT, 0, ctx().stream>>>(n, x, incx, pk, pi);
    FM_LAUNCH_CHECK();
    iamax_final_kernel<<<1, 32, 0, ctx().stream>>>(nb, pk, pi, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_swap(long long n, double *x, long long incx, double *y, long long incy) {
    if (n <= 0) return FM_OK;
    swap_kernel<<<blocks_for(n, 256 * 4), 256, 0, ctx().stream>>>(n, x, incx, y, incy);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

// y := alpha op(A) x + beta y;  x == nullptr means a vector of ones (row / column sums)
int k_gemv(int trans, int m, int n, double alpha, const double *a, int lda, const double *x, long long incx, double beta, double *y, long long incy) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1)) return set_error(FM_ERR_ARG, "gemv: bad dimensions");
    const int leny = trans ? n : m;
    if (leny == 0) return FM_OK;
    if (trans) {
        gemv_t_kernel<<<n < ctx().sm_count * 8 ? n : ctx().sm_count * 8, RT, 0, ctx().stream>>>(m, n, a, lda, x, incx, alpha, beta, y, incy);
        FM_LAUNCH_CHECK();
        return FM_OK;
    }
    const bool v2 = ((reinterpret_cast<uintptr_t>(a) & 15u) == 0) && (lda % 2 == 0);
    const int rb = ceil_div(m, v2 ? 512 : 256);
    int splits = (ctx().sm_count * 4 + rb - 1) / rb;
    if (splits > ceil_div(n, 16)) splits = ceil_div(n, 16);
    if (splits < 1) splits = 1;
    const int chunk = n > 0 ? ceil_div(n, splits) : 1;
    splits = n > 0 ? ceil_div(n, chunk) : 1;
    double *part = static_cast<double *>(scratch((size_t)splits * m * sizeof(double)));
    if (!part) return FM_ERR_NOMEM;
    if (v2) gemv_n_partial_kernel<2><<<dim3(rb, splits), 256, 0, ctx().stream>>>(m, n, a, lda, x, incx, chunk, part);
    else gemv_n_partial_kernel<1><<<dim3(rb, splits), 256, 0, ctx().stream>>>(m, n, a, lda, x, incx, chunk, part);
    FM_LAUNCH_CHECK();
    gemv_finish_kernel<<<ceil_div(m, 256), 256, 0, ctx().stream>>>(m, splits, part, alpha, beta, y, incy);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_larnv(int idist, int32_t *iseed_host, long long n, double *x) {
    if (idist < 1 || idist > 3) return set_error(FM_ERR_ARG, "larnv: idist must be 1, 2 or 3");
    if (n < 0) return set_error(FM_ERR_ARG, "larnv: negative length");
    for (int i = 0; i < 4; ++i)
        if (iseed_host[i] < 0 || iseed_host[i] > 4095) return set_error(FM_ERR_ARG, "larnv: iseed entries must be in 0..4095");
    const unsigned long long seed = ((unsigned long long)iseed_host[0] << 36) | ((unsigned long long)iseed_host[1] << 24) |
                                    ((unsigned long long)iseed_host[2] << 12) | (unsigned long long)iseed_host[3];
    if (n > 0) {
        larnv_kernel<<<blocks_for(n, 256 * 4), 256, 0, ctx().stream>>>(idist, seed, n, x);
        FM_LAUNCH_CHECK();
    }
    // DLARNV draws 64 values per DLARUV call; for idist 3 each value consumes two uniforms.  The state after the call is
    // s0 * a^(uniforms consumed), written back as four base-4096 digits.
    const unsigned long long used = (unsigned long long)(idist == 3 ? 2 * n : n);
    const unsigned long long ns = (seed * lcg_pow(used)) & LCG_MASK;
    iseed_host[0] = (int32_t)((ns >> 36) & 4095);
    iseed_host[1] = (int32_t)((ns >> 24) & 4095);
    iseed_host[2] = (int32_t)((ns >> 12) & 4095);
    iseed_host[3] = (int32_t)(ns & 4095);
    return FM_OK;
}

}  // namespace fm
