// lu.cu -- blocked right-looking LU with partial pivoting and the solves built on it: the device side of
// dgetrf_ (flomat.rkt:1505, `flomat-lu!` :1525), dgesv_ (:2112, `mldivide` :3289 / `flomat-solve-many` :2145),
// dgetri_ (:1754, `inv`) and `det` (:1837-1848).  Everything takes a batch dimension (equal strides) so the same
// kernels serve the batched expm.
//
// Structure (B200):
//   * outer panels of NB=128 columns; the trailing update A22 -= L21*U12 is the TMA/DMMA GEMM of dgemm.cu
//     (>= 98% of the flops at n=8192);
//   * a panel is factored recursively (128 -> 64 -> 32 -> 16 columns, like DGETRF2) so that the updates inside
//     the panel are GEMM/TRSM shaped too;
//   * the 16-column leaf is ONE kernel on ONE thread-block cluster: every thread owns one matrix row in
//     registers (16 doubles), the per-column pivot search is a warp-shuffle argmax (lowest index wins ties, like
//     IDAMAX) + one smem stage + one cluster barrier; the winning row is pushed to every CTA of the cluster
//     through distributed shared memory, so a column step costs one cluster.sync, not a kernel launch;
//   * row interchanges (DLASWP) and the small triangular solves run as their own kernels.
// ipiv is int32, 1-based, global row numbers (flomat.rkt:1457-1463).  info: first exactly-zero pivot, 1-based.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cfloat>

namespace cg = cooperative_groups;

namespace fm {
namespace {

constexpr int NB = 128;       // outer panel width
constexpr int LEAF = 16;      // columns factored by one leaf kernel
constexpr int MAXC = 16;      // largest cluster

// ---------------------------------------------------------------------------------------------
// leaf: m x w (w <= 16) panel, one row per thread, rows spread over the CTAs of one cluster
// ---------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
lu_leaf_kernel(double *a, long long lda, long long sa, int m, int w, int row0, int col0, int32_t *ipiv, long long sp, int32_t *info) {
    constexpr int W = LEAF;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int csize = (int)cluster.num_blocks();
    const int item = blockIdx.y;
    a += item * sa;
    ipiv += item * sp;
    info += item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = crank * NT + tid;

    __shared__ double s_wval[NT / 32];
    __shared__ int s_wrow[NT / 32];
    __shared__ double s_cand[2][MAXC][W + 2];  // per parity, per source CTA: |pivot|, row, the candidate row itself
    __shared__ double s_rowj[2][W];            // contents of row j before the interchange (owner CTA only)

    double v[W];
#pragma unroll
    for (int c = 0; c < W; ++c) v[c] = (row < m && c < w) ? a[c * lda + row] : 0.0;

#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (j < w) {  // uniform across the cluster
            const int par = j & 1;
            // ---- local argmax of |v[j]| over the rows still in play (row >= j) ----
            double val = (row >= j && row < m) ? fabs(v[j]) : -1.0;
            int r = row;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, val, o);
                const int orow = __shfl_xor_sync(0xffffffffu, r, o);
                if (ov > val || (ov == val && orow < r)) { val = ov; r = orow; }
            }
            if (lane == 0) { s_wval[warp] = val; s_wrow[warp] = r; }
            __syncthreads();
            val = lane < NT / 32 ? s_wval[lane] : -2.0;
            r = lane < NT / 32 ? s_wrow[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, val, o);
                const int orow = __shfl_xor_sync(0xffffffffu, r, o);
                if (ov > val || (ov == val && orow < r)) { val = ov; r = orow; }
            }
            // ---- the CTA's candidate row is pushed into every CTA of the cluster (DSMEM) ----
            if (row == r) {
                for (int d = 0; d < csize; ++d) {
                    double *dst = cluster.map_shared_rank(&s_cand[par][crank][0], d);
                    dst[0] = val;
                    dst[1] = (double)r;
#pragma unroll
                    for (int c = 0; c < W; ++c) dst[2 + c] = v[c];
                }
            }
            if (row == j) {
#pragma unroll
                for (int c = 0; c < W; ++c) s_rowj[par][c] = v[c];
            }
            cluster.sync();
            // ---- global winner (same decision in every thread) ----
            double best = -2.0;
            int p = 0x7fffffff, wr = 0;
            for (int d = 0; d < csize; ++d) {
                const double cv = s_cand[par][d][0];
                const int cr = (int)s_cand[par][d][1];
                if (cv > best || (cv == best && cr < p)) { best = cv; p = cr; wr = d; }
            }
            const double *pr = &s_cand[par][wr][2];
            const double pval = pr[j];
            if (p != j) {
                if (row == p) {  // the pivot row's old slot receives row j
                    const double *rj = cluster.map_shared_rank(&s_rowj[par][0], j / NT);
#pragma unroll
                    for (int c = 0; c < W; ++c) v[c] = rj[c];
                } else if (row == j) {
#pragma unroll
                    for (int c = 0; c < W; ++c) v[c] = pr[c];
                }
            }
            if (crank == 0 && tid == 0) {
                ipiv[col0 + j] = row0 + p + 1;
                if (pval == 0.0 && *info == 0) *info = col0 + j + 1;
            }
            if (row > j && row < m && pval != 0.0) {
                // DGETF2: scale by the reciprocal when |pivot| >= sfmin, else divide
                const double l = (fabs(pval) >= DBL_MIN) ? v[j] * (1.0 / pval) : v[j] / pval;
                v[j] = l;
#pragma unroll
                for (int c = j + 1; c < W; ++c) v[c] -= l * pr[c];
            }
        }
    }
    if (row < m) {
#pragma unroll
        for (int c = 0; c < W; ++c)
            if (c < w) a[c * lda + row] = v[c];
    }
    cluster.sync();  // nobody exits while a peer may still read its shared memory
}

// ---------------------------------------------------------------------------------------------
// DLASWP: for i in [k1,k2): swap rows i and ipiv[i]-1 in `ncols` columns starting at `a` (row 0 = matrix row 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) laswp_kernel(double *a, long long lda, long long sa, int ncols, int k1, int k2, const int32_t *ipiv,
                                                     long long sp) {
    __shared__ int piv[NB];
    a += blockIdx.y * sa;
    ipiv += blockIdx.y * sp;
    for (int base = k1; base < k2; base += NB) {
        const int cnt = min(NB, k2 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) piv[i] = ipiv[base + i] - 1;
        __syncthreads();
        const int col = blockIdx.x * blockDim.x + threadIdx.x;
        if (col < ncols) {
            double *c = a + col * lda;
            for (int i = 0; i < cnt; ++i) {
                const int p = piv[i], r = base + i;
                if (p != r) {
                    const double t = c[r];
                    c[r] = c[p];
                    c[p] = t;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// small triangular solves, T is nb x nb (nb <= 128) in smem, B is nb x ncols solved in place, 32 columns per CTA
//   UPPER=false: unit lower (forward substitution)   UPPER=true: upper, non-unit (back substitution with division)
// ---------------------------------------------------------------------------------------------
constexpr int TRSM_CB = 32;
template <bool UPPER>
__global__ void __launch_bounds__(256) trsm_kernel(int nb, int ncols, const double *__restrict__ T, long long ldt, long long st, double *B,
                                                    long long ldb, long long sb) {
    extern __shared__ double sm[];
    double *Ts = sm;                       // Ts[k*nb + i] = T(i,k)
    double *Bs = sm + (size_t)nb * nb;     // Bs[col*(nb+1) + i]
    const int ldbs = nb + 1;
    T += blockIdx.y * st;
    B += blockIdx.y * sb;
    const int c0 = blockIdx.x * TRSM_CB;
    const int nc = min(TRSM_CB, ncols - c0);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nb * nb; idx += 256) {
        const int i = idx % nb, k = idx / nb;
        Ts[idx] = T[k * ldt + i];
    }
    for (int idx = tid; idx < nb * TRSM_CB; idx += 256) {
        const int i = idx % nb, c = idx / nb;
        Bs[c * ldbs + i] = (c < nc) ? B[(c0 + c) * ldb + i] : 0.0;
    }
    __syncthreads();
    const int col = tid & 31, wg = tid >> 5;  // lanes = columns, warps stride over rows
    double *bc = Bs + col * ldbs;
    if (!UPPER) {
        for (int k = 0; k < nb; ++k) {
            const double bk = bc[k];
            const double *tk = Ts + k * nb;
            for (int i = k + 1 + wg; i < nb; i += 8) bc[i] -= tk[i] * bk;
            __syncthreads();
        }
    } else {
        for (int k = nb - 1; k >= 0; --k) {
            const double *tk = Ts + k * nb;
            const double bk = bc[k] / tk[k];
            for (int i = wg; i < k; i += 8) bc[i] -= tk[i] * bk;
            __syncthreads();
            if (wg == 0) bc[k] = bk;
            // row k is never read again in this sweep, so no barrier is needed after the store
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nb * TRSM_CB; idx += 256) {
        const int i = idx % nb, c = idx / nb;
        if (c < nc) B[(c0 + c) * ldb + i] = Bs[c * ldbs + i];
    }
}

__global__ void __launch_bounds__(256) det_kernel(int n, const double *lu, long long lda, const int32_t *ipiv, double *out) {
    // pivots-sign (flomat.rkt:1487) times the LEFT-TO-RIGHT product of the diagonal (:1837-1842): the product is taken
    // sequentially by one thread so the rounding sequence is the reference's; the loads are staged by the whole CTA.
    __shared__ double d[2048];
    __shared__ int neg[256];
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += 256) cnt += (ipiv[i] - 1 != i);
    neg[threadIdx.x] = cnt;
    double prod = 1.0;
    for (int base = 0; base < n; base += 2048) {
        const int len = min(2048, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += 256) d[i] = lu[(base + i) * (lda + 1)];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < len; ++i) prod *= d[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
        for (int i = 0; i < 256; ++i) total += neg[i];
        // (* sign prod): the sign is an exact integer +-1
        *out = (total & 1) ? -prod : prod;
        if (n == 0) *out = 1.0;
    }
}

__global__ void zero_info_kernel(int32_t *info, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) info[i] = 0;
}

// ---------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------
struct Mat {
    double *a;
    int lda;
    long long stride;
    int batch;
    double *at(int i, int j) const { return a + (size_t)j * lda + i; }
};

template <int NT>
int launch_leaf(const Mat &A, int i0, int j0, int m, int w, int32_t *ipiv, long long sp, int32_t *info, int csize) {
    static bool attr = false;
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(lu_leaf_kernel<NT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, A.batch, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    FM_CUDA(cudaLaunchKernelEx(&cfg, lu_leaf_kernel<NT>, A.at(i0, j0), (long long)A.lda, A.stride, m, w, i0, j0, ipiv, sp, info));
    count_launch();
    return FM_OK;
}

int leaf(const Mat &A, int i0, int j0, int m, int w, int32_t *ipiv, long long sp, int32_t *info) {
    if (m <= 256) return launch_leaf<256>(A, i0, j0, m, w, ipiv, sp, info, 1);
    if (m <= 1024) return launch_leaf<1024>(A, i0, j0, m, w, ipiv, sp, info, 1);
    int csize = 2;
    while (csize * 1024 < m) csize *= 2;
    if (csize > MAXC)
        return set_error(FM_ERR_ARG, "getrf: panels taller than %d rows are not supported yet (m=%d)", MAXC * 1024, m);
    return launch_leaf<1024>(A, i0, j0, m, w, ipiv, sp, info, csize);
}

int laswp(const Mat &A, int j0, int ncols, int k1, int k2, const int32_t *ipiv, long long sp) {
    if (ncols <= 0 || k2 <= k1) return FM_OK;
    dim3 grid(ceil_div(ncols, 256), A.batch);
    laswp_kernel<<<grid, 256, 0, ctx().stream>>>(A.at(0, j0), A.lda, A.stride, ncols, k1, k2, ipiv, sp);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

template <bool UPPER>
int trsm(int nb, int ncols, const double *T, int ldt, long long st, double *B, int ldb, long long sb, int batch) {
    if (nb <= 0 || ncols <= 0) return FM_OK;
    static bool attr = false;
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(trsm_kernel<UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)((NB * NB + TRSM_CB * (NB + 1)) * sizeof(double))));
        attr = true;
    }
    const size_t smem = ((size_t)nb * nb + (size_t)TRSM_CB * (nb + 1)) * sizeof(double);
    dim3 grid(ceil_div(ncols, TRSM_CB), batch);
    trsm_kernel<UPPER><<<grid, 256, smem, ctx().stream>>>(nb, ncols, T, ldt, st, B, ldb, sb);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int gemm_update(const Mat &A, int ci, int cj, int m, int n, int k, int ai, int aj, int bi, int bj) {
    // A[ci:ci+m, cj:cj+n] -= A[ai:ai+m, aj:aj+k] * A[bi:bi+k, bj:bj+n]
    if (m <= 0 || n <= 0 || k <= 0) return FM_OK;
    GemmEpilogue ep;
    ep.alpha = -1.0;
    ep.beta = 1.0;
    return k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, n, k, A.at(ai, aj), A.lda, A.stride, A.at(bi, bj), A.lda, A.stride, A.at(ci, cj), A.lda,
                   A.stride, A.batch, ep);
}

// factor columns [j0, j0+w) of A (rows j0..m), w <= NB, recursively; pivots are applied inside the panel only
int panel(const Mat &A, int m, int j0, int w, int32_t *ipiv, long long sp, int32_t *info) {
    if (w <= LEAF) return leaf(A, j0, j0, m - j0, w, ipiv, sp, info);
    int w1 = ((w / 2 + LEAF - 1) / LEAF) * LEAF;
    if (w1 >= w) w1 = w - LEAF > 0 ? ((w - 1) / LEAF) * LEAF : w;
    const int w2 = w - w1;
    FM_TRY(panel(A, m, j0, w1, ipiv, sp, info));
    FM_TRY(laswp(A, j0 + w1, w2, j0, j0 + w1, ipiv, sp));
    FM_TRY(trsm<false>(w1, w2, A.at(j0, j0), A.lda, A.stride, A.at(j0, j0 + w1), A.lda, A.stride, A.batch));
    FM_TRY(gemm_update(A, j0 + w1, j0 + w1, m - j0 - w1, w2, w1, j0 + w1, j0, j0, j0 + w1));
    FM_TRY(panel(A, m, j0 + w1, w2, ipiv, sp, info));
    FM_TRY(laswp(A, j0, w1, j0 + w1, j0 + w, ipiv, sp));
    return FM_OK;
}

}  // namespace

int k_getrf(int m, int n, double *a, int lda, long long sa, int32_t *ipiv, long long sp, int32_t *info, int batch) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1) || batch < 0) return set_error(FM_ERR_ARG, "getrf: bad dimensions");
    if (batch == 0) return FM_OK;
    zero_info_kernel<<<ceil_div(batch, 256), 256, 0, ctx().stream>>>(info, batch);
    FM_LAUNCH_CHECK();
    if (m == 0 || n == 0) return FM_OK;
    Mat A{a, lda, sa, batch};
    const int mn = m < n ? m : n;
    for (int j0 = 0; j0 < mn; j0 += NB) {
        const int jb = (mn - j0) < NB ? (mn - j0) : NB;
        FM_TRY(panel(A, m, j0, jb, ipiv, sp, info));
        FM_TRY(laswp(A, 0, j0, j0, j0 + jb, ipiv, sp));
        const int nr = n - j0 - jb;
        if (nr > 0) {
            FM_TRY(laswp(A, j0 + jb, nr, j0, j0 + jb, ipiv, sp));
            FM_TRY(trsm<false>(jb, nr, A.at(j0, j0), lda, sa, A.at(j0, j0 + jb), lda, sa, batch));
            FM_TRY(gemm_update(A, j0 + jb, j0 + jb, m - j0 - jb, nr, jb, j0 + jb, j0, j0, j0 + jb));
        }
    }
    return FM_OK;
}

int k_getrs(int n, int nrhs, const double *lu, int lda, long long sa, const int32_t *ipiv, long long sp, double *b, int ldb, long long sb,
            int batch) {
    if (n < 0 || nrhs < 0 || lda < (n > 1 ? n : 1) || ldb < (n > 1 ? n : 1) || batch < 0) return set_error(FM_ERR_ARG, "getrs: bad dimensions");
    if (n == 0 || nrhs == 0 || batch == 0) return FM_OK;
    Mat B{b, ldb, sb, batch};
    FM_TRY(laswp(B, 0, nrhs, 0, n, ipiv, sp));
    GemmEpilogue ep;
    ep.alpha = -1.0;
    ep.beta = 1.0;
    // L (unit lower) forward sweep
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int kb = (n - k0) < NB ? (n - k0) : NB;
        FM_TRY(trsm<false>(kb, nrhs, lu + (size_t)k0 * lda + k0, lda, sa, B.at(k0, 0), ldb, sb, batch));
        const int rest = n - k0 - kb;
        if (rest > 0)
            FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, rest, nrhs, kb, lu + (size_t)k0 * lda + k0 + kb, lda, sa, B.at(k0, 0), ldb, sb,
                           B.at(k0 + kb, 0), ldb, sb, batch, ep));
    }
    // U (upper) backward sweep
    const int nblk = ceil_div(n, NB);
    for (int blk = nblk - 1; blk >= 0; --blk) {
        const int k0 = blk * NB;
        const int kb = (n - k0) < NB ? (n - k0) : NB;
        FM_TRY(trsm<true>(kb, nrhs, lu + (size_t)k0 * lda + k0, lda, sa, B.at(k0, 0), ldb, sb, batch));
        if (k0 > 0)
            FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, k0, nrhs, kb, lu + (size_t)k0 * lda, lda, sa, B.at(k0, 0), ldb, sb, B.at(0, 0), ldb, sb,
                           batch, ep));
    }
    return FM_OK;
}

int k_det_device(int n, const double *lu, int lda, const int32_t *ipiv, double *d_out) {
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "det: bad dimensions");
    det_kernel<<<1, 256, 0, ctx().stream>>>(n, lu, lda, ipiv, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

}  // namespace fm
