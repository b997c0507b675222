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
// leaf: m x w (w <= 16) panel.  Every thread owns RPT matrix rows in registers; the rows are spread over the CTAs of one
// thread-block cluster.  One column step costs ONE __syncthreads and ONE mbarrier wait:
//   warp-shuffle argmax -> each warp publishes (|pivot|, row, the candidate row's 16 values) in smem -> __syncthreads ->
//   warp 0 picks the CTA's candidate and pushes it into EVERY CTA of the cluster with st.async (DSMEM stores that
//   complete a transaction count on the destination's mbarrier: data and signal travel together) -> everybody waits on
//   its own mbarrier, picks the same global winner from its local copy of the candidates and updates its rows.
// The old contents of row j (needed by the thread whose row is swapped out) ride along in the same push.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lsmem(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async(uint32_t raddr, double v, uint32_t rbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr), "l"(__double_as_longlong(v)),
                 "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

template <int NT, int RPT>
__global__ void __launch_bounds__(NT, 1)
lu_leaf_kernel(double *a, long long lda, long long sa, int m, int w, int row0, int col0, int32_t *ipiv, long long sp, int32_t *info) {
    constexpr int W = LEAF, NWARP = NT / 32, ROWS = NT * RPT;
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int item = blockIdx.y;
    a += item * sa;
    ipiv += item * sp;
    info += item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rbase = crank * ROWS + tid;  // rows rbase + k*NT, k < RPT

    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ double s_wval[NWARP];
    __shared__ int s_wrow[NWARP];
    __shared__ double s_wdata[NWARP][W];
    __shared__ double s_cand[2][MAXC][W + 2];  // per parity, per source CTA: |pivot|, row, the candidate row
    __shared__ double s_rowj[2][W];            // row j before the interchange (pushed by CTA 0)
    __shared__ double s_myrowj[W];

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsmem(&s_bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsmem(&s_bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    double v[RPT][W];
#pragma unroll
    for (int k = 0; k < RPT; ++k)
#pragma unroll
        for (int c = 0; c < W; ++c) {
            const int r = rbase + k * NT;
            v[k][c] = (r < m && c < w) ? a[c * lda + r] : 0.0;
        }
    // every CTA's barriers are initialised before anybody pushes
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t tx_bytes = csize * (W + 2) * 8 + W * 8;

    // The column loop is NOT unrolled (an unrolled body is ~350 KB of straight-line SASS that runs once: the kernel was
    // instruction-fetch bound).  To keep every register index static the row is ROTATED by one slot per step: slot 0 is
    // always the current column, updated trailing entries move down one slot, and the finished value of the step (the
    // multiplier l, or the U entry for rows that are already done) is appended at slot 15.  After w steps slot 16-w+c holds
    // column c, for every row.  Whole-row interchanges automatically carry the multipliers of earlier columns along.
#pragma unroll 1
    for (int j = 0; j < w; ++j) {
        const int par = j & 1;
        const uint32_t bar = lsmem(&s_bar[par]);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
        // ---- candidate among this thread's rows still in play (row >= j); lowest row wins ties (IDAMAX) ----
        double val = -1.0;
        int row = 0x7fffffff, kb = 0;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = rbase + k * NT;
            const double x = fabs(v[k][0]);
            if (r >= j && r < m && x > val) { val = x; row = r; kb = k; }
        }
        const int myrow = row;
        double bval = val;
        int brow = row;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bval, o);
            const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (ov > bval || (ov == bval && orow < brow)) { bval = ov; brow = orow; }
        }
        if (bval >= 0.0 && myrow == brow) {  // exactly one lane: publish the warp's candidate row
#pragma unroll
            for (int c = 0; c < W; ++c) {
                double x = v[0][c];
#pragma unroll
                for (int k = 1; k < RPT; ++k) x = (kb == k) ? v[k][c] : x;
                s_wdata[warp][c] = x;
            }
        }
        if (lane == 0) { s_wval[warp] = bval; s_wrow[warp] = brow; }
        if (crank == 0 && tid == j) {  // row j lives in CTA 0, thread j, k = 0 (j < 16 <= NT)
#pragma unroll
            for (int c = 0; c < W; ++c) s_myrowj[c] = v[0][c];
        }
        __syncthreads();
        if (warp == 0) {
            double x = lane < NWARP ? s_wval[lane] : -2.0;
            int r = lane < NWARP ? s_wrow[lane] : 0x7fffffff;
            int src = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, x, o);
                const int orow = __shfl_xor_sync(0xffffffffu, r, o);
                const int osrc = __shfl_xor_sync(0xffffffffu, src, o);
                if (ov > x || (ov == x && orow < r)) { x = ov; r = orow; src = osrc; }
            }
            if (lane < W + 2) {
                const double payload = lane == 0 ? x : (lane == 1 ? (double)r : (x >= 0.0 ? s_wdata[src][lane - 2] : 0.0));
                const uint32_t dst = lsmem(&s_cand[par][crank][lane]);
                for (uint32_t d = 0; d < csize; ++d) st_async(mapa(dst, d), payload, mapa(bar, d));
            }
            if (crank == 0 && lane < W) {
                const double payload = s_myrowj[lane];
                const uint32_t dst = lsmem(&s_rowj[par][lane]);
                for (uint32_t d = 0; d < csize; ++d) st_async(mapa(dst, d), payload, mapa(bar, d));
            }
        }
        bar_wait(bar, (j >> 1) & 1);
        // ---- global winner: same decision in every thread ----
        double best = -2.0;
        int p = 0x7fffffff, wr = 0;
        for (int d = 0; d < (int)csize; ++d) {
            const double cv = s_cand[par][d][0];
            const int cr = (int)s_cand[par][d][1];
            if (cv > best || (cv == best && cr < p)) { best = cv; p = cr; wr = d; }
        }
        const double *pr = &s_cand[par][wr][2];
        const double pval = pr[0];
        if (crank == 0 && tid == 0) {
            ipiv[col0 + j] = row0 + p + 1;
            if (pval == 0.0 && *info == 0) *info = col0 + j + 1;
        }
        const double rcp = 1.0 / pval;
        const bool use_rcp = fabs(pval) >= DBL_MIN;  // DGETF2: scale by the reciprocal unless the pivot is tiny
        const int last = W - 1 - j;                   // slots 1..last are trailing columns; beyond them: earlier multipliers
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = rbase + k * NT;
            if (p != j) {
                if (r == p) {  // the pivot row's old slot receives row j
#pragma unroll
                    for (int c = 0; c < W; ++c) v[k][c] = s_rowj[par][c];
                } else if (r == j) {
#pragma unroll
                    for (int c = 0; c < W; ++c) v[k][c] = pr[c];
                }
            }
            const bool active = (r > j && r < m && pval != 0.0);
            const double l = active ? (use_rcp ? v[k][0] * rcp : v[k][0] / pval) : 0.0;
            const double app = active ? l : v[k][0];
#pragma unroll
            for (int c = 1; c < W; ++c) v[k][c - 1] = (c <= last) ? v[k][c] - l * pr[c] : v[k][c];
            v[k][W - 1] = app;
        }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int r = rbase + k * NT;
        if (r < m) {
#pragma unroll
            for (int c = 0; c < W; ++c)
                if (c >= W - w) a[(c - (W - w)) * lda + r] = v[k][c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// DLASWP for the <= 16 pivots of one leaf, applied to every column outside [skip0, skip1): one thread per column.
// All loads are issued up front (16 contiguous top rows + the distinct far rows), the interchange sequence is then
// replayed on the thread's private copy in shared memory, and everything is written back: one memory round trip
// instead of 16 dependent ones.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) laswp_leaf_kernel(double *a, long long lda, long long sa, int ncols, int skip0, int skip1, int k1,
                                                          int cnt, const int32_t *ipiv, long long sp) {
    extern __shared__ double S[];  // S[slot * 256 + tid], slots 0..15 top rows, 16..31 far rows
    __shared__ int other[LEAF];    // slot swapped with top slot j
    __shared__ int farrow[LEAF];   // far row loaded into slot 16+j (or -1)
    a += blockIdx.y * sa;
    ipiv += blockIdx.y * sp;
    const int tid = threadIdx.x;
    if (tid < LEAF) {
        int o = tid, f = -1;
        if (tid < cnt) {
            const int p = ipiv[k1 + tid] - 1;
            if (p < k1 + cnt) o = p - k1;
            else {
                int first = tid;
                for (int i = 0; i < tid; ++i)
                    if (ipiv[k1 + i] - 1 == p) { first = i; break; }
                o = LEAF + first;
                if (first == tid) f = p;
            }
        }
        other[tid] = o;
        farrow[tid] = f;
    }
    __syncthreads();
    int col = blockIdx.x * 256 + tid;
    if (col >= skip0) col += skip1 - skip0;
    if (col >= ncols) return;
    double *c = a + col * lda;
    for (int j = 0; j < cnt; ++j) S[j * 256 + tid] = c[k1 + j];
    for (int j = 0; j < cnt; ++j)
        if (farrow[j] >= 0) S[(LEAF + j) * 256 + tid] = c[farrow[j]];
    for (int j = 0; j < cnt; ++j) {
        const int o = other[j];
        if (o != j) {
            const double t = S[j * 256 + tid];
            S[j * 256 + tid] = S[o * 256 + tid];
            S[o * 256 + tid] = t;
        }
    }
    for (int j = 0; j < cnt; ++j) c[k1 + j] = S[j * 256 + tid];
    for (int j = 0; j < cnt; ++j)
        if (farrow[j] >= 0) c[farrow[j]] = S[(LEAF + j) * 256 + tid];
}

// General DLASWP (used by getrs on the right-hand sides): for i in [k1,k2): swap rows i and ipiv[i]-1, thread per column.
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

template <int NT, int RPT>
int launch_leaf(const Mat &A, int i0, int j0, int m, int w, int32_t *ipiv, long long sp, int32_t *info, int csize) {
    static bool attr = false;
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(lu_leaf_kernel<NT, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    FM_CUDA(cudaLaunchKernelEx(&cfg, lu_leaf_kernel<NT, RPT>, A.at(i0, j0), (long long)A.lda, A.stride, m, w, i0, j0, ipiv, sp, info));
    count_launch();
    return FM_OK;
}

int leaf(const Mat &A, int i0, int j0, int m, int w, int32_t *ipiv, long long sp, int32_t *info) {
    if (m <= 256) return launch_leaf<256, 1>(A, i0, j0, m, w, ipiv, sp, info, 1);
    if (m <= 512) return launch_leaf<256, 2>(A, i0, j0, m, w, ipiv, sp, info, 1);
    int csize = 1;
    while (csize * 1024 < m) csize *= 2;
    if (csize > MAXC)
        return set_error(FM_ERR_ARG, "getrf: panels taller than %d rows are not supported yet (m=%d)", MAXC * 1024, m);
    return launch_leaf<256, 4>(A, i0, j0, m, w, ipiv, sp, info, csize);
}

// apply the pivots [k1, k1+cnt) of one leaf to all columns except [skip0, skip1)
int laswp_leaf(const Mat &A, int ncols_total, int skip0, int skip1, int k1, int cnt, const int32_t *ipiv, long long sp) {
    const int ncols = ncols_total - (skip1 - skip0);
    if (ncols <= 0 || cnt <= 0) return FM_OK;
    static bool attr = false;
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(laswp_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * LEAF * 256 * (int)sizeof(double)));
        attr = true;
    }
    dim3 grid(ceil_div(ncols, 256), A.batch);
    laswp_leaf_kernel<<<grid, 256, 2 * LEAF * 256 * sizeof(double), ctx().stream>>>(A.a, A.lda, A.stride, ncols_total, skip0, skip1, k1, cnt, ipiv, sp);
    FM_LAUNCH_CHECK();
    return FM_OK;
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

// factor columns [j0, j0+w) of A (rows j0..m), w <= NB, recursively (DGETRF2-style splitting).  Each 16-column leaf
// applies its row interchanges to ALL other columns of the matrix right away, so no interchanges are pending anywhere.
int panel(const Mat &A, int m, int n, int j0, int w, int32_t *ipiv, long long sp, int32_t *info) {
    if (w <= LEAF) {
        FM_TRY(leaf(A, j0, j0, m - j0, w, ipiv, sp, info));
        return laswp_leaf(A, n, j0, j0 + w, j0, w, ipiv, sp);
    }
    int w1 = ((w / 2 + LEAF - 1) / LEAF) * LEAF;
    const int w2 = w - w1;
    FM_TRY(panel(A, m, n, j0, w1, ipiv, sp, info));
    FM_TRY(trsm<false>(w1, w2, A.at(j0, j0), A.lda, A.stride, A.at(j0, j0 + w1), A.lda, A.stride, A.batch));
    FM_TRY(gemm_update(A, j0 + w1, j0 + w1, m - j0 - w1, w2, w1, j0 + w1, j0, j0, j0 + w1));
    return panel(A, m, n, j0 + w1, w2, ipiv, sp, info);
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
        FM_TRY(panel(A, m, n, j0, jb, ipiv, sp, info));
        const int nr = n - j0 - jb;
        if (nr > 0) {
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
