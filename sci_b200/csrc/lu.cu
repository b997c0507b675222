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
#include <cstdlib>

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

// Replays the interchanges piv[0..cnt) (absolute row numbers; pivot i swaps rows k1+i and piv[i]) on one column `c` whose
// top rows k1..k1+15 and distinct far rows are staged in the caller's shared-memory slots S[slot*stride + lane]:
// every load is issued before the first dependent use, so a group of 16 interchanges costs one memory round trip.
// `first` = first pivot index to apply (own-leaf columns skip the pivots that were applied in registers).
__device__ __forceinline__ void swap_group(double *c, int k1, int cnt, const int *other, const int *farrow, double *S, int stride, int lane,
                                           int first) {
    // stage through registers so that the (generic-pointer) shared stores cannot serialise the global loads
    double top[LEAF], far[LEAF];
#pragma unroll
    for (int j = 0; j < LEAF; ++j) top[j] = j < cnt ? c[k1 + j] : 0.0;
#pragma unroll
    for (int j = 0; j < LEAF; ++j) far[j] = (j >= first && j < cnt && farrow[j] >= 0) ? c[farrow[j]] : 0.0;
#pragma unroll
    for (int j = 0; j < LEAF; ++j) {
        S[j * stride + lane] = top[j];
        S[(LEAF + j) * stride + lane] = far[j];
    }
    for (int j = first; j < cnt; ++j) {
        const int o = other[j];
        if (o != j) {
            const double t = S[j * stride + lane];
            S[j * stride + lane] = S[o * stride + lane];
            S[o * stride + lane] = t;
        }
    }
#pragma unroll
    for (int j = 0; j < LEAF; ++j) {
        top[j] = S[j * stride + lane];
        far[j] = S[(LEAF + j) * stride + lane];
    }
#pragma unroll
    for (int j = 0; j < LEAF; ++j)
        if (j < cnt) c[k1 + j] = top[j];
#pragma unroll
    for (int j = 0; j < LEAF; ++j)
        if (j >= first && j < cnt && farrow[j] >= 0) c[farrow[j]] = far[j];
}

// For pivots piv[0..cnt) of rows k1..k1+cnt: other[j] = slot exchanged with top slot j (a top slot, or LEAF + first
// occurrence of that far row), farrow[j] = far row to load into slot LEAF+j (or -1).  Called by threads j < LEAF.
__device__ __forceinline__ void swap_plan(int j, int k1, int cnt, const int *piv, int *other, int *farrow) {
    int o = j, f = -1;
    if (j < cnt) {
        const int p = piv[j];
        if (p < k1 + cnt) o = p - k1;
        else {
            int first = j;
            for (int i = 0; i < j; ++i)
                if (piv[i] == p) { first = i; break; }
            o = LEAF + first;
            if (first == j) f = p;
        }
    }
    other[j] = o;
    farrow[j] = f;
}

// ---------------------------------------------------------------------------------------------
// Panel kernel: factors an m x pw panel (pw <= 128) in ONE launch on ONE cluster.  Leaves of 16 columns are factored
// in registers (column steps as described above); between leaves, inside the same kernel, the leaf's interchanges are
// applied to the other panel columns, the 16 x rest block row of U is obtained by a register-level triangular solve,
// and every thread applies the rank-16 update to its own rows of the remaining panel columns (right-looking).
// ---------------------------------------------------------------------------------------------
constexpr int PAY = LEAF + 3;

#ifdef FM_LU_PROF
__device__ long long g_prof[32];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0, t1) do { if (tid == 0 && crank == 0 && blockIdx.y == 0) g_prof[slot] += (t1) - (t0); } while (0)
#else
#define PROF_T(var) do {} while (0)
#define PROF_ADD(slot, t0, t1) do {} while (0)
#endif
  // payload of a candidate: |pivot|, row, 1/pivot, the 16 row values

template <int NT, int RPT>
__global__ void __launch_bounds__(NT, 1)
lu_panel_kernel(double *a, long long lda, long long sa, int m, int pw, int row0, int col0, int32_t *ipiv, long long sp, int32_t *info) {
    constexpr int W = LEAF, NWARP = NT / 32, ROWS = NT * RPT;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NOROW = 0x7fffffff;
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int item = blockIdx.y;
    a += item * sa;
    ipiv += item * sp;
    info += item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rbase = crank * ROWS + tid;  // panel-local rows rbase + k*NT, k < RPT

    extern __shared__ double dyn[];  // 32 x 128 doubles: swap staging, later the U block s_U[t*NB + c]
    __shared__ __align__(8) unsigned long long s_bar[3];
    __shared__ unsigned s_whi[2][NWARP], s_wlo[2][NWARP];
    __shared__ int s_wrow[2][NWARP];
    __shared__ double s_wdata[2][NWARP][W + 1];  // candidate row of each warp (+ reciprocal of its pivot)
    __shared__ double s_cand[3][MAXC][PAY];      // triple buffered: column q uses buffer q % 3 (see the pipelining note below)
    __shared__ double s_rowj[3][W];
    __shared__ double s_myrowj[2][W];
    __shared__ double s_L[W][W + 1];
    __shared__ int s_piv[W], s_other[W], s_far[W];

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 3; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsmem(&s_bar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t tx_bytes = csize * PAY * 8 + W * 8;

    for (int lc0 = 0; lc0 < pw; lc0 += W) {
        const int w = min(W, pw - lc0);
        PROF_T(tL0);
        // ---- 1. this thread's rows of the leaf columns ----
        double v[RPT][W];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = rbase + k * NT;
#pragma unroll
            for (int c = 0; c < W; ++c) v[k][c] = (r >= lc0 && r < m && c < w) ? a[(lc0 + c) * lda + r] : 0.0;
        }
        PROF_T(tL1);
        PROF_ADD(0, tL0, tL1);
        // ---- 2. column steps.
        //  * The row is ROTATED one slot per step so that all register indices stay static and the loop is not unrolled:
        //    slot 0 = current column, updated trailing entries move down one slot, the finished value is appended at slot 15.
        //  * The loop is SOFTWARE PIPELINED: as soon as pivot q is known every thread updates only the next column (one FMA
        //    per row), the candidates for pivot q+1 are reduced and pushed, and the remaining 14 FMAs per row of update q run
        //    while that push is in flight.  A CTA may therefore still read the candidates of pivot q while its peers already
        //    push pivot q+2: candidate buffers and mbarriers are triple buffered (q % 3).
        double l[RPT], app[RPT], pm[W];
#pragma unroll
        for (int k = 0; k < RPT; ++k) { l[k] = 0.0; app[k] = 0.0; }
#pragma unroll
        for (int c = 0; c < W; ++c) pm[c] = 0.0;

        // stage the candidates of panel column q: x0[k] = value of column q in row k of this thread; when `upd` the row to
        // publish is the row AFTER update q-1 (computed on the fly for the one publishing lane only)
        auto candidate_and_push = [&](const int q, const double (&x0)[RPT], const bool upd) {
            const int buf = q % 3, wb = q & 1;
            const uint32_t bar = lsmem(&s_bar[buf]);
            if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
            double val = -1.0, piv0 = 1.0;
            int row = NOROW, kb = 0;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = rbase + k * NT;
                const double x = fabs(x0[k]);
                if (r >= q && r < m && x > val) { val = x; row = r; kb = k; piv0 = x0[k]; }
            }
            const double myrcp = 1.0 / piv0;  // issued early: its latency hides behind the REDUX chain
            // warp argmax on the bit pattern of |x| (monotonic for non-negative doubles); lowest row wins ties (IDAMAX)
            const unsigned long long key = row != NOROW ? (unsigned long long)__double_as_longlong(val) : 0ull;
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mh = __reduce_max_sync(FULL, hi);
            const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
            const int wrow = __reduce_min_sync(FULL, (hi == mh && lo == ml) ? row : NOROW);
            auto row_value = [&](int k, int c) -> double {  // slot c of row k after the pending update (if any)
                if (!upd) return v[k][c];
                return c < W - 1 ? v[k][c + 1] - l[k] * pm[c + 1] : app[k];
            };
            if (row == wrow && row != NOROW) {  // exactly one lane: publish the warp's candidate row
#pragma unroll
                for (int k = 0; k < RPT; ++k)
                    if (kb == k) {
#pragma unroll
                        for (int c = 0; c < W; ++c) s_wdata[wb][warp][c] = row_value(k, c);
                    }
                s_wdata[wb][warp][W] = myrcp;
            }
            if (lane == 0) { s_whi[wb][warp] = mh; s_wlo[wb][warp] = ml; s_wrow[wb][warp] = wrow; }
            if (crank == 0 && tid == q) {  // row q lives in CTA 0, thread q, k = 0 (q < 128 <= NT)
#pragma unroll
                for (int c = 0; c < W; ++c) s_myrowj[wb][c] = row_value(0, c);
            }
            __syncthreads();
            // every warp finds the CTA's candidate (redundantly) and pushes it to the destination CTAs warp, warp+NWARP, ...
            const unsigned h2 = lane < NWARP ? s_whi[wb][lane] : 0u, l2 = lane < NWARP ? s_wlo[wb][lane] : 0u;
            const int r2 = lane < NWARP ? s_wrow[wb][lane] : NOROW;
            const unsigned MH = __reduce_max_sync(FULL, r2 != NOROW ? h2 : 0u);
            const unsigned ML = __reduce_max_sync(FULL, (r2 != NOROW && h2 == MH) ? l2 : 0u);
            const int MR = __reduce_min_sync(FULL, (r2 != NOROW && h2 == MH && l2 == ML) ? r2 : NOROW);
            const unsigned who = __ballot_sync(FULL, r2 == MR && r2 != NOROW);
            const int src = who ? __ffs(who) - 1 : 0;
            double payload = 0.0;
            if (lane == 0) payload = (double)MR;
            else if (lane < PAY) payload = (MR == NOROW) ? 0.0 : (lane == 1 ? s_wdata[wb][src][W] : s_wdata[wb][src][lane - 2]);
            const double rowj_payload = (crank == 0 && lane < W) ? s_myrowj[wb][lane] : 0.0;
            const uint32_t dst = lsmem(&s_cand[buf][crank][lane < PAY ? lane : 0]);
            const uint32_t dstj = lsmem(&s_rowj[buf][lane < W ? lane : 0]);
            for (uint32_t d = warp; d < csize; d += NWARP) {
                const uint32_t rbar = mapa(bar, d);
                if (lane < PAY) st_async(mapa(dst, d), payload, rbar);
                if (crank == 0 && lane < W) st_async(mapa(dstj, d), rowj_payload, rbar);
            }
        };

        {
            double x0[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) x0[k] = v[k][0];
            candidate_and_push(lc0, x0, false);
        }
#pragma unroll 1
        for (int j = 0; j < w; ++j) {
            PROF_T(tc0);
            const int jj = lc0 + j;  // panel-local index of the pivot position
            const int buf = jj % 3;
            bar_wait(lsmem(&s_bar[buf]), (jj / 3) & 1);
            PROF_T(tc3);
            PROF_ADD(3, tc0, tc3);
            // global winner: same decision in every warp (lane d looks at CTA d's candidate)
            int p, wr;
            {
                const int cr = lane < (int)csize ? (int)s_cand[buf][lane][0] : NOROW;
                const double cv = (lane < (int)csize && cr != NOROW) ? fabs(s_cand[buf][lane][2]) : 0.0;
                const bool ok = cr != NOROW;
                const unsigned long long ck = ok ? (unsigned long long)__double_as_longlong(cv) : 0ull;
                const unsigned ch = (unsigned)(ck >> 32), cl = (unsigned)ck;
                const unsigned GH = __reduce_max_sync(FULL, ch);
                const unsigned GL = __reduce_max_sync(FULL, (ok && ch == GH) ? cl : 0u);
                p = __reduce_min_sync(FULL, (ok && ch == GH && cl == GL) ? cr : NOROW);
                const unsigned who = __ballot_sync(FULL, ok && cr == p);
                wr = who ? __ffs(who) - 1 : 0;
            }
            const double *pr = &s_cand[buf][wr][2];
            const double pval = pr[0];
            const double rcp = s_cand[buf][wr][1];
            if (tid == 0) {
                s_piv[j] = p;
                if (crank == 0) {
                    ipiv[col0 + jj] = row0 + p + 1;
                    if (pval == 0.0 && *info == 0) *info = col0 + jj + 1;
                }
            }
            PROF_T(tc4);
            PROF_ADD(4, tc3, tc4);
            const bool use_rcp = fabs(pval) >= DBL_MIN;  // DGETF2: scale by the reciprocal unless the pivot is tiny
            const int last = W - 1 - j;                   // slots 1..last hold trailing columns; beyond them earlier multipliers
#pragma unroll
            for (int c = 1; c < W; ++c) pm[c] = (c <= last) ? pr[c] : 0.0;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int r = rbase + k * NT;
                if (p != jj) {
                    if (r == p) {  // the pivot row's old slot receives row jj
#pragma unroll
                        for (int c = 0; c < W; ++c) v[k][c] = s_rowj[buf][c];
                    } else if (r == jj) {
#pragma unroll
                        for (int c = 0; c < W; ++c) v[k][c] = pr[c];
                    }
                }
                const bool active = (r > jj && r < m && pval != 0.0);
                l[k] = active ? (use_rcp ? v[k][0] * rcp : v[k][0] / pval) : 0.0;
                app[k] = active ? l[k] : v[k][0];
            }
            if (j + 1 < w) {  // look ahead: next column first, its candidates go out before the rest of this update
                double x0[RPT];
#pragma unroll
                for (int k = 0; k < RPT; ++k) x0[k] = v[k][1] - l[k] * pm[1];
                candidate_and_push(jj + 1, x0, true);
            }
            PROF_T(tc5);
            PROF_ADD(1, tc4, tc5);
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
#pragma unroll
                for (int c = 1; c < W; ++c) v[k][c - 1] = v[k][c] - l[k] * pm[c];
                v[k][W - 1] = app[k];
            }
            PROF_T(tc6);
            PROF_ADD(5, tc5, tc6);
        }
        PROF_T(tL2);
        // ---- 3. write the leaf back (slot W-w+c holds column c); the multipliers stay in registers for step 9 ----
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = rbase + k * NT;
            if (r >= lc0 && r < m) {
#pragma unroll
                for (int c = 0; c < W; ++c)
                    if (c >= W - w) a[(lc0 + c - (W - w)) * lda + r] = v[k][c];
            }
        }
        PROF_T(tL3);
        PROF_ADD(6, tL2, tL3);
        if (pw <= W) break;  // single-leaf panel: nothing else to do (uniform)
        // ---- 4. CTA 0 applies the leaf's interchanges to the other columns of the panel ----
        if (crank == 0) {
            if (tid >= lc0 && tid < lc0 + W) {  // rows lc0..lc0+15 are final: keep L_ii for the triangular solve
#pragma unroll
                for (int c = 0; c < W; ++c) s_L[tid - lc0][c] = v[0][c];
            }
            __syncthreads();  // s_piv complete (written by tid 0 during the column steps)
            if (tid < W) swap_plan(tid, lc0, w, s_piv, s_other, s_far);
            __syncthreads();
            if (tid < pw && (tid < lc0 || tid >= lc0 + w)) swap_group(a + tid * lda, lc0, w, s_other, s_far, dyn, NB, tid, 0);
            __threadfence();
        }
        PROF_T(tL4);
        PROF_ADD(7, tL3, tL4);
        asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
        PROF_T(tL5);
        PROF_ADD(8, tL4, tL5);
        const int rest0 = lc0 + W, ncols = pw - rest0;
        if (ncols <= 0) break;  // last leaf (uniform)
        // ---- 6. CTA 0: U(lc0:lc0+16, rest) = L_ii^{-1} * A(lc0:lc0+16, rest), one column per thread, in registers ----
        if (crank == 0) {
            if (tid >= rest0 && tid < pw) {
                double x[W];
                double *col = a + tid * lda + lc0;
#pragma unroll
                for (int t = 0; t < W; ++t) x[t] = col[t];
#pragma unroll
                for (int t = 0; t < W; ++t)
#pragma unroll
                    for (int q = t + 1; q < W; ++q) x[q] -= s_L[q][t] * x[t];
#pragma unroll
                for (int t = 0; t < W; ++t) col[t] = x[t];
            }
            __threadfence();
        }
        PROF_T(tL6);
        PROF_ADD(9, tL5, tL6);
        asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
        PROF_T(tL7);
        PROF_ADD(10, tL6, tL7);
        // ---- 8. everybody stages the U block row, 9. and applies the rank-16 update to its own rows of the rest columns ----
        for (int idx = tid; idx < W * ncols; idx += NT) {
            const int t = idx % W, c = idx / W;
            dyn[t * NB + c] = a[(rest0 + c) * lda + lc0 + t];
        }
        __syncthreads();
        PROF_T(tL8);
        PROF_ADD(11, tL7, tL8);
        {
            constexpr int CU = 16 / RPT;  // columns per batch: RPT*CU = 16 independent loads in flight per thread
            bool rowok[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) rowok[k] = (rbase + k * NT >= rest0) && (rbase + k * NT < m);
            for (int c0 = 0; c0 < ncols; c0 += CU) {
                double x[RPT][CU];
#pragma unroll
                for (int k = 0; k < RPT; ++k)
#pragma unroll
                    for (int u = 0; u < CU; ++u)
                        x[k][u] = (rowok[k] && c0 + u < ncols) ? a[(rest0 + c0 + u) * lda + rbase + k * NT] : 0.0;
#pragma unroll
                for (int t = 0; t < W; ++t)
#pragma unroll
                    for (int u = 0; u < CU; ++u) {
                        const double ut = dyn[t * NB + c0 + u];
#pragma unroll
                        for (int k = 0; k < RPT; ++k) x[k][u] -= v[k][t] * ut;
                    }
#pragma unroll
                for (int k = 0; k < RPT; ++k)
#pragma unroll
                    for (int u = 0; u < CU; ++u)
                        if (rowok[k] && c0 + u < ncols) a[(rest0 + c0 + u) * lda + rbase + k * NT] = x[k][u];
            }
        }
        __syncthreads();  // dyn is reused by the next leaf's interchanges
        PROF_T(tL9);
        PROF_ADD(12, tL8, tL9);
    }
}

// ---------------------------------------------------------------------------------------------
// DLASWP for the cnt (<= 128) pivots of one panel, applied to every column outside [skip0, skip1): one thread per column,
// the pivots are replayed in groups of 16 (one memory round trip per group, see swap_group).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) laswp_panel_kernel(double *a, long long lda, long long sa, int ncols, int skip0, int skip1, int k1,
                                                           int cnt, const int32_t *ipiv, long long sp) {
    extern __shared__ double S[];  // S[slot * 256 + tid], slots 0..15 top rows, 16..31 far rows
    __shared__ int piv[LEAF], other[LEAF], farrow[LEAF];
    a += blockIdx.y * sa;
    ipiv += blockIdx.y * sp;
    const int tid = threadIdx.x;
    int col = blockIdx.x * 256 + tid;
    if (col >= skip0) col += skip1 - skip0;
    for (int g0 = 0; g0 < cnt; g0 += LEAF) {
        const int gc = min(LEAF, cnt - g0);
        __syncthreads();
        if (tid < LEAF) piv[tid] = tid < gc ? ipiv[k1 + g0 + tid] - 1 : 0;
        __syncthreads();
        if (tid < LEAF) swap_plan(tid, k1 + g0, gc, piv, other, farrow);
        __syncthreads();
        if (col < ncols) swap_group(a + col * lda, k1 + g0, gc, other, farrow, S, 256, tid, 0);
    }
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

constexpr int PANEL_DYN_SMEM = 2 * LEAF * NB * (int)sizeof(double);

template <int NT, int RPT>
int launch_panel(const Mat &A, int i0, int j0, int m, int pw, int32_t *ipiv, long long sp, int32_t *info, int csize) {
    static bool attr = false;
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(lu_panel_kernel<NT, RPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        FM_CUDA(cudaFuncSetAttribute(lu_panel_kernel<NT, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_DYN_SMEM));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, A.batch, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = PANEL_DYN_SMEM;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    FM_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_kernel<NT, RPT>, A.at(i0, j0), (long long)A.lda, A.stride, m, pw, i0, j0, ipiv, sp, info));
    count_launch();
    return FM_OK;
}

// factor the m x pw panel whose top-left element is A(j0, j0): one cluster launch.  One row per thread whenever the cluster
// (<= 16 CTAs) can cover the panel: the column step is latency/issue bound, so fewer rows per thread is faster.
int panel_factor(const Mat &A, int j0, int m, int pw, int32_t *ipiv, long long sp, int32_t *info) {
    auto pow2_ctas = [](int rows, int per_cta) {
        int cs = 1;
        while (cs * per_cta < rows) cs *= 2;
        return cs;
    };
    if (m <= 4096) return launch_panel<256, 1>(A, j0, j0, m, pw, ipiv, sp, info, pow2_ctas(m, 256));
    if (m <= 8192) return launch_panel<256, 2>(A, j0, j0, m, pw, ipiv, sp, info, pow2_ctas(m, 512));
    if (m <= 16384) return launch_panel<256, 4>(A, j0, j0, m, pw, ipiv, sp, info, pow2_ctas(m, 1024));
    return set_error(FM_ERR_ARG, "getrf: panels taller than 16384 rows are not supported yet (m=%d)", m);
}

// apply the pivots [k1, k1+cnt) of one panel to all columns except [skip0, skip1)
int laswp_panel(const Mat &A, int ncols_total, int skip0, int skip1, int k1, int cnt, const int32_t *ipiv, long long sp) {
    const int ncols = ncols_total - (skip1 - skip0);
    if (ncols <= 0 || cnt <= 0) return FM_OK;
    static bool attr = false;
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(laswp_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * LEAF * 256 * (int)sizeof(double)));
        attr = true;
    }
    dim3 grid(ceil_div(ncols, 256), A.batch);
    laswp_panel_kernel<<<grid, 256, 2 * LEAF * 256 * sizeof(double), ctx().stream>>>(A.a, A.lda, A.stride, ncols_total, skip0, skip1, k1, cnt, ipiv, sp);
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
        FM_TRY(panel_factor(A, j0, m - j0, jb, ipiv, sp, info));
        FM_TRY(laswp_panel(A, n, j0, j0 + jb, j0, jb, ipiv, sp));
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

#ifdef FM_LU_PROF
}  // namespace fm
extern "C" int fm_lu_prof(long long *out32) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, fm::g_prof, sizeof(long long) * 32);
    long long z[32] = {};
    cudaMemcpyToSymbol(fm::g_prof, z, sizeof(z));
    return 0;
}
namespace fm {
#endif

int k_det_device(int n, const double *lu, int lda, const int32_t *ipiv, double *d_out) {
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "det: bad dimensions");
    det_kernel<<<1, 256, 0, ctx().stream>>>(n, lu, lda, ipiv, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

}  // namespace fm
