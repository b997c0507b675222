// lu.cu -- recursive blocked LU with partial pivoting and the solves built on it: the device side of
// dgetrf_ (flomat.rkt:1505, `flomat-lu!` :1525), dgesv_ (:2112, `mldivide` :3289 / `flomat-solve-many` :2145),
// dgetri_ (:1754, `inv`) and `det` (:1837-1848).  Everything takes a batch dimension (equal strides) so the same
// kernels serve the batched expm.
//
// Structure (B200):
//   * the column range is split recursively (DGETRF2 style) down to panels of NB=128 columns, so ~95% of the flops run
//     in TMA/DMMA GEMMs (dgemm.cu) with a large inner dimension; triangular solves are recursive as well
//     (GEMM + 128x128 leaf solves);
//   * a 128-column panel is ONE kernel on ONE thread-block cluster (up to 16 CTAs): every thread owns matrix rows in
//     registers, 16 columns at a time; a column step is a warp REDUX argmax -> CTA argmax -> st.async push of every
//     CTA's candidate row into every CTA's shared memory (data and mbarrier signal travel together) -> the same global
//     decision everywhere.  Pivoting inside the panel is IMPLICIT: rows never move while the panel is factored (a
//     pivot row just drops out of play), each row tracks the position LAPACK's interchanges would have given it
//     (needed for IDAMAX's lowest-index tie rule and for ipiv), and the elimination arithmetic is exactly DGETF2's;
//   * the panel kernel emits ipiv (1-based, LAPACK) and the panel's NET row permutation as <= 256 (dst, src) pairs;
//     one gather/scatter kernel applies it to any column range (all loads in flight at once instead of 128 dependent
//     row swaps).  getrs rebuilds the same maps from ipiv.
// ipiv is int32, 1-based, global row numbers (flomat.rkt:1457-1463).  info: first exactly-zero pivot, 1-based.
#include "common.cuh"
#include <cfloat>
#include <cstdlib>
#include <algorithm>

namespace fm {
namespace {

constexpr int NB = 128;       // panel width
constexpr int LEAF = 16;      // columns held in registers at a time
constexpr int MAXC = 16;      // largest cluster
constexpr int MAPN = 2 * NB;  // (dst, src) pairs per panel: NB pivot rows + up to NB displaced top rows
constexpr int PAY = LEAF + 2; // candidate payload: packed (position, physical row) | 1/pivot | 16 row values
constexpr int NOROW = 0x7fffffff;
constexpr unsigned FULL = 0xffffffffu;
constexpr int UP = NB + 4;    // row stride of the U12 block in shared memory (k-major): 8 words mod 32, conflict-free DMMA B fragments
constexpr int LP = LEAF + 4;  // row stride of the multiplier tile: conflict-free DMMA A fragments
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// -DFM_LU_PROF: clock64 phase counters of warp 0 / CTA 0 of the panel kernel (developer build, tools/build_prof.sh)
#ifdef FM_LU_PROF
__device__ long long g_prof[32];
#define PROF_T(var) const long long var = clock64()
#define PROF_ADD(slot, t0, t1) do { if (tid == 0 && crank == 0 && blockIdx.y == 0) g_prof[slot] += (t1) - (t0); } while (0)
#else
#define PROF_T(var) do {} while (0)
#define PROF_ADD(slot, t0, t1) do {} while (0)
#endif
#ifdef FM_LU_PROF  // wavefront solve: thread 0 of the chunk-0 CTAs, critical section only (last off-diagonal tile + diagonal tile)
#define PROFW_DECL long long pw_last = 0; bool pw_on = false
#define PROFW(slot) do { if (tid == 0 && chunk == 0) { const long long now_ = clock64(); if (pw_on) atomicAdd((unsigned long long *)&g_prof[slot], (unsigned long long)(now_ - pw_last)); pw_last = now_; } } while (0)
#define PROFW_ON(c) do { pw_on = (c); } while (0)
#else
#define PROFW_DECL do {} while (0)
#define PROFW(slot) do {} while (0)
#define PROFW_ON(c) do {} while (0)
#endif

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
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 8-byte global -> shared copy; !valid writes zeros without touching global memory
__device__ __forceinline__ void cp_async8(uint32_t dst, const double *src, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// argmax over (key, then lowest position) of one candidate per lane; lanes without a candidate pass valid = false
// (and key = 0).  Returns true in exactly one lane (the winner), or in none when no lane has a candidate.  The common
// case (a unique maximum of the high word) costs one REDUX and one vote.
// (returns the vote itself -- one bit set, or none -- so that a caller that needs the winner's LANE does not vote a second time)
__device__ __forceinline__ unsigned warp_argmax_mask(unsigned long long key, int pos, bool valid = true) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned tie = __ballot_sync(FULL, valid && hi == mh);
    if (__popc(tie) <= 1) return tie;
    const unsigned ml = __reduce_max_sync(FULL, (valid && hi == mh) ? lo : 0u);
    const bool top = valid && hi == mh && lo == ml;
    const int mp = __reduce_min_sync(FULL, top ? pos : NOROW);
    return __ballot_sync(FULL, top && pos == mp);
}
__device__ __forceinline__ bool warp_argmax(unsigned long long key, int pos, bool valid = true) {
    return (warp_argmax_mask(key, pos, valid) >> (threadIdx.x & 31)) & 1u;
}
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
    return r;
}

// ---------------------------------------------------------------------------------------------
// Panel kernel: factors an m x pw panel (pw <= 128) in ONE launch on ONE cluster, rows in their ORIGINAL places.
// Physical panel row r belongs to CTA r / ROWS, thread (r % ROWS) % NT, slot (r % ROWS) / NT.
// Per 16-column leaf:
//   1. load the leaf columns of the rows still in play;
//   2. 16 column steps (see candidate_and_push / the loop below), software pipelined: the candidates of step q+1 go out
//      before the remaining 14 FMAs per row of update q are executed;
//   3. store the leaf columns;
//   4. every CTA forms U12 = L11^-1 * A[pivot rows, rest columns] redundantly (pivot rows are read through L2; L11 came
//      with the candidate broadcasts) and applies the rank-16 update to its own rows of the remaining panel columns.
// U12 is written back to the pivot rows one leaf later (after the cluster barrier that proves every CTA has read the old
// values).  ipiv gets LAPACK's interchange sequence, map[] the panel's net permutation.
// ---------------------------------------------------------------------------------------------
// NOPIV (round 2): the no-interchange variant for matrices on which partial pivoting provably never interchanges (the Pade denominator
// inside expm, see k_getrf_nopiv below).  The pivot of step q is panel row q, known in advance: its owner thread publishes the row and
// its warp pushes it to every CTA -- no thread/warp/CTA/cluster argmax, no __syncthreads in the column step of a cluster.  Every
// multiplier is checked (|l| <= 0.999, finite nonzero pivot: exactly "IDAMAX would have picked the diagonal") and *viol raised otherwise.
template <int NT, int RPT, int MINB = 1, bool NOPIV = false>
__global__ void __launch_bounds__(NT, MINB)
lu_panel_kernel(double *a, long long lda, long long sa, int m, int pw, int j0, int32_t *ipiv, long long sp, int32_t *info, int2 *map,
                long long smap, int32_t *viol) {
    constexpr int W = LEAF, NWARP = NT / 32, ROWS = NT * RPT;
    uint32_t crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int item = blockIdx.y;
    a += item * sa;
    ipiv += item * sp;
    info += item;
    map += item * smap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rbase = crank * ROWS + tid;  // panel-local rows rbase + k*NT, k < RPT

    // dynamic smem: the U12 blocks, double buffered, U(i, c) at sU[buf][i * UP + c]; then (MINB == 1) the multiplier tile
    // sLm[local row * LP + k] of the DMMA rank-16 update
    extern __shared__ double dyn[];
    // NOPIV: a ring of 32 (two leaves) -- only the owner of a pivot row sends, so nothing holds a fast CTA back within a leaf (the
    // cluster barrier of step 4 does between leaves); with an all-to-all exchange per step (pivoting) three buffers are enough
    constexpr int NBAR = NOPIV ? 32 : 3;
    __shared__ __align__(8) unsigned long long s_bar[NBAR];
    __shared__ int s_wpos[2][NWARP];                         // position of each warp's candidate (NOROW: none)
    __shared__ __align__(16) double s_wdata[2][NWARP][PAY];  // each warp's candidate: head | 1/pivot | 16 row values
    __shared__ __align__(16) double s_cand[3][MAXC][PAY];    // every CTA's candidate; step q uses buffer q % 3 (NOPIV: a flat ring of 32 rows)
    static_assert(3 * MAXC >= 32, "the NOPIV ring of 32 pivot rows lives in s_cand");
    __shared__ double s_L[W][W + 1];                         // L11 of the current leaf (row j: multipliers of pivot row j)
    __shared__ int s_pphys[NB];                              // physical panel row of every pivot

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBAR; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsmem(&s_bar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * NWARP) (&s_wpos[0][0])[tid] = NOROW;
    // look-ahead: the trailing-matrix GEMM queued behind this kernel (launched as a programmatic dependent) may start now
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    cluster_arrive();
    cluster_wait();
    const uint32_t tx_bytes = (NOPIV ? 1u : csize) * PAY * 8;  // NOPIV: one sender per step (the owner of the pivot row)
    // the lean batched configuration (MINB > 1) is only ever launched as a single CTA: the exchange code is compiled out
    const bool solo = (MINB > 1) || (csize == 1);
    // this warp pushes the CTA's candidate to CTAs warp, warp + NWARP, ...: remote addresses of s_cand[0][crank][lane] and
    // s_bar[0] there
    constexpr int NDEST = (MAXC + NWARP - 1) / NWARP;
    uint32_t rcand[NDEST], rbar[NDEST];
#pragma unroll
    for (int i = 0; i < NDEST; ++i) {
        const uint32_t d = warp + i * NWARP;
        rcand[i] = mapa(lsmem(&s_cand[0][crank][lane < PAY ? lane : 0]), d < csize ? d : 0);
        rbar[i] = mapa(lsmem(&s_bar[0]), d < csize ? d : 0);
    }

    // pos >= 0: the row is in play and LAPACK would hold it at panel row `pos` now;  pos < 0: out of play (~pos = the
    // pivot index it became, or a large value for rows beyond m)
    int pos[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) pos[k] = (rbase + k * NT < m) ? rbase + k * NT : ~0x3fffffff;

    int pend_lc0 = -1;  // leaf whose U12 still has to be written to the pivot rows
    bool nopiv_bad = false;

    for (int lc0 = 0; lc0 < pw; lc0 += W) {
        const int w = min(W, pw - lc0);
        PROF_T(tl0);
        // ---- 1. this thread's rows of the leaf columns ----
        double v[RPT][W];
        bool live[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            live[k] = pos[k] >= 0;
            const int r = rbase + k * NT;
#pragma unroll
            for (int c = 0; c < W; ++c) v[k][c] = (live[k] && c < w) ? a[(lc0 + c) * lda + r] : 0.0;
        }
        // ---- 2. column steps.
        //  * The row is ROTATED one slot per step so that all register indices stay static and the loop is not unrolled:
        //    slot 0 = current column, updated trailing entries move down one slot, the finished value is appended at slot 15.
        //  * The loop is SOFTWARE PIPELINED: as soon as pivot q is known every thread updates the next column (one FMA per
        //    row) and the warp reduction for pivot q+1 is issued; the remaining 14 FMAs per row of update q execute while the
        //    REDUX is in flight, then the winning lanes publish their (already updated) rows.  A CTA may still read the
        //    candidates of pivot q while its peers already push pivot q+2: candidate buffers and mbarriers are triple
        //    buffered (q % 3).

        PROF_T(tl1);
        PROF_ADD(8, tl0, tl1);
        // thread-level candidate among this thread's rows in play: largest |x| (x = slot 0), lowest position on ties (IDAMAX).
        // key = bits(|x|); rows out of play carry position NOROW and never win.
        auto thread_best = [&](const int qn, unsigned long long &bkey, int &bpos, int &kb, double &brcp) {
            bkey = 0ull; bpos = NOROW; kb = 0;
            double bval = 1.0;
            if (NOPIV) {  // the candidate is row qn itself: this thread has it or has nothing
#pragma unroll
                for (int k = 0; k < RPT; ++k)
                    if (rbase + k * NT == qn && pos[k] >= 0) { bpos = qn; kb = k; bval = v[k][0]; bkey = 1ull; }
                brcp = 1.0 / bval;
                return;
            }
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v[k][0]));
                const int pk = pos[k] >= 0 ? pos[k] : NOROW;
                if (pk != NOROW && (bpos == NOROW || key > bkey || (key == bkey && pk < bpos))) { bkey = key; bpos = pk; kb = k; bval = v[k][0]; }
            }
            brcp = 1.0 / bval;  // issued early: its latency hides behind the reduction
        };
        // the warp's winner publishes its row; the CTA's candidate goes to every CTA of the cluster.  A cluster of one CTA
        // (`solo`: panels of <= 256 rows, the batched expm) skips the exchange: the CTA's candidate IS the pivot row and is
        // read where the winning warp published it (solo_src = that warp).
        int solo_src = 0;
        auto publish_and_push = [&](const int q, const bool iwin, const int bpos, const int kb, const double brcp) {
            const int buf = NOPIV ? q % 32 : q % 3, wb = q & 1;
            PROF_T(tp0);
            if (tid == 0 && !solo)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lsmem(&s_bar[buf])), "r"(tx_bytes) : "memory");
            if (iwin) {  // at most one lane per warp
                double2 *wd = reinterpret_cast<double2 *>(s_wdata[wb][warp]);
#pragma unroll
                for (int k = 0; k < RPT; ++k)
                    if (kb == k) {
                        wd[0] = make_double2(__longlong_as_double(((long long)bpos << 32) | (long long)(unsigned)(rbase + k * NT)), brcp);
#pragma unroll
                        for (int c = 0; c < W; c += 2) wd[1 + c / 2] = make_double2(v[k][c], v[k][c + 1]);
                    }
                s_wpos[wb][warp] = bpos;
            }
            PROF_T(tp1);
            PROF_ADD(4, tp0, tp1);
            if (NOPIV) {
                // the owner of row q is one thread of one CTA: its warp pushes the row to every CTA (itself included); everybody else
                // goes straight to the mbarrier.  A single CTA reads the row where it was published, behind one barrier.
                if (solo) {
                    __syncthreads();
                    solo_src = ((q % ROWS) % NT) >> 5;
                    return;
                }
                if (__any_sync(FULL, iwin)) {
                    __syncwarp();
                    const double payload = s_wdata[wb][warp][lane < PAY ? lane : 0];
                    if (lane < PAY) {
                        const uint32_t la = lsmem(&s_cand[0][0][0]) + (uint32_t)((buf * PAY + lane) * 8), lb = lsmem(&s_bar[buf]);
                        for (uint32_t d = 0; d < csize; ++d) st_async(mapa(la, d), payload, mapa(lb, d));
                    }
                }
                return;
            }
            __syncthreads();
            PROF_T(tp2);
            PROF_ADD(5, tp1, tp2);
            // every warp finds the CTA's candidate (redundantly) and pushes it to its destination CTAs
            const int p2 = lane < NWARP ? s_wpos[wb][lane] : NOROW;
            const unsigned long long k2 = p2 != NOROW ? (unsigned long long)__double_as_longlong(fabs(s_wdata[wb][lane < NWARP ? lane : 0][2])) : 0ull;
            const unsigned who = warp_argmax_mask(k2, p2, p2 != NOROW);
            if (solo) {
                solo_src = who ? __ffs(who) - 1 : 0;
                return;
            }
            double payload;
            if (who) payload = s_wdata[wb][__ffs(who) - 1][lane < PAY ? lane : 0];
            else payload = __longlong_as_double((long long)NOROW << 32);
            const uint32_t off = buf * (MAXC * PAY * 8), boff = buf * 8;
            if (lane < PAY) {
#pragma unroll
                for (int i = 0; i < NDEST; ++i)
                    if (warp + i * NWARP < (int)csize) st_async(rcand[i] + off, payload, rbar[i] + boff);
            }
            PROF_T(tp3);
            PROF_ADD(6, tp2, tp3);
        };

        {
            unsigned long long bkey; int bpos, kb; double brcp;
            thread_best(lc0, bkey, bpos, kb, brcp);
            const bool iwin = NOPIV ? bpos != NOROW : warp_argmax(bpos != NOROW ? bkey : 0ull, bpos, bpos != NOROW);
            publish_and_push(lc0, iwin, bpos, kb, brcp);
        }
#pragma unroll 1
        for (int j = 0; j < w; ++j) {
            const int q = lc0 + j;  // pivot index inside the panel
            const int buf = NOPIV ? q % 32 : q % 3;
            if (lane == 0) s_wpos[(q + 1) & 1][warp] = NOROW;  // reset for the next publish (last read before the previous barrier)
            PROF_T(tc0);
            if (!solo) bar_wait(lsmem(&s_bar[buf]), NOPIV ? (q / 32) & 1 : (q / 3) & 1);
            PROF_T(tc1);
            PROF_ADD(0, tc0, tc1);
            // global winner: same decision in every warp (lane d looks at CTA d's candidate).  The winner's row is read through an
            // explicit shared-space address: a generic pointer that may point into either array costs a window-base computation
            // (S2UR SR_CgaCtaId + LEA) on the critical chain of every step.
            uint32_t prow;
            if (NOPIV && !solo) {
                prow = lsmem(&s_cand[0][0][0]) + (uint32_t)(buf * PAY * 8);  // ring slot of this step
            } else if (!solo) {
                const long long c0 = lane < (int)csize ? __double_as_longlong(s_cand[buf][lane][0]) : ((long long)NOROW << 32);
                const int cp = (int)(c0 >> 32);
                const unsigned long long ck = cp != NOROW ? (unsigned long long)__double_as_longlong(fabs(s_cand[buf][lane < (int)csize ? lane : 0][2])) : 0ull;
                const unsigned who = warp_argmax_mask(ck, cp, cp != NOROW);
                prow = lsmem(&s_cand[0][0][0]) + (uint32_t)((buf * MAXC + (who ? __ffs(who) - 1 : 0)) * PAY * 8);
            } else {
                prow = lsmem(&s_wdata[0][0][0]) + (uint32_t)(((q & 1) * NWARP + solo_src) * PAY * 8);
            }
            const double2 head = lds_v2(prow);
            const long long hd = __double_as_longlong(head.x);
            const int p = (int)(hd >> 32);           // LAPACK position of the pivot row (ipiv)
            const int pphys = (int)(unsigned)hd;     // where it physically lives
            const double rcp = head.y;
            double pm[W];
#pragma unroll
            for (int c = 0; c < W; c += 2) {
                const double2 t = lds_v2(prow + 16u + 8u * c);
                pm[c] = t.x;
                pm[c + 1] = t.y;
            }
            const double pval = pm[0];
            if (warp == 0) {
                if (lane < j) s_L[j][lane] = lds_f64(prow + 8u * (2 + W - j + lane));  // multipliers of pivot row j for the leaf's earlier steps
                if (lane == 0) {
                    s_pphys[q] = pphys;
                    if (crank == 0) {
                        ipiv[j0 + q] = j0 + p + 1;
                        if (pval == 0.0 && *info == 0) *info = j0 + q + 1;
                    }
                }
            }
            PROF_T(tc2);
            PROF_ADD(1, tc1, tc2);
            const bool use_rcp = fabs(pval) >= DBL_MIN;  // DGETF2: scale by the reciprocal unless the pivot is tiny
            const int last = W - 1 - j;                   // slots 1..last hold trailing columns; beyond them earlier multipliers
            if (NOPIV && (!(fabs(pval) > 0.0) || !(fabs(pval) < DBL_MAX))) nopiv_bad = true;  // a zero / non-finite pivot is not a dominant one
            double l[RPT], app[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const bool mine = (pos[k] == p);
                if (pos[k] == q && !mine) pos[k] = p;  // the row LAPACK swaps out of position q
                if (mine) pos[k] = ~q;
                const bool active = (pos[k] >= 0 && pval != 0.0);
                l[k] = active ? (use_rcp ? v[k][0] * rcp : v[k][0] / pval) : 0.0;
                if (NOPIV && !(fabs(l[k]) <= 0.999)) nopiv_bad = true;  // |l| <= 1 everywhere is "IDAMAX picks the diagonal"
                app[k] = active ? l[k] : v[k][0];
                v[k][0] = v[k][1] - l[k] * (last >= 1 ? pm[1] : 0.0);  // next column first
            }
            // reduction for pivot q+1 on the freshly updated next column, issued before the rest of the update
            unsigned long long bkey; int bpos, kb; double brcp;
            thread_best(q + 1, bkey, bpos, kb, brcp);
            const bool more = j + 1 < w;
            PROF_T(tc3);
            PROF_ADD(2, tc2, tc3);
            const bool iwin = !more ? false : (NOPIV ? bpos != NOROW : warp_argmax(bpos != NOROW ? bkey : 0ull, bpos, bpos != NOROW));
#pragma unroll
            for (int c = 2; c < W; ++c) {
                const double pc = (c <= last) ? pm[c] : 0.0;
#pragma unroll
                for (int k = 0; k < RPT; ++k) v[k][c - 1] = v[k][c] - l[k] * pc;
            }
#pragma unroll
            for (int k = 0; k < RPT; ++k) v[k][W - 1] = app[k];
            PROF_T(tc4);
            PROF_ADD(3, tc3, tc4);
            if (more) publish_and_push(q + 1, iwin, bpos, kb, brcp);
        }
        PROF_T(tl2);
        PROF_ADD(9, tl1, tl2);
        // ---- 3. write the leaf back (slot W-w+c holds column c); the multipliers stay in registers for step 4 ----
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = rbase + k * NT;
            if (live[k]) {
#pragma unroll
                for (int c = 0; c < W; ++c)
                    if (c >= W - w) a[(lc0 + c - (W - w)) * lda + r] = v[k][c];
            }
        }
        const int rest0 = lc0 + W, ncols = pw - rest0;
        if (ncols <= 0) break;  // last leaf (uniform)
        // ---- 4. U12 and the rank-16 update of the remaining panel columns ----
        PROF_T(tl3);
        PROF_ADD(10, tl2, tl3);
        if (lc0 > 0) cluster_wait();  // every CTA finished the previous leaf's update (and its reads of the old pivot rows)
        __syncthreads();              // s_L, s_pphys of this leaf are complete
        const int cur = (lc0 / W) & 1;
        double *sU = dyn + cur * (W * UP);
        if (pend_lc0 >= 0 && crank == 0) {  // delayed write of the previous leaf's U12 into its pivot rows
            const double *pU = dyn + (cur ^ 1) * (W * UP);
            const int pc = pw - (pend_lc0 + W);
            if (tid < pc) {
#pragma unroll
                for (int i = 0; i < W; ++i) a[(pend_lc0 + W + tid) * lda + s_pphys[pend_lc0 + i]] = pU[i * UP + tid];
            }
        }
        pend_lc0 = lc0;
        PROF_T(tl4);
        PROF_ADD(11, tl3, tl4);
        if (tid < ncols) {
            double x[W];
            const double *col = a + (rest0 + tid) * lda;
#pragma unroll
            for (int i = 0; i < W; ++i) x[i] = __ldcg(col + s_pphys[lc0 + i]);
#pragma unroll
            for (int t = 0; t < W; ++t)
#pragma unroll
                for (int i = t + 1; i < W; ++i) x[i] -= s_L[i][t] * x[t];
#pragma unroll
            for (int i = 0; i < W; ++i) sU[i * UP + tid] = x[i];
        }
        __syncthreads();
        PROF_T(tl5);
        PROF_ADD(12, tl4, tl5);
        if constexpr (MINB == 1) {
            // Rank-16 update on the FP64 tensor pipe: C(rows of this warp, 8 columns) -= L(rows, 16) U(16, 8 columns) as 4 DMMAs
            // per 8 x 8 tile.  With plain FMAs every U value is a broadcast LDS per two FMAs and the update is bound by the
            // shared-memory pipe (19k cycles per leaf against 7k of FP64 work); here the operands are distributed fragments.
            double *sLm = dyn + 2 * W * UP;
            unsigned act[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                double *lr = sLm + (k * NT + tid) * LP;
#pragma unroll
                for (int c = 0; c < W; c += 2) *reinterpret_cast<double2 *>(lr + c) = make_double2(-v[k][c], -v[k][c + 1]);
                act[k] = __ballot_sync(FULL, pos[k] >= 0);
            }
            __syncwarp();  // a warp consumes exactly the rows its own lanes own
            const int g = lane >> 2, t = lane & 3;
            double afr[RPT * 4][4];
#pragma unroll
            for (int k = 0; k < RPT; ++k)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int sk = 0; sk < 4; ++sk) afr[k * 4 + mt][sk] = sLm[(k * NT + 32 * warp + 8 * mt + g) * LP + 4 * sk + t];
            constexpr int NTL = RPT >= 4 ? 1 : (RPT == 2 ? 2 : 4);  // n-tiles per pass: all their C loads are issued before the first DMMA (L2 latency is the bound)
            for (int n0 = 0; n0 < ncols; n0 += 8 * NTL) {
                double bfr[NTL][4];
                double cf[NTL][RPT * 4][2];
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const int nb0 = n0 + 8 * nt, c0 = nb0 + 2 * t;
#pragma unroll
                    for (int sk = 0; sk < 4; ++sk) bfr[nt][sk] = (nb0 + g < ncols) ? sU[(4 * sk + t) * UP + nb0 + g] : 0.0;
                    const double *col0 = a + (rest0 + c0) * lda + crank * ROWS + 32 * warp + g;
#pragma unroll
                    for (int k = 0; k < RPT; ++k)
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            const bool ok = (act[k] >> (8 * mt + g)) & 1u;
                            const double *p = col0 + k * NT + 8 * mt;
                            cf[nt][k * 4 + mt][0] = (ok && c0 < ncols) ? p[0] : 0.0;
                            cf[nt][k * 4 + mt][1] = (ok && c0 + 1 < ncols) ? p[lda] : 0.0;
                        }
                }
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                    for (int i = 0; i < RPT * 4; ++i)
#pragma unroll
                        for (int sk = 0; sk < 4; ++sk) dmma884(cf[nt][i][0], cf[nt][i][1], afr[i][sk], bfr[nt][sk]);
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const int c0 = n0 + 8 * nt + 2 * t;
                    double *col0 = a + (rest0 + c0) * lda + crank * ROWS + 32 * warp + g;
#pragma unroll
                    for (int k = 0; k < RPT; ++k)
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            const bool ok = (act[k] >> (8 * mt + g)) & 1u;
                            double *p = col0 + k * NT + 8 * mt;
                            if (ok && c0 < ncols) p[0] = cf[nt][k * 4 + mt][0];
                            if (ok && c0 + 1 < ncols) p[lda] = cf[nt][k * 4 + mt][1];
                        }
                }
            }
            __syncwarp();
        } else
        {
            // columns per batch; two batches are in flight (the loads of batch i+1 are issued before the FMAs of batch i):
            // the update is bound by the L2 round trip of these loads, not by the FP64 pipe
            constexpr int CU = MINB > 1 ? 2 : (RPT >= 4 ? 2 : (RPT == 2 ? 4 : 12));
            bool rowok[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) rowok[k] = pos[k] >= 0;
            auto load = [&](double (&x)[RPT][CU], int c0) {
#pragma unroll
                for (int k = 0; k < RPT; ++k)
#pragma unroll
                    for (int u = 0; u < CU; ++u)
                        x[k][u] = (rowok[k] && c0 + u < ncols) ? a[(rest0 + c0 + u) * lda + rbase + k * NT] : 0.0;
            };
            auto update_store = [&](double (&x)[RPT][CU], int c0) {
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    const double *up = sU + (c0 + u < ncols ? c0 + u : 0);
#pragma unroll
                    for (int t2 = 0; t2 < W; ++t2) {
                        const double uu = up[t2 * UP];
#pragma unroll
                        for (int k = 0; k < RPT; ++k) x[k][u] -= v[k][t2] * uu;
                    }
                }
#pragma unroll
                for (int k = 0; k < RPT; ++k)
#pragma unroll
                    for (int u = 0; u < CU; ++u)
                        if (rowok[k] && c0 + u < ncols) a[(rest0 + c0 + u) * lda + rbase + k * NT] = x[k][u];
            };
            double xa[RPT][CU], xb[RPT][CU];
            load(xa, 0);
            for (int c0 = 0; c0 < ncols; c0 += 2 * CU) {
                load(xb, c0 + CU);
                update_store(xa, c0);
                load(xa, c0 + 2 * CU);
                update_store(xb, c0 + CU);
            }
        }
        PROF_T(tl6);
        PROF_ADD(13, tl5, tl6);
        cluster_arrive();
    }
    // ---- the last pending U12 and the panel's net permutation ----
    if (pend_lc0 >= 0) {
        cluster_wait();
        if (crank == 0) {
            const double *pU = dyn + ((pend_lc0 / W) & 1) * (W * UP);
            const int pc = pw - (pend_lc0 + W);
            if (tid < pc) {
#pragma unroll
                for (int i = 0; i < W; ++i) a[(pend_lc0 + W + tid) * lda + s_pphys[pend_lc0 + i]] = pU[i * UP + tid];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int r = rbase + k * NT;
        if (pos[k] < 0) {
            const int q = ~pos[k];
            if (q < pw) map[q] = make_int2(j0 + q, j0 + r);  // final row q comes from physical row r
        }
    }
    if (NOPIV) {
        if (__syncthreads_or(nopiv_bad) && tid == 0 && viol != nullptr) atomicOr(viol, 1);
    }
    if (crank == 0) {
        // a top row (physical row < 128: thread r % NT, slot r / NT of CTA 0) that never became a pivot ends where the
        // interchanges pushed it
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = tid + k * NT;
            if (r < NB) {
                if (r >= pw) map[r] = make_int2(-1, -1);
                map[NB + r] = (pos[k] >= 0 && pos[k] != r) ? make_int2(j0 + pos[k], j0 + r) : make_int2(-1, -1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Applies the net row permutations of panels [0, npanels) of `map` (in order) to the columns [c0, c0+ncols): per panel
// all <= 256 source rows of a column are loaded before any destination row is stored.  8 columns per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int PERM_PC = 8;
// `lower_cols` (the deferred pass over the L part of a finished factorization): panel p only touches the columns left of
// it, i.e. a column group starts at the first panel behind its own columns.
__global__ void __launch_bounds__(MAPN) permute_rows_kernel(double *a, long long lda, long long sa, int ncols, const int2 *map, long long smap,
                                                            int npanels, int lower_cols) {
    a += blockIdx.y * sa;
    map += blockIdx.y * smap;
    const int tid = threadIdx.x;
    const int col0 = blockIdx.x * PERM_PC;
    const int nc = min(PERM_PC, ncols - col0);
    double *base = a + (long long)col0 * lda;
    for (int p = lower_cols ? (col0 + nc - 1) / NB + 1 : 0; p < npanels; ++p) {
        const int2 e = map[p * MAPN + tid];
        const bool on = e.x >= 0 && e.x != e.y;
        double t[PERM_PC];
#pragma unroll
        for (int j = 0; j < PERM_PC; ++j) t[j] = (on && j < nc) ? base[j * lda + e.y] : 0.0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PERM_PC; ++j)
            if (on && j < nc) base[j * lda + e.x] = t[j];
        if (p + 1 < npanels) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// ipiv -> per-block net permutations (same format the panel kernel emits): one warp replays the <= 128 interchanges of
// block blockIdx.x on a sparse model (the 128 block rows + the far rows they touch).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) ipiv_to_maps_kernel(int n, const int32_t *ipiv, long long sp, int2 *map, long long smap, int block0,
                                                          int npiv /* pivots in the block; 0 = up to NB */) {
    __shared__ int piv[NB], top[NB], fpos[NB], fval[NB];
    ipiv += blockIdx.y * sp;
    map += blockIdx.y * smap + (long long)(blockIdx.x + block0) * MAPN;
    const int lane = threadIdx.x;
    const int k0 = (blockIdx.x + block0) * NB, cnt = min(npiv > 0 ? npiv : NB, n - k0);
    for (int i = lane; i < NB; i += 32) {
        piv[i] = i < cnt ? ipiv[k0 + i] - 1 : k0 + i;
        top[i] = k0 + i;
        fpos[i] = -1;
        fval[i] = -1;
    }
    __syncwarp();
    int nfar = 0;
    for (int i = 0; i < cnt; ++i) {
        const int p = piv[i];
        if (p == k0 + i || p < 0 || p >= n) continue;
        if (p >= k0 && p < k0 + NB) {
            if (lane == 0) { const int t = top[i]; top[i] = top[p - k0]; top[p - k0] = t; }
        } else {
            unsigned hit = 0;
            int e_found = -1;
            for (int e0 = 0; e0 < nfar; e0 += 32) {
                hit = __ballot_sync(FULL, e0 + lane < nfar && fpos[e0 + lane] == p);
                if (hit) { e_found = e0 + __ffs(hit) - 1; break; }
            }
            if (lane == 0) {
                if (e_found >= 0) { const int t = top[i]; top[i] = fval[e_found]; fval[e_found] = t; }
                else { fpos[nfar] = p; fval[nfar] = top[i]; top[i] = p; }
            }
            if (e_found < 0) ++nfar;
        }
        __syncwarp();
    }
    for (int i = lane; i < NB; i += 32) {
        map[i] = (k0 + i < n && top[i] != k0 + i) ? make_int2(k0 + i, top[i]) : make_int2(-1, -1);  // rows of the block beyond the last pivot can be displaced too
        map[NB + i] = (i < nfar && fval[i] != fpos[i]) ? make_int2(fpos[i], fval[i]) : make_int2(-1, -1);
    }
}

// ---------------------------------------------------------------------------------------------
// leaf triangular solves: T is nb x nb (nb <= 128), B is nb x ncols, solved in place.
//   UPPER=false: unit lower (forward substitution)   UPPER=true: upper, non-unit (back substitution)
// One CTA = 8 warps = 32 columns: lane = column, warp r owns rows 16b + 2r, 2r+1 of every 16-row block b, and keeps
// them in registers for the whole solve (16 doubles).  T is staged in shared memory as 16x16 blocks (triangle only) laid
// out so that the 2 rows a warp needs are one broadcast LDS.128 per column; the 16x16 DIAGONAL blocks are inverted
// in shared memory once per CTA (the standard GPU TRSM formulation: the only place a reciprocal/inverse replaces a
// division-based substitution), so every block step is the same 2x16 register mat-vec; the 16 solved values of a block
// step are exchanged between the 8 warps through a small padded smem tile (2 __syncthreads per block step).
// ---------------------------------------------------------------------------------------------
constexpr int TR_NT = 256;           // threads per CTA
constexpr int TR_XLD = 18;           // padded column stride of the exchange tile (conflict-free LDS.128)
template <int CPT>
constexpr int tr_smem() { return (36 * 256 + 2 * 32 * CPT * TR_XLD) * (int)sizeof(double); }
// CPT = columns per thread (a CTA covers 32 * CPT columns).  With CPT = 2 every broadcast T load feeds two columns: half the
// shared-memory traffic per column, used when there are enough columns to fill the machine with 64-column CTAs.
// RSIDE (with UPPER = false): the right-side solve X T^T = B with T lower and NON-unit -- every ROW of B is a right-hand side
// (B is `ncols` rows x nb columns; the step L21 = A21 L11^-T of the blocked Cholesky).  Same kernel with the roles of the
// two indices of B swapped; those loads and stores are coalesced as they are.
template <bool UPPER, int CPT, bool RSIDE = false>
__global__ void __launch_bounds__(TR_NT) trsm_leaf_kernel(int nb, int ncols, const double *__restrict__ T, long long ldt, long long st, double *B,
                                                          long long ldb, long long sb) {
    constexpr int TR_CB = 32 * CPT;
    constexpr bool UNIT = !UPPER && !RSIDE;
    extern __shared__ double sm[];
    double *Ts = sm;             // block (bi, bk), warp r, column j, row i:  Ts[blk*256 + r*32 + j*2 + i] = T(16bi + 2r + i, 16bk + j)
    double *Xs = sm + 36 * 256;  // two exchange tiles: Xs[buf][col * TR_XLD + row]
    T += blockIdx.y * st;
    B += blockIdx.y * sb;
    const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
    const int ngroups = (ncols + TR_CB - 1) / TR_CB;  // a CTA walks the column groups blockIdx.x, blockIdx.x + gridDim.x, ...
    int grp = blockIdx.x;
    auto blk_index = [](int bi, int bk) { return UPPER ? bk * (bk + 1) / 2 + bi : bi * (bi + 1) / 2 + bk; };
    // ---- stage T (triangle only, all 36 blocks: blocks beyond nb are zero with a unit diagonal) with cp.async: every copy
    // of the CTA is in flight at once; meanwhile load B.  thread -> (pair of rows i2, column j) of a block: the 16-byte
    // destination is contiguous, the source is when T allows.  Two blocks per pass: (big, small) advance incrementally.
    const bool t16 = ((reinterpret_cast<uintptr_t>(T) & 15u) == 0) && (ldt % 2 == 0);
    {
        int big = 0, small = 0;  // block tid >> 7 of the triangle enumeration (big >= small)
        if (tid >> 7) { big = 1; small = 0; }
        const int e = tid & 127, i2 = e & 7, j = e >> 3;
        for (int blk = tid >> 7; blk < 36; blk += 2) {
            const int bi = UPPER ? small : big, bk = UPPER ? big : small;
            const int gi = bi * 16 + 2 * i2, gj = bk * 16 + j;
            double *dst = Ts + blk * 256 + i2 * 32 + j * 2;
            const double *src = T + gj * ldt + gi;
            if (t16 && gi + 1 < nb && gj < nb) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(lsmem(dst)), "l"(src) : "memory");
            } else {
                cp_async8(lsmem(dst), gi < nb && gj < nb ? src : T, gi < nb && gj < nb);
                cp_async8(lsmem(dst + 1), gi + 1 < nb && gj < nb ? src + 1 : T, gi + 1 < nb && gj < nb);
            }
            small += 2;
            while (small > big) { small -= big + 1; ++big; }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    double y[CPT][8][2];
    auto column = [&](int cp) { return grp * TR_CB + 32 * cp + lane; };
    auto load_y = [&]() {
#pragma unroll
        for (int cp = 0; cp < CPT; ++cp) {
            const int col = column(cp);
            const double *bcol = RSIDE ? B + col + (long long)(2 * r) * ldb : B + (long long)col * ldb + 2 * r;
            const long long rs = RSIDE ? ldb : 1;  // stride between consecutive unknowns of one right-hand side
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
                for (int i = 0; i < 2; ++i) y[cp][b][i] = (col < ncols && 16 * b + 2 * r + i < nb) ? bcol[(16 * b + i) * rs] : 0.0;
        }
    };
    load_y();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- invert the diagonal blocks in place: 16 lanes per block, lane j owns column j of the inverse ----
    if (r < 4) {
        const int d = r * 2 + (lane >> 4);  // 4 warps x 2 half-warps = the 8 diagonal blocks
        double *tb = Ts + blk_index(d, d) * 256;
        auto el = [&](int i, int k) -> double & { return tb[(i >> 1) * 32 + k * 2 + (i & 1)]; };
        const int j = lane & 15;
        double z[16];
        if (!UPPER) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < i) acc -= el(i, k) * ((k >= j) ? z[k] : 0.0);
                const double dg = (UNIT || 16 * d + i >= nb) ? 1.0 : el(i, i);
                z[i] = (i >= j) ? (UNIT ? acc : acc / dg) : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 15; i >= 0; --i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 15; k >= 0; --k)
                    if (k > i) acc -= el(i, k) * ((k <= j) ? z[k] : 0.0);
                const double dg = (16 * d + i < nb) ? el(i, i) : 1.0;
                z[i] = (i <= j) ? acc / dg : 0.0;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) el(i, j) = z[i];
    }
    __syncthreads();
    // ---- block steps ----
    auto read16 = [&](const double *xs, double (&x)[CPT][16]) {
#pragma unroll
        for (int cp = 0; cp < CPT; ++cp)
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const double2 t = *reinterpret_cast<const double2 *>(xs + (32 * cp + lane) * TR_XLD + 2 * h);
                x[cp][2 * h] = t.x;
                x[cp][2 * h + 1] = t.y;
            }
    };
    auto write2 = [&](double *xs, int bk) {
#pragma unroll
        for (int cp = 0; cp < CPT; ++cp)
            *reinterpret_cast<double2 *>(xs + (32 * cp + lane) * TR_XLD + 2 * r) = make_double2(y[cp][bk][0], y[cp][bk][1]);
    };
    double *XA = Xs, *XB = Xs + TR_CB * TR_XLD;
  for (;;) {  // column groups of this CTA (T and the inverted diagonal blocks are staged once)
    // The 8 block steps are fully unrolled (block coordinates are compile-time constants), so the register-resident y[b]
    // are indexed statically and the updates of all remaining blocks of a step form independent FMA chains.
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int bk = UPPER ? 7 - s : s;
        double x[CPT][16];
        write2(XA, bk);
        __syncthreads();
        read16(XA, x);
        {   // x_bk = inv(T_kk) b_k, this warp's 2 rows; four partial sums shorten the dependent chain
            const double2 *tr = reinterpret_cast<const double2 *>(Ts + blk_index(bk, bk) * 256 + r * 32);
            double z0[CPT], z1[CPT], w0[CPT], w1[CPT];
#pragma unroll
            for (int cp = 0; cp < CPT; ++cp) z0[cp] = z1[cp] = w0[cp] = w1[cp] = 0.0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const double2 ta = tr[j], tb2 = tr[j + 1];
#pragma unroll
                for (int cp = 0; cp < CPT; ++cp) {
                    z0[cp] += ta.x * x[cp][j];
                    z1[cp] += ta.y * x[cp][j];
                    w0[cp] += tb2.x * x[cp][j + 1];
                    w1[cp] += tb2.y * x[cp][j + 1];
                }
            }
#pragma unroll
            for (int cp = 0; cp < CPT; ++cp) {
                y[cp][bk][0] = z0[cp] + w0[cp];
                y[cp][bk][1] = z1[cp] + w1[cp];
            }
        }
        write2(XB, bk);
        __syncthreads();
        if (s == 7) break;
        read16(XB, x);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (UPPER ? (b < bk) : (b > bk)) {
                    const double2 t = reinterpret_cast<const double2 *>(Ts + blk_index(b, bk) * 256 + r * 32)[j];
#pragma unroll
                    for (int cp = 0; cp < CPT; ++cp) {
                        y[cp][b][0] -= t.x * x[cp][j];
                        y[cp][b][1] -= t.y * x[cp][j];
                    }
                }
            }
        }
    }
    const int next = grp + gridDim.x;
    if (next < ngroups) {  // more groups to come: T stays staged, plain stores (one 16-byte piece per block and thread)
#pragma unroll
        for (int cp = 0; cp < CPT; ++cp) {
            const int col = column(cp);
            double *bcol = RSIDE ? B + col + (long long)(2 * r) * ldb : B + (long long)col * ldb + 2 * r;
            const long long rs = RSIDE ? ldb : 1;
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (col < ncols && 16 * b + 2 * r + i < nb) bcol[(16 * b + i) * rs] = y[cp][b][i];
        }
        grp = next;
        load_y();
        continue;
    }
    if (RSIDE) {  // rows of B: the plain stores are already coalesced across the lanes
#pragma unroll
        for (int cp = 0; cp < CPT; ++cp) {
            const int col = column(cp);
            double *bcol = B + col + (long long)(2 * r) * ldb;
#pragma unroll
            for (int b = 0; b < 8; ++b)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (col < ncols && 16 * b + 2 * r + i < nb) bcol[(long long)(16 * b + i) * ldb] = y[cp][b][i];
        }
        break;
    }
    // ---- last group: transpose through shared memory (T is no longer needed) so that the stores are full 256-byte row segments ----
    double *Bs = Ts;  // Bs[col * TR_OLD + row]
    constexpr int TR_OLD = NB + 2;
#pragma unroll
    for (int cp = 0; cp < CPT; ++cp)
#pragma unroll
        for (int b = 0; b < 8; ++b)
            *reinterpret_cast<double2 *>(Bs + (32 * cp + lane) * TR_OLD + 16 * b + 2 * r) = make_double2(y[cp][b][0], y[cp][b][1]);
    __syncthreads();
    for (int idx = tid; idx < NB * TR_CB; idx += TR_NT) {
        const int i = idx & (NB - 1), c = idx >> 7;
        if (i < nb && grp * TR_CB + c < ncols) B[(long long)(grp * TR_CB + c) * ldb + i] = Bs[c * TR_OLD + i];
    }
    break;
  }
}

// ---------------------------------------------------------------------------------------------
// leaf triangular solve on the FP64 tensor pipe.  Same contract as trsm_leaf_kernel (and the same 16 x 16 blocking with
// inverted diagonal blocks), but every 16 x 16 block product is 32 DMMAs on distributed fragments instead of 512 FMAs fed by
// broadcast LDS.128 -- the FMA version is bound by the shared-memory pipe (one 16-byte broadcast per two FMAs).
// One CTA = 8 warps = 32 columns; warp w owns block-row w (rows 16w..16w+15) as C fragments (16 doubles per lane).
// Step bk: warp bk turns its block into B fragments through a small smem tile, multiplies by inv(T_kk) (A fragments straight
// from the staged T, row stride 20 doubles: conflict free), publishes -x_bk in a double-buffered tile; one __syncthreads;
// every warp still to be solved accumulates T(w, bk) (-x_bk).
// ---------------------------------------------------------------------------------------------
constexpr int TD_LP = 20;            // row stride of a staged T block
constexpr int TD_BLK = 16 * TD_LP;   // doubles per block
constexpr int TD_XLD = 36;           // row stride of the 16 x 32 exchange tiles
constexpr int TD_SMEM = (36 * TD_BLK + 3 * 16 * TD_XLD) * (int)sizeof(double);
template <bool UPPER, bool RSIDE>
__global__ void __launch_bounds__(256) trsm_leaf_dmma_kernel(int nb, int ncols, const double *__restrict__ T, long long ldt, long long st, double *B,
                                                             long long ldb, long long sb) {
    constexpr bool UNIT = !UPPER && !RSIDE;
    extern __shared__ double sm[];
    double *Ts = sm;                  // block (bi, bk): Ts[blk * TD_BLK + i * TD_LP + k] = T(16 bi + i, 16 bk + k)
    double *XC = sm + 36 * TD_BLK;    // conversion tile of the solving warp
    double *XB = XC + 16 * TD_XLD;    // two broadcast tiles
    T += blockIdx.y * st;
    B += blockIdx.y * sb;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int ngroups = (ncols + 31) / 32;
    auto blk_index = [](int bi, int bk) { return UPPER ? bk * (bk + 1) / 2 + bi : bi * (bi + 1) / 2 + bk; };
    // ---- stage the triangle of T (zeros beyond nb) ----
    {
        int big = 0, small = 0;
        const int i = tid & 15, k = tid >> 4;  // one element per thread and block; consecutive threads walk a column of T
        for (int blk = 0; blk < 36; ++blk) {
            const int bi = UPPER ? small : big, bk = UPPER ? big : small;
            const int gi = bi * 16 + i, gk = bk * 16 + k;
            const bool ok = gi < nb && gk < nb;
            cp_async8(lsmem(Ts + blk * TD_BLK + i * TD_LP + k), ok ? T + gk * ldt + gi : T, ok);
            if (++small > big) { small = 0; ++big; }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    int grp = blockIdx.x;
    double yc[2][4][2];
    auto elem = [&](int row, int col) -> double * { return RSIDE ? B + col + (long long)row * ldb : B + (long long)col * ldb + row; };
    auto load_y = [&]() {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = 16 * w + 8 * mt + g, col = grp * 32 + 8 * nt + 2 * t + e;
                    yc[mt][nt][e] = (row < nb && col < ncols) ? *elem(row, col) : 0.0;
                }
    };
    load_y();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- invert the diagonal blocks in place: 16 lanes per block, lane j owns column j of the inverse ----
    if (w < 4) {
        const int d = w * 2 + (lane >> 4);
        double *tb = Ts + blk_index(d, d) * TD_BLK;
        auto el = [&](int i, int k) -> double & { return tb[i * TD_LP + k]; };
        const int j = lane & 15;
        double z[16];
        if (!UPPER) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < i) acc -= el(i, k) * ((k >= j) ? z[k] : 0.0);
                const double dg = (UNIT || 16 * d + i >= nb) ? 1.0 : el(i, i);
                z[i] = (i >= j) ? (UNIT ? acc : acc / dg) : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 15; i >= 0; --i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 15; k >= 0; --k)
                    if (k > i) acc -= el(i, k) * ((k <= j) ? z[k] : 0.0);
                const double dg = (16 * d + i < nb) ? el(i, i) : 1.0;
                z[i] = (i <= j) ? acc / dg : 0.0;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) el(i, j) = z[i];
    }
    __syncthreads();
    auto frag_a = [&](const double *tb, double (&af)[2][4]) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) af[mt][ks] = tb[(8 * mt + g) * TD_LP + 4 * ks + t];
    };
    auto frag_b = [&](const double *xs, double (&bf)[4][4]) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) bf[ks][nt] = xs[(4 * ks + t) * TD_XLD + 8 * nt + g];
    };
    auto put_c = [&](double *xs, const double (&c)[2][4][2], double sign) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<double2 *>(xs + (8 * mt + g) * TD_XLD + 8 * nt + 2 * t) = make_double2(sign * c[mt][nt][0], sign * c[mt][nt][1]);
    };
    for (;;) {  // column groups of this CTA (T and the inverted diagonal blocks are staged once)
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
            const int bk = UPPER ? 7 - s : s;
            double *xb = XB + (s & 1) * 16 * TD_XLD;
            if (w == bk) {
                double af[2][4], bf[4][4];
                put_c(XC, yc, 1.0);
                __syncwarp();
                frag_b(XC, bf);
                frag_a(Ts + blk_index(bk, bk) * TD_BLK, af);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        yc[mt][nt][0] = yc[mt][nt][1] = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) dmma884(yc[mt][nt][0], yc[mt][nt][1], af[mt][ks], bf[ks][nt]);
                    }
                put_c(xb, yc, -1.0);  // the others accumulate T (-x)
            }
            __syncthreads();
            if (UPPER ? (w < bk) : (w > bk)) {
                double af[2][4], bf[4][4];
                frag_b(xb, bf);
                frag_a(Ts + blk_index(w, bk) * TD_BLK, af);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) dmma884(yc[mt][nt][0], yc[mt][nt][1], af[mt][ks], bf[ks][nt]);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int row = 16 * w + 8 * mt + g, col = grp * 32 + 8 * nt + 2 * t + e;
                    if (row < nb && col < ncols) *elem(row, col) = yc[mt][nt][e];
                }
        grp += gridDim.x;
        if (grp >= ngroups) break;
        __syncthreads();  // the last step's broadcast tile is still being read by slower warps
        load_y();
    }
}

// ---------------------------------------------------------------------------------------------
// leaf triangular solve, right-hand sides owned by warps (the default).  The columns of B are independent, so a warp that owns
// 8 of them never has to talk to another warp: the solve is done on the transposed system  X^T = Y^T T^-T  with the 8 columns
// as the M dimension of the m8n8k4 DMMA and the 128 unknowns as N (16 accumulator tiles, 32 doubles per lane).  The C fragment
// of a solved 16-row block (lane (g,t): column g, rows 2t, 2t+1, 8+2t, 8+2t+1) is, as it stands, a valid A fragment for the
// next product if the k index of slice s is read as {2t, 2t+1, 8+2t, 8+2t+1}[s] -- the B fragments (rows of the staged T block)
// are simply loaded with the same k order, one conflict-free LDS.128 per two slices (row stride 24 doubles).  No shared-memory
// exchange, no __syncthreads after the staging, perfectly balanced DMMA work (the block-row version above serialises on the
// solving warp and loads the four sub-partitions unevenly: 217 us against 65 us of DMMA time for 512 x (128 x 128, 256 rhs)).
// Same 16 x 16 blocking and inverted diagonal blocks as the other two leaf kernels, so the same rounding behaviour.
// ---------------------------------------------------------------------------------------------
constexpr int TC_LP = 24;            // row stride of a staged T block: 16-byte loads of a quarter warp (2 rows x 64 B) hit all banks once
constexpr int TC_BLK = 16 * TC_LP;
constexpr int TC_SMEM = 36 * TC_BLK * (int)sizeof(double);
template <bool UPPER, bool RSIDE>
__global__ void __launch_bounds__(256, 2) trsm_leaf_cols_kernel(int nb, int ncols, const double *__restrict__ T, long long ldt, long long st, double *B,
                                                                long long ldb, long long sb, int vec_ok, int32_t *viol) {
    constexpr bool UNIT = !UPPER && !RSIDE;
    extern __shared__ double sm[];
    double *Ts = sm;  // block (bi, bk): Ts[blk * TC_BLK + i * TC_LP + k] = T(16 bi + i, 16 bk + k); diagonal blocks hold their inverse
    T += blockIdx.y * st;
    B += blockIdx.y * sb;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    auto blk_index = [](int bi, int bk) { return UPPER ? bk * (bk + 1) / 2 + bi : bi * (bi + 1) / 2 + bk; };
    const int nblk = (nb + 15) >> 4;
    {
        int big = 0, small = 0;
        const int i = tid & 15, k = tid >> 4;  // consecutive threads walk a column of T
        for (int blk = 0; blk < 36; ++blk) {
            const int bi = UPPER ? small : big, bk = UPPER ? big : small;
            const int gi = bi * 16 + i, gk = bk * 16 + k;
            const bool ok = gi < nb && gk < nb;
            if (big < nblk) cp_async8(lsmem(Ts + blk * TC_BLK + i * TC_LP + k), ok ? T + gk * ldt + gi : T, ok);
            if (++small > big) { small = 0; ++big; }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- invert the diagonal blocks in place: 16 lanes per block, lane j owns column j of the inverse ----
    if (w < 4) {
        const int d = w * 2 + (lane >> 4);
        double *tb = Ts + blk_index(d, d) * TC_BLK;
        auto el = [&](int i, int k) -> double & { return tb[i * TC_LP + k]; };
        const int j = lane & 15;
        double z[16];
        if (d < nblk) {
            if (!UPPER) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < i) acc -= el(i, k) * ((k >= j) ? z[k] : 0.0);
                    const double dg = (UNIT || 16 * d + i >= nb) ? 1.0 : el(i, i);
                    z[i] = (i >= j) ? (UNIT ? acc : acc / dg) : 0.0;
                }
            } else {
#pragma unroll
                for (int i = 15; i >= 0; --i) {
                    double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                    for (int k = 15; k >= 0; --k)
                        if (k > i) acc -= el(i, k) * ((k <= j) ? z[k] : 0.0);
                    const double dg = (16 * d + i < nb) ? el(i, i) : 1.0;
                    z[i] = (i <= j) ? acc / dg : 0.0;
                }
            }
        }
        __syncwarp();
        if (d < nblk) {
#pragma unroll
            for (int i = 0; i < 16; ++i) el(i, j) = z[i];
        }
    }
    __syncthreads();
    // B fragments of one 16 x 16 block for both 8-wide n-tiles: bf[nt2][s] = Tb[8 nt2 + g][k_s(t)]
    auto frag_b = [&](const double *tb, double (&bf)[2][4]) {
#pragma unroll
        for (int nt2 = 0; nt2 < 2; ++nt2) {
            const double2 lo = *reinterpret_cast<const double2 *>(tb + (8 * nt2 + g) * TC_LP + 2 * t);
            const double2 hi = *reinterpret_cast<const double2 *>(tb + (8 * nt2 + g) * TC_LP + 8 + 2 * t);
            bf[nt2][0] = lo.x; bf[nt2][1] = lo.y; bf[nt2][2] = hi.x; bf[nt2][3] = hi.y;
        }
    };
    const int chunks = (ncols + 63) / 64;
    for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
        const int col = chunk * 64 + 8 * w + g;  // this lane's right-hand side
        if (chunk * 64 + 8 * w >= ncols) continue;  // warp-uniform
        double wc[16][2];  // wc[nt][e] = Y(8 nt + 2 t + e, col)
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const int row = 8 * nt + 2 * t;
            wc[nt][0] = wc[nt][1] = 0.0;
            if (col < ncols && nt < 2 * nblk) {
                if (!RSIDE && vec_ok && row + 1 < nb) {
                    const double2 v = *reinterpret_cast<const double2 *>(B + (long long)col * ldb + row);
                    wc[nt][0] = v.x; wc[nt][1] = v.y;
                } else {
                    if (row < nb) wc[nt][0] = RSIDE ? B[col + (long long)row * ldb] : B[(long long)col * ldb + row];
                    if (row + 1 < nb) wc[nt][1] = RSIDE ? B[col + (long long)(row + 1) * ldb] : B[(long long)col * ldb + row + 1];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int bk = UPPER ? 7 - s : s;
            if (bk >= nblk) continue;  // block rows beyond nb (short leaves): warp-uniform
            double bf[2][4];
            // x_bk^T = y_bk^T inv(T_kk)^T
            frag_b(Ts + blk_index(bk, bk) * TC_BLK, bf);
            const double a0 = wc[2 * bk][0], a1 = wc[2 * bk][1], a2 = wc[2 * bk + 1][0], a3 = wc[2 * bk + 1][1];
            {
                double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;  // the two n-tiles interleaved
                dmma884(c0, c1, a0, bf[0][0]); dmma884(d0, d1, a0, bf[1][0]);
                dmma884(c0, c1, a1, bf[0][1]); dmma884(d0, d1, a1, bf[1][1]);
                dmma884(c0, c1, a2, bf[0][2]); dmma884(d0, d1, a2, bf[1][2]);
                dmma884(c0, c1, a3, bf[0][3]); dmma884(d0, d1, a3, bf[1][3]);
                wc[2 * bk][0] = c0; wc[2 * bk][1] = c1; wc[2 * bk + 1][0] = d0; wc[2 * bk + 1][1] = d1;
            }
            const double n0 = -wc[2 * bk][0], n1 = -wc[2 * bk][1], n2 = -wc[2 * bk + 1][0], n3 = -wc[2 * bk + 1][1];
            // y_bi^T -= x_bk^T T(bi, bk)^T for the block rows still to be solved (the next one to be solved first), two block
            // rows = four accumulators at a time, slice-major, so that consecutive DMMAs never depend on each other
            const double ng[4] = {n0, n1, n2, n3};
#pragma unroll
            for (int q = 1; q < 8; q += 2) {
                double bf2[2][2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int bi = UPPER ? bk - (q + j) : bk + (q + j);
                    if (bi < 0 || bi > 7) continue;  // compile time after unrolling
                    if (bi < nblk) frag_b(Ts + blk_index(bi, bk) * TC_BLK, bf2[j]);
                }
#pragma unroll
                for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int bi = UPPER ? bk - (q + j) : bk + (q + j);
                        if (bi < 0 || bi > 7) continue;
                        if (bi >= nblk) continue;  // warp-uniform
#pragma unroll
                        for (int nt2 = 0; nt2 < 2; ++nt2) dmma884(wc[2 * bi + nt2][0], wc[2 * bi + nt2][1], ng[sl], bf2[j][nt2][sl]);
                    }
            }
        }
        if (RSIDE && viol != nullptr) {
            // the no-interchange LU (k_getrf_nopiv): these are its multipliers L21; any |l| > 0.999 (or NaN) means partial pivoting
            // would have interchanged and the factorization must be discarded
            bool bad = false;
#pragma unroll
            for (int nt = 0; nt < 16; ++nt)
                if (col < ncols && nt < 2 * nblk) bad = bad || !(fabs(wc[nt][0]) <= 0.999) || !(fabs(wc[nt][1]) <= 0.999);
            if (__any_sync(FULL, bad) && lane == 0) atomicOr(viol, 1);
        }
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const int row = 8 * nt + 2 * t;
            if (col < ncols && nt < 2 * nblk) {
                if (!RSIDE && vec_ok && row + 1 < nb) {
                    *reinterpret_cast<double2 *>(B + (long long)col * ldb + row) = make_double2(wc[nt][0], wc[nt][1]);
                } else {
                    if (row < nb) (RSIDE ? B[col + (long long)row * ldb] : B[(long long)col * ldb + row]) = wc[nt][0];
                    if (row + 1 < nb) (RSIDE ? B[col + (long long)(row + 1) * ldb] : B[(long long)col * ldb + row + 1]) = wc[nt][1];
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Whole-triangle solve for FEW right-hand sides (dgetrs with nrhs <= 256: `solve`/`mldivide` with a handful of columns,
// BASELINE config 3).  With 64 columns a blocked sweep is one leaf solve + one thin GEMM per 128 rows, 2 x 64 dependent
// launches of ~18 us each for n = 8192; the flops are nothing, the chain of launches is everything.  Here the whole sweep is
// ONE launch: CTA r owns block row r (128 rows x 64 columns, kept in registers as in trsm_leaf_cols_kernel: 8 columns per
// warp, C fragments), streams the tiles T(r, k) of its block row through a 3-deep cp.async ring (the tiles do not depend on
// the solution, so they are prefetched ahead of the wavefront), and for every k waits for CTA k to publish x_k
// (release/acquire flag in global memory), subtracts T(r,k) x_k on the DMMA pipe, finally solves with its own diagonal tile
// (16 x 16 diagonal blocks inverted up front, off the critical path) and publishes x_r.  All CTAs are co-resident
// (grid <= SM count, checked by the host), so the spin waits cannot deadlock; they are bounded anyway.
// Same blocking and inverted 16 x 16 diagonal blocks as the leaf kernels: same rounding behaviour.
// ---------------------------------------------------------------------------------------------
constexpr int TW_LD = 132;                 // column stride of a staged 128 x 64 half tile: LDS.64 fragment loads hit every bank once per wavefront
constexpr int TW_HALF = 64 * TW_LD;
constexpr int TW_DLP = 24;                 // row stride of the inverted diagonal blocks (LDS.128, as TC_LP)
constexpr int TW_SMEM = (3 * TW_HALF + 8 * 16 * TW_DLP) * (int)sizeof(double);
template <bool UPPER, int WPC, bool PAIR>
__global__ void __launch_bounds__(32 * WPC * (PAIR ? 2 : 1), 1) trsm_wave_kernel(int n, int ncols, const double *__restrict__ T, long long ldt, double *B, long long ldb,
                                                           int *flags, int epoch, int vec_ok, int tvec_ok) {
    constexpr bool UNIT = !UPPER;
    // CW right-hand sides per CTA: 8 per owner warp.  PAIR: every owner warp w has a helper warp w + WPC on the same sub-partition that
    // takes rows 64..127 of the off-diagonal updates (its own running sum, handed over once before the diagonal tile): one warp per
    // sub-partition issues a DMMA every ~26 cycles (4-deep dependent chains), two interleave to the 16-cycle rate of the pipe.
    constexpr int THREADS = 32 * WPC * (PAIR ? 2 : 1), CW = 8 * WPC;
    extern __shared__ double sm[];
    double *ring = sm;
    double *Dinv = sm + 3 * TW_HALF;  // Dinv[d * 16 * TW_DLP + i * TW_DLP + k] = inv(T_rr block d)(i, k)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int nblk = gridDim.x, cchunks = gridDim.y, chunk = blockIdx.y;
    const int r = UPPER ? nblk - 1 - (int)blockIdx.x : (int)blockIdx.x;  // low block index = early in the wavefront
    const int rbase = r * 128;
    const int nk = UPPER ? nblk - 1 - r : r;  // off-diagonal tiles of this block row, in the order their x_k become available
    const int njobs = 2 * (nk + 1);           // half tiles, the diagonal tile last
    auto issue = [&](int q) {
        if (q < njobs) {
            const int i = q >> 1, k = i < nk ? (UPPER ? nblk - 1 - i : i) : r, h = UPPER ? 1 - (q & 1) : (q & 1);
            double *dst = ring + (q % 3) * TW_HALF;
            const int cbase = k * 128 + 64 * h;
            if (tvec_ok && rbase + 128 <= n && cbase + 64 <= n) {  // full, 16-byte aligned half tile: 16-byte copies
                const int rp = tid & 63;
#pragma unroll 8
                for (int jj = 0; jj < 64 * 64 / THREADS; ++jj) {
                    const int col = (tid >> 6) + (THREADS / 64) * jj;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lsmem(dst + col * TW_LD + 2 * rp)),
                                 "l"(T + (long long)(cbase + col) * ldt + rbase + 2 * rp) : "memory");
                }
            } else {  // edge tiles of a ragged n (zero fill) and unaligned triangles
                const int row = tid & 127, gr = rbase + row;
#pragma unroll 8
                for (int jj = 0; jj < 64 * 128 / THREADS; ++jj) {
                    const int col = (tid >> 7) + (THREADS / 128) * jj, gc = cbase + col;
                    const bool ok = gr < n && gc < n;
                    cp_async8(lsmem(dst + col * TW_LD + row), ok ? T + (long long)gc * ldt + gr : T, ok);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto landed = [&](int) { asm volatile("cp.async.wait_group 1;" ::: "memory"); };  // this thread's part of half tile q is in shared memory
    issue(0);
    issue(1);
    // ---- the diagonal 16 x 16 blocks of T_rr, inverted (identity beyond n) ----
    for (int e = tid; e < 8 * 256; e += THREADS) {
        const int d = e >> 8, i = e & 15, k = (e >> 4) & 15, gi = rbase + 16 * d + i, gk = rbase + 16 * d + k;
        Dinv[d * 16 * TW_DLP + i * TW_DLP + k] = (gi < n && gk < n) ? T[(long long)gk * ldt + gi] : 0.0;
    }
    __syncthreads();
    if (w < 4) {
        const int d = w * 2 + (lane >> 4), j = lane & 15;
        double *tb = Dinv + d * 16 * TW_DLP;
        auto el = [&](int i, int k) -> double & { return tb[i * TW_DLP + k]; };
        double z[16];
        if (!UPPER) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    if (k < i) acc -= el(i, k) * ((k >= j) ? z[k] : 0.0);
                const double dg = (UNIT || rbase + 16 * d + i >= n) ? 1.0 : el(i, i);
                z[i] = (i >= j) ? (UNIT ? acc : acc / dg) : 0.0;
            }
        } else {
#pragma unroll
            for (int i = 15; i >= 0; --i) {
                double acc = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 15; k >= 0; --k)
                    if (k > i) acc -= el(i, k) * ((k <= j) ? z[k] : 0.0);
                const double dg = (rbase + 16 * d + i < n) ? el(i, i) : 1.0;
                z[i] = (i <= j) ? acc / dg : 0.0;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) el(i, j) = z[i];
    }
    // ---- this warp's 8 right-hand sides of block row r ----
    const bool helper = PAIR && w >= WPC;
    const int cw = helper ? w - WPC : w;
    const int col = chunk * CW + 8 * cw + g;
    const bool have = chunk * CW + 8 * cw < ncols;  // warp-uniform; idle warps still stage tiles and meet the barriers
    auto ld2 = [&](int row, double &v0, double &v1) {  // rows row, row+1 of this lane's column, bypassing L1 (x_k comes from another SM)
        v0 = v1 = 0.0;
        if (col < ncols) {
            const double *p = B + (long long)col * ldb + row;
            if (vec_ok && row + 1 < n) {
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
                v0 = v.x; v1 = v.y;
            } else {
                if (row < n) v0 = __ldcg(p);
                if (row + 1 < n) v1 = __ldcg(p + 1);
            }
        }
    };
    double wc[16][2];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
        wc[nt][0] = wc[nt][1] = 0.0;
        if (have && !helper) ld2(rbase + 8 * nt + 2 * t, wc[nt][0], wc[nt][1]);  // a helper starts from zero: wc[8..15] is its running sum
    }
    PROFW_DECL;
    for (int i = 0; i <= nk; ++i) {
        const bool diag = i == nk;
        const int k = diag ? r : (UPPER ? nblk - 1 - i : i);
        double xf[16][2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int q = 2 * i + hh, h = UPPER ? 1 - hh : hh;
            landed(q);
            __syncthreads();  // half tile q has landed for every thread; everyone is done with half tile q-1; x_k is visible
            PROFW(diag ? 21 : 17);
            issue(q + 2);  // into the buffer of half tile q-1; ahead of the flag wait: the tiles do not depend on the solution
            PROFW(diag ? 22 : 18);
            if (hh == 0 && !diag) {
                if (tid == 0) {  // x_k published?
                    const int *f = flags + k * cchunks + chunk;
                    int v, spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    } while (v != epoch && ++spins < (1 << 22));
                    if (v != epoch) asm volatile("trap;");  // seconds without progress under a cooperative launch: fail loudly, never continue on stale data
                }
                PROFW_ON(false); PROFW(16); PROFW_ON(nk > 0 && i >= nk - 1);
                __syncthreads();  // x_k is visible to every thread
                if (have) {       // one L2 round trip for the whole x_k
#pragma unroll
                    for (int nt = 0; nt < 16; ++nt) ld2(k * 128 + 8 * nt + 2 * t, xf[nt][0], xf[nt][1]);
                }
                PROFW(20);
            }
            const double *buf = ring + (q % 3) * TW_HALF;
            if (PAIR && diag && hh == 0 && nk > 0) {
                // hand-over of the helpers' sums through the ring slot of half tile q-1 (free: nothing is staged after the diagonal tile)
                double *xch = ring + ((q + 2) % 3) * TW_HALF + (cw * 32 + lane) * 16;
                if (helper && have) {
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<double2 *>(xch + 2 * nt) = make_double2(wc[8 + nt][0], wc[8 + nt][1]);
                }
                __syncthreads();
                if (!helper && have) {
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const double2 hv = *reinterpret_cast<const double2 *>(xch + 2 * nt);
                        wc[8 + nt][0] += hv.x; wc[8 + nt][1] += hv.y;
                    }
                }
            }
            if (!have || (helper && diag)) continue;
            if (!diag) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int xn = 8 * h + 2 * kk;
                    const double ng[4] = {-xf[xn][0], -xf[xn][1], -xf[xn + 1][0], -xf[xn + 1][1]};
                    const double *cb = buf + (16 * kk + 2 * t) * TW_LD + g;
                    // four accumulators at a time, slice-major: consecutive DMMAs never depend on each other
#pragma unroll
                    for (int nt0 = 0; nt0 < 16; nt0 += 4) {
                        if (PAIR && (helper ? nt0 < 8 : nt0 >= 8)) continue;  // warp-uniform: owner rows 0..63, helper rows 64..127
                        double bv[4][4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            bv[j][0] = cb[8 * (nt0 + j)]; bv[j][1] = cb[TW_LD + 8 * (nt0 + j)];
                            bv[j][2] = cb[8 * TW_LD + 8 * (nt0 + j)]; bv[j][3] = cb[9 * TW_LD + 8 * (nt0 + j)];
                        }
#pragma unroll
                        for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                            for (int j = 0; j < 4; ++j) dmma884(wc[nt0 + j][0], wc[nt0 + j][1], ng[sl], bv[j][sl]);
                    }
                }
                PROFW(19);
                continue;
            }
            // ---- diagonal tile: the four block steps whose columns are in this half ----
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const int bk = 4 * h + (UPPER ? 3 - s4 : s4);
                {
                    const double *db = Dinv + bk * 16 * TW_DLP;
                    const double a0 = wc[2 * bk][0], a1 = wc[2 * bk][1], a2 = wc[2 * bk + 1][0], a3 = wc[2 * bk + 1][1];
                    const double2 lo0 = *reinterpret_cast<const double2 *>(db + g * TW_DLP + 2 * t);
                    const double2 hi0 = *reinterpret_cast<const double2 *>(db + g * TW_DLP + 8 + 2 * t);
                    const double2 lo1 = *reinterpret_cast<const double2 *>(db + (8 + g) * TW_DLP + 2 * t);
                    const double2 hi1 = *reinterpret_cast<const double2 *>(db + (8 + g) * TW_DLP + 8 + 2 * t);
                    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
                    dmma884(c0, c1, a0, lo0.x); dmma884(d0, d1, a0, lo1.x);
                    dmma884(c0, c1, a1, lo0.y); dmma884(d0, d1, a1, lo1.y);
                    dmma884(c0, c1, a2, hi0.x); dmma884(d0, d1, a2, hi1.x);
                    dmma884(c0, c1, a3, hi0.y); dmma884(d0, d1, a3, hi1.y);
                    wc[2 * bk][0] = c0; wc[2 * bk][1] = c1; wc[2 * bk + 1][0] = d0; wc[2 * bk + 1][1] = d1;
                }
                const double ng[4] = {-wc[2 * bk][0], -wc[2 * bk][1], -wc[2 * bk + 1][0], -wc[2 * bk + 1][1]};
                const double *cb = buf + (16 * (bk & 3) + 2 * t) * TW_LD + g;
                // the block rows still to be solved, nearest first, four accumulators at a time
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    double bv[4][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int nt = UPPER ? 2 * bk - 1 - (4 * q4 + j) : 2 * bk + 2 + 4 * q4 + j;
                        if (nt < 0 || nt > 15) continue;  // compile time
                        bv[j][0] = cb[8 * nt]; bv[j][1] = cb[TW_LD + 8 * nt]; bv[j][2] = cb[8 * TW_LD + 8 * nt]; bv[j][3] = cb[9 * TW_LD + 8 * nt];
                    }
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int nt = UPPER ? 2 * bk - 1 - (4 * q4 + j) : 2 * bk + 2 + 4 * q4 + j;
                            if (nt < 0 || nt > 15) continue;
                            dmma884(wc[nt][0], wc[nt][1], ng[sl], bv[j][sl]);
                        }
                }
            }
            PROFW(23);
        }
    }
    // ---- publish x_r ----
    if (have && !helper && col < ncols) {
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            const int row = rbase + 8 * nt + 2 * t;
            double *p = B + (long long)col * ldb + row;
            if (vec_ok && row + 1 < n) *reinterpret_cast<double2 *>(p) = make_double2(wc[nt][0], wc[nt][1]);
            else {
                if (row < n) p[0] = wc[nt][0];
                if (row + 1 < n) p[1] = wc[nt][1];
            }
        }
    }
    PROFW(24);
    __syncthreads();
    PROFW(25);
    if (tid == 0) {  // the barrier orders every thread's stores before this thread's fence; the fence is cumulative
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + r * cchunks + chunk), "r"(epoch) : "memory");
    }
    PROFW(26);
#ifdef FM_LU_PROF
    if (tid == 0 && chunk == 0 && nk > 0) atomicAdd((unsigned long long *)&g_prof[27], 1ull);
#endif
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
// Panels taller than 16384 rows (more than the 16 x 1024 rows a cluster can hold in registers): plain DGETF2 with one IDAMAX
// reduction and two small kernels per column, rows physically interchanged inside the panel.  Slow (launch bound, ~2.5 ms
// per panel) but only the first (m - 16384)/128 panels of a very large factorization take this path.
// ---------------------------------------------------------------------------------------------
// one CTA: turn the IDAMAX result into ipiv / info, interchange rows j and p across the panel's columns, publish p and 1/pivot
__global__ void __launch_bounds__(NB) tall_pivot_kernel(double *a, long long lda, int pw, int j, int grow0, const long long *amax_idx, int32_t *ipiv,
                                                         int32_t *info, double *pivinfo /* [0] = pivot value */) {
    const int p = j + (int)amax_idx[0];  // panel-local row of the pivot
    const int c = threadIdx.x;
    if (c < pw && p != j) {
        const double t = a[c * lda + j];
        a[c * lda + j] = a[c * lda + p];
        a[c * lda + p] = t;
    }
    __syncthreads();
    if (c == 0) {
        const double pv = a[j * lda + j];
        ipiv[grow0 + j] = grow0 + p + 1;
        if (pv == 0.0 && *info == 0) *info = grow0 + j + 1;
        pivinfo[0] = pv;
    }
}
// rows below the pivot: multiplier (DGETF2's reciprocal rule) and the rank-1 update of the remaining panel columns
__global__ void __launch_bounds__(256) tall_update_kernel(double *a, long long lda, int m, int pw, int j, const double *pivinfo) {
    __shared__ double urow[NB];
    const double pv = pivinfo[0];
    for (int c = threadIdx.x; c < pw; c += 256) urow[c] = c > j ? a[c * lda + j] : 0.0;
    __syncthreads();
    if (pv == 0.0) return;
    const bool use_rcp = fabs(pv) >= DBL_MIN;
    const double rcp = 1.0 / pv;
    for (int r = j + 1 + blockIdx.x * 256 + threadIdx.x; r < m; r += gridDim.x * 256) {
        const double x = a[j * lda + r];
        const double l = use_rcp ? x * rcp : x / pv;
        a[j * lda + r] = l;
        for (int c = j + 1; c < pw; ++c) a[c * lda + r] -= l * urow[c];
    }
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

// device workspace for the row maps (lives for the life of the library, grown on demand)
struct MapPool {
    int2 *buf = nullptr;
    size_t entries = 0;
    int2 *get(size_t want) {
        if (want <= entries) return buf;
        if (buf) cudaFree(buf);  // implicit device sync: nothing in flight still uses it
        buf = nullptr;
        entries = 0;
        if (cudaMalloc(&buf, want * sizeof(int2)) != cudaSuccess) {
            cudaGetLastError();
            set_error(FM_ERR_NOMEM, "LU row-map workspace of %zu bytes failed", want * sizeof(int2));
            return nullptr;
        }
        entries = want;
        return buf;
    }
};
PerDevice<MapPool> g_maps_getrf, g_maps_getrs;

struct Maps {
    int2 *base;
    long long item_stride;  // in int2 units
    int2 *panel(int j0) const { return base + (long long)(j0 / NB) * MAPN; }
};

template <int NT, int RPT, int MINB>
constexpr int panel_dyn_smem() { return (2 * LEAF * UP + (MINB == 1 ? NT * RPT * LP : 0)) * (int)sizeof(double); }

template <int NT, int RPT, int MINB = 1, bool NOPIV = false>
int launch_panel(const Mat &A, int j0, int m, int pw, int32_t *ipiv, long long sp, int32_t *info, const Maps &maps, int csize,
                 int32_t *viol = nullptr) {
    static PerDevice<bool> attr_pd;  // function attributes are per device
    bool &attr = attr_pd();
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(lu_panel_kernel<NT, RPT, MINB, NOPIV>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        FM_CUDA(cudaFuncSetAttribute(lu_panel_kernel<NT, RPT, MINB, NOPIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_dyn_smem<NT, RPT, MINB>()));
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, A.batch, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = panel_dyn_smem<NT, RPT, MINB>();
    cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    FM_CUDA(cudaLaunchKernelEx(&cfg, lu_panel_kernel<NT, RPT, MINB, NOPIV>, A.at(j0, j0), (long long)A.lda, A.stride, m, pw, j0, ipiv, sp, info,
                               maps.panel(j0), maps.item_stride, viol));
    count_launch();
    return FM_OK;
}

int panel_tall(const Mat &A, int j0, int m, int pw, int32_t *ipiv, int32_t *info, const Maps &maps);
int panel_tall_entry(const Mat &A, int j0, int m, int pw, int32_t *ipiv, int32_t *info, const Maps &maps, int *csize_out);

// factor the m x pw panel whose top-left element is A(j0, j0): one cluster launch.  One row per thread whenever the cluster
// (<= 16 CTAs) can cover the panel: the column step is latency bound, so fewer rows per thread is faster.
// viol != nullptr: the no-interchange variant of the cluster kernel (single matrices, m <= 16384; see lu_panel_kernel<.., NOPIV>)
int panel_factor(const Mat &A, int j0, int m, int pw, int32_t *ipiv, long long sp, int32_t *info, const Maps &maps, int *csize_out = nullptr,
                 int32_t *viol = nullptr) {
    if (viol) {
        auto pow2 = [](int rows, int per_cta) { int cs = 1; while (cs * per_cta < rows) cs *= 2; return cs; };
        if (A.batch != 1 || m > 16384) return set_error(FM_ERR_ARG, "getrf_nopiv: no-interchange cluster panels are for single matrices of <= 16384 rows");
        if (csize_out) *csize_out = m <= 2048 ? pow2(m, 128) : (m <= 4096 ? pow2(m, 256) : 16);
        if (m <= 2048) return launch_panel<128, 1, 1, true>(A, j0, m, pw, ipiv, sp, info, maps, pow2(m, 128), viol);
        if (m <= 4096) return launch_panel<256, 1, 1, true>(A, j0, m, pw, ipiv, sp, info, maps, pow2(m, 256), viol);
        if (m <= 8192) return launch_panel<256, 2, 1, true>(A, j0, m, pw, ipiv, sp, info, maps, 16, viol);
        return launch_panel<256, 4, 1, true>(A, j0, m, pw, ipiv, sp, info, maps, 16, viol);
    }
    if (const char *tm = getenv("FLOMAT_LU_TALL_MIN"))  // test hook: send panels of more than this many rows to the tall fallback
        if (m > atoi(tm) && A.batch == 1) return panel_tall_entry(A, j0, m, pw, ipiv, info, maps, csize_out);
    auto pow2_ctas = [](int rows, int per_cta) {
        int cs = 1;
        while (cs * per_cta < rows) cs *= 2;
        return cs;
    };
    if (csize_out) *csize_out = m <= 4096 ? pow2_ctas(m, 256) : 16;
    // many small independent panels (batched expm): a lean 128-thread variant, four CTAs per SM, one wave for 592 items
    if (A.batch >= 64 && m <= 256) return launch_panel<128, 2, 4>(A, j0, m, pw, ipiv, sp, info, maps, 1);
    static const int solo_max = getenv("FLOMAT_LU_SOLO_MAX") ? atoi(getenv("FLOMAT_LU_SOLO_MAX")) : 256;  // developer switch
    if (m <= solo_max && m > 256) {  // one CTA, no cluster exchange, more rows per thread
        if (csize_out) *csize_out = 1;
        if (m <= 512) return launch_panel<256, 2>(A, j0, m, pw, ipiv, sp, info, maps, 1);
        if (m <= 1024) return launch_panel<256, 4>(A, j0, m, pw, ipiv, sp, info, maps, 1);
    }
    // 257..2048 rows: 128-row CTAs, i.e. twice the SMs of the 256-row shape.  The in-panel rank-16 update streams each CTA's rows of
    // the remaining panel columns through L2 at the per-SM rate (~24 B/clk), so more SMs is more bandwidth: getrf n = 2048 3.54 -> 3.18 ms,
    // n = 1024 1.69 -> 1.54 ms (64-row CTAs lose again: 1.59 ms).  FLOMAT_LU_NO_SMALL_WIDE restores the 256-row shape for A/B timing.
    static const bool wide_small = getenv("FLOMAT_LU_NO_SMALL_WIDE") == nullptr;
    if (wide_small && m <= 2048 && m > 256) {
        if (csize_out) *csize_out = pow2_ctas(m, 128);
        return launch_panel<128, 1>(A, j0, m, pw, ipiv, sp, info, maps, pow2_ctas(m, 128));
    }
    if (m <= 4096) return launch_panel<256, 1>(A, j0, m, pw, ipiv, sp, info, maps, pow2_ctas(m, 256));
    if (m <= 8192) {
        static const char *variant = getenv("FLOMAT_LU_PANEL");  // developer switch for A/B timing: 512 | 128
        if (variant && atoi(variant) == 512) return launch_panel<512, 1>(A, j0, m, pw, ipiv, sp, info, maps, 16);
        if (variant && atoi(variant) == 128) return launch_panel<128, 4>(A, j0, m, pw, ipiv, sp, info, maps, 16);
        return launch_panel<256, 2>(A, j0, m, pw, ipiv, sp, info, maps, 16);
    }
    if (m <= 16384) return launch_panel<256, 4>(A, j0, m, pw, ipiv, sp, info, maps, 16);
    return panel_tall_entry(A, j0, m, pw, ipiv, info, maps, csize_out);
}

int panel_tall_entry(const Mat &A, int j0, int m, int pw, int32_t *ipiv, int32_t *info, const Maps &maps, int *csize_out) {
    if (csize_out) *csize_out = 0;  // 0 = the tall fallback: no cluster, own columns already interchanged
    return panel_tall(A, j0, m, pw, ipiv, info, maps);
}

// m > 16384: unblocked panel with explicit interchanges (see tall_pivot_kernel); the panel's own columns end up permuted,
// its map (for the other columns) is rebuilt from ipiv
int panel_tall(const Mat &A, int j0, int m, int pw, int32_t *ipiv, int32_t *info, const Maps &maps) {
    if (A.batch != 1) return set_error(FM_ERR_ARG, "getrf: batched panels taller than 16384 rows are not supported (m=%d)", m);
    double *a = A.at(j0, j0);
    long long *d_idx = reinterpret_cast<long long *>(ctx().d_scalar + 8);
    double *d_piv = ctx().d_scalar + 10;
    for (int j = 0; j < pw; ++j) {
        FM_TRY(k_iamax_device(m - j, a + (size_t)j * A.lda + j, 1, d_idx));
        tall_pivot_kernel<<<1, NB, 0, ctx().stream>>>(a, A.lda, pw, j, j0, d_idx, ipiv, info, d_piv);
        FM_LAUNCH_CHECK();
        if (m - j - 1 > 0) {
            tall_update_kernel<<<std::min(ceil_div(m - j - 1, 256), 4 * ctx().sm_count), 256, 0, ctx().stream>>>(a, A.lda, m, pw, j, d_piv);
            FM_LAUNCH_CHECK();
        }
    }
    // the net permutation of this block of pivots, in the format the cluster kernel emits (global row numbers)
    // (`n` bounds the pivots replayed in the block: j0 + pw; every interchange target of this panel is a valid row)
    ipiv_to_maps_kernel<<<dim3(1, 1), 32, 0, ctx().stream>>>(j0 + m, ipiv, 0, maps.base, maps.item_stride, j0 / NB, pw);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

// apply the row permutations of the panels covering pivots [p0, p1) to columns [c0, c0+ncols) of A
int permute_rows(const Mat &A, int c0, int ncols, const Maps &maps, int p0, int p1, bool lower_cols = false) {
    if (ncols <= 0 || p1 <= p0) return FM_OK;
    const int npanels = ceil_div(p1, NB) - p0 / NB;
    dim3 grid(ceil_div(ncols, PERM_PC), A.batch);
    permute_rows_kernel<<<grid, MAPN, 0, ctx().stream>>>(A.at(0, c0), A.lda, A.stride, ncols, maps.panel(p0), maps.item_stride, npanels,
                                                         lower_cols ? 1 : 0);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

template <bool UPPER, int CPT, bool RSIDE = false>
int trsm_leaf_launch(int nb, int ncols, const double *T, int ldt, long long st, double *B, int ldb, long long sb, int batch) {
    static PerDevice<bool> attr_pd;  // function attributes are per device
    bool &attr = attr_pd();
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(trsm_leaf_kernel<UPPER, CPT, RSIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tr_smem<CPT>()));
        attr = true;
    }
    // enough CTAs to fill the machine (2 per SM); beyond that a CTA walks several column groups with T staged once
    const int groups = ceil_div(ncols, 32 * CPT);
    int gx = ceil_div(2 * ctx().sm_count, batch);
    if (gx > groups) gx = groups;
    if (gx < 1) gx = 1;
    dim3 grid(gx, batch);
    trsm_leaf_kernel<UPPER, CPT, RSIDE><<<grid, TR_NT, tr_smem<CPT>(), ctx().stream>>>(nb, ncols, T, ldt, st, B, ldb, sb);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

template <bool UPPER, bool RSIDE>
int trsm_leaf_dmma_launch(int nb, int ncols, const double *T, int ldt, long long st, double *B, int ldb, long long sb, int batch) {
    static PerDevice<bool> attr_pd;  // function attributes are per device
    bool &attr = attr_pd();
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(trsm_leaf_dmma_kernel<UPPER, RSIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM));
        attr = true;
    }
    const int groups = ceil_div(ncols, 32);
    int gx = ceil_div(2 * ctx().sm_count, batch);  // 2 CTAs per SM fill the machine; beyond that a CTA walks several groups
    if (gx > groups) gx = groups;
    if (gx < 1) gx = 1;
    trsm_leaf_dmma_kernel<UPPER, RSIDE><<<dim3(gx, batch), 256, TD_SMEM, ctx().stream>>>(nb, ncols, T, ldt, st, B, ldb, sb);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

template <bool UPPER, bool RSIDE>
int trsm_leaf_cols_launch(int nb, int ncols, const double *T, int ldt, long long st, double *B, int ldb, long long sb, int batch,
                          int32_t *viol = nullptr) {
    static PerDevice<bool> attr_pd;  // function attributes are per device
    bool &attr = attr_pd();
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(trsm_leaf_cols_kernel<UPPER, RSIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        attr = true;
    }
    const int chunks = ceil_div(ncols, 64);
    int gx = ceil_div(2 * ctx().sm_count, batch);  // 2 CTAs per SM fill the machine; beyond that a CTA walks several 64-column chunks
    if (gx > chunks) gx = chunks;
    if (gx < 1) gx = 1;
    const int vec_ok = ((reinterpret_cast<uintptr_t>(B) & 15u) == 0 && (ldb % 2) == 0 && (sb % 2) == 0) ? 1 : 0;
    trsm_leaf_cols_kernel<UPPER, RSIDE><<<dim3(gx, batch), 256, TC_SMEM, ctx().stream>>>(nb, ncols, T, ldt, st, B, ldb, sb, vec_ok, viol);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int trsm_variant() {  // developer switch: 0 = right-hand sides per warp (default), 1 = block rows per warp (DMMA), 2 = FMA leaf
    static const int v = getenv("FLOMAT_TRSM_FMA") ? 2 : getenv("FLOMAT_TRSM_ROWS") ? 1 : 0;
    return v;
}
bool trsm_use_fma() {
    static const bool fma = getenv("FLOMAT_TRSM_FMA") != nullptr;  // developer switch: the FMA leaf kernel
    return fma;
}

template <bool UPPER>
int trsm_leaf(int nb, int ncols, const double *T, int ldt, long long st, double *B, int ldb, long long sb, int batch) {
    if (nb <= 0 || ncols <= 0) return FM_OK;
    if (trsm_variant() == 0) return trsm_leaf_cols_launch<UPPER, false>(nb, ncols, T, ldt, st, B, ldb, sb, batch);
    if (!trsm_use_fma()) return trsm_leaf_dmma_launch<UPPER, false>(nb, ncols, T, ldt, st, B, ldb, sb, batch);
    // two columns per thread once 64-column CTAs already fill the machine (half the shared-memory traffic per column)
    const long long groups64 = (long long)ceil_div(ncols, 64) * batch;
    if (groups64 >= ctx().sm_count / 2) return trsm_leaf_launch<UPPER, 2>(nb, ncols, T, ldt, st, B, ldb, sb, batch);
    return trsm_leaf_launch<UPPER, 1>(nb, ncols, T, ldt, st, B, ldb, sb, batch);
}

int gemm_minus(int m, int n, int k, const double *a, int lda, long long sa, const double *b, int ldb, long long sb, double *c, int ldc,
               long long sc, int batch, int max_sms = 0, bool pdl_dependent = false) {
    if (m <= 0 || n <= 0 || k <= 0) return FM_OK;
    GemmEpilogue ep;
    ep.alpha = -1.0;
    ep.beta = 1.0;
    ep.max_sms = max_sms;
    ep.pdl_dependent = pdl_dependent;
    ep.dynamic_tiles = pdl_dependent;  // the trailing update that runs beside the panel kernel
    return k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch, ep);
}

// the single-launch wavefront solve: one CTA per (128-row block, 64-column chunk), all co-resident
template <bool UPPER>
int trsm_wave(int w, int ncols, const double *T, int ldt, double *B, int ldb, bool *done) {
    *done = false;
    static const bool off = getenv("FLOMAT_TRSM_NO_WAVE") != nullptr;  // developer switches
    static const bool wide_only = getenv("FLOMAT_TRSM_WAVE_WIDE") != nullptr;
    const int nblk = ceil_div(w, 128);
    if (off || nblk < 2 || ctx().max_smem_optin < TW_SMEM) return FM_OK;
    // 32-column CTAs (four owner warps + four helper warps) when that many CTAs still fit on the machine at once, 64-column CTAs
    // (eight unpaired warps) otherwise
    static const bool no_pair = getenv("FLOMAT_TRSM_WAVE_NO_PAIR") != nullptr;  // developer switch: four unpaired warps
    const bool narrow = !wide_only && (long long)nblk * ceil_div(ncols, 32) <= ctx().sm_count;
    const int cchunks = ceil_div(ncols, narrow ? 32 : 64);
    struct Flags { int *dev = nullptr; int epoch = 0; };
    static PerDevice<Flags> flags_pd;
    int *&d_flags = flags_pd().dev;
    int &epoch = flags_pd().epoch;
    const int nflags = 4096;
    if ((long long)nblk * cchunks > ctx().sm_count || nblk * cchunks > nflags) return FM_OK;
    if (!d_flags) {
        FM_CUDA(cudaMalloc(&d_flags, nflags * sizeof(int)));
        FM_CUDA(cudaMemsetAsync(d_flags, 0, nflags * sizeof(int), ctx().stream));
        epoch = 0;
    }
    static PerDevice<bool> attr_pd;  // function attributes are per device
    bool &attr = attr_pd();
    if (!attr) {
        FM_CUDA(cudaFuncSetAttribute(trsm_wave_kernel<UPPER, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
        FM_CUDA(cudaFuncSetAttribute(trsm_wave_kernel<UPPER, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
        FM_CUDA(cudaFuncSetAttribute(trsm_wave_kernel<UPPER, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
        attr = true;
    }
    if (++epoch == 0x7fffffff) {  // keep the published values distinct from anything an earlier launch left behind
        FM_CUDA(cudaMemsetAsync(d_flags, 0, nflags * sizeof(int), ctx().stream));
        epoch = 1;
    }
    const int vec_ok = ((reinterpret_cast<uintptr_t>(B) & 15u) == 0 && (ldb % 2) == 0) ? 1 : 0;
    const int tvec_ok = ((reinterpret_cast<uintptr_t>(T) & 15u) == 0 && (ldt % 2) == 0) ? 1 : 0;
    // cooperative launch: the runtime guarantees that all CTAs are resident at once (the flag waits depend on it) or refuses the
    // launch, in which case the caller falls back to the blocked sweep
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblk, cchunks, 1);
    cfg.blockDim = dim3(narrow && no_pair ? 128 : 256, 1, 1);
    cfg.dynamicSmemBytes = TW_SMEM;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const long long ldt_ll = ldt, ldb_ll = ldb;
    cudaError_t e;
    if (narrow && !no_pair) e = cudaLaunchKernelEx(&cfg, trsm_wave_kernel<UPPER, 4, true>, w, ncols, T, ldt_ll, B, ldb_ll, d_flags, epoch, vec_ok, tvec_ok);
    else if (narrow) e = cudaLaunchKernelEx(&cfg, trsm_wave_kernel<UPPER, 4, false>, w, ncols, T, ldt_ll, B, ldb_ll, d_flags, epoch, vec_ok, tvec_ok);
    else e = cudaLaunchKernelEx(&cfg, trsm_wave_kernel<UPPER, 8, false>, w, ncols, T, ldt_ll, B, ldb_ll, d_flags, epoch, vec_ok, tvec_ok);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();
        return FM_OK;  // not done: blocked sweep
    }
    if (e != cudaSuccess) return set_error(FM_ERR_CUDA, "trsm_wave: launch failed: %s", cudaGetErrorString(e));
    count_launch();
    *done = true;
    return FM_OK;
}

int split(int w) { return ((w / 2 + NB - 1) / NB) * NB; }  // first half of a recursive split, a multiple of the panel width

// B := T^-1 B with T the w x w triangle at `T` (unit lower or upper), B w x ncols; recursive: GEMMs with a large inner
// dimension carry the flops, the 128 x 128 leaves are substitution kernels
template <bool UPPER>
int rtrsm(int w, int ncols, const double *T, int ldt, long long st, double *B, int ldb, long long sb, int batch) {
    if (w <= 0 || ncols <= 0) return FM_OK;
    if (w <= NB) return trsm_leaf<UPPER>(w, ncols, T, ldt, st, B, ldb, sb, batch);
    if (batch == 1) {
        // one launch for the whole triangle whenever one CTA per (128-row block, 32- or 64-column chunk) fits on the machine at once:
        // few right-hand sides at any n (dgetrs with nrhs <= 256 up to n = 18944), or as many columns as rows for n <= 1024 (the solve
        // inside expm n = 512: 4 leaves + 3 GEMMs per triangle, ~110 us, become one ~50 us launch)
        bool done = false;
        FM_TRY(trsm_wave<UPPER>(w, ncols, T, ldt, B, ldb, &done));
        if (done) return FM_OK;
    }
    if (ncols <= 2 * NB) {
        // few right-hand sides: a recursive split would leave GEMMs with a handful of tiles and a long inner dimension;
        // the right-looking sweep (rank-128 updates of everything still to be solved) exposes rows/128 tiles per update
        const int nblk = ceil_div(w, NB);
        for (int s = 0; s < nblk; ++s) {
            const int k0 = (UPPER ? nblk - 1 - s : s) * NB, kb = std::min(NB, w - k0);
            FM_TRY(trsm_leaf<UPPER>(kb, ncols, T + (size_t)k0 * ldt + k0, ldt, st, B + k0, ldb, sb, batch));
            if (!UPPER) FM_TRY(gemm_minus(w - k0 - kb, ncols, kb, T + (size_t)k0 * ldt + k0 + kb, ldt, st, B + k0, ldb, sb, B + k0 + kb, ldb, sb, batch));
            else FM_TRY(gemm_minus(k0, ncols, kb, T + (size_t)k0 * ldt, ldt, st, B + k0, ldb, sb, B, ldb, sb, batch));
        }
        return FM_OK;
    }
    const int w1 = split(w), w2 = w - w1;
    const double *T11 = T, *T22 = T + (size_t)w1 * ldt + w1;
    double *B1 = B, *B2 = B + w1;
    if (!UPPER) {
        FM_TRY(rtrsm<false>(w1, ncols, T11, ldt, st, B1, ldb, sb, batch));
        FM_TRY(gemm_minus(w2, ncols, w1, T + w1, ldt, st, B1, ldb, sb, B2, ldb, sb, batch));  // B2 -= T21 B1
        return rtrsm<false>(w2, ncols, T22, ldt, st, B2, ldb, sb, batch);
    }
    FM_TRY(rtrsm<true>(w2, ncols, T22, ldt, st, B2, ldb, sb, batch));
    FM_TRY(gemm_minus(w1, ncols, w2, T + (size_t)w1 * ldt, ldt, st, B2, ldb, sb, B1, ldb, sb, batch));  // B1 -= T12 B2
    return rtrsm<true>(w1, ncols, T11, ldt, st, B1, ldb, sb, batch);
}

// factor columns [j0, j0+w) of the m-row matrix A (all updates from the columns left of j0 already applied).  On return
// those columns hold L\U with the interchanges of [j0, j0+w) applied to them and to nothing else.
int rgetrf(const Mat &A, int m, int j0, int w, int32_t *ipiv, long long sp, int32_t *info, const Maps &maps) {
    if (w <= NB) {
        FM_TRY(panel_factor(A, j0, m - j0, w, ipiv, sp, info, maps));
        return permute_rows(A, j0, w, maps, j0, j0 + w);
    }
    const int w1 = split(w), w2 = w - w1;
    const int j1 = j0 + w1;
    FM_TRY(rgetrf(A, m, j0, w1, ipiv, sp, info, maps));
    FM_TRY(permute_rows(A, j1, w2, maps, j0, j1));
    FM_TRY(rtrsm<false>(w1, w2, A.at(j0, j0), A.lda, A.stride, A.at(j0, j1), A.lda, A.stride, A.batch));
    FM_TRY(gemm_minus(m - j1, w2, w1, A.at(j1, j0), A.lda, A.stride, A.at(j0, j1), A.lda, A.stride, A.at(j1, j1), A.lda, A.stride, A.batch));
    FM_TRY(rgetrf(A, m, j1, w2, ipiv, sp, info, maps));
    return permute_rows(A, j0, w1, maps, j1, j0 + w);
}

// Right-looking blocked LU with one panel of look-ahead: as soon as panel k is applied to the columns of panel k+1, panel
// k+1 is factored by its cluster while the rest of the trailing matrix is updated on the remaining SMs.  The trailing GEMM
// is launched as a programmatic dependent of the panel kernel (the panel's CTAs are resident before the GEMM's first CTA
// is placed, so the two never fight over the 16 SMs of the cluster) and waits for the panel before it exits, which keeps
// plain stream order valid for everything queued afterwards.
// jbeg > 0 / jend >= 0 (multiples of NB; k_getrf_streamed): factor the panels [jbeg, jend) only -- the updates of the columns left of
// jbeg have been applied to [jbeg, n) already; the interchanges and updates of these panels go to the columns [jbeg, n) and to
// nothing else, and the panel at jend is left unfactored.
int getrf_lookahead(const Mat &A, int m, int n, int32_t *ipiv, long long sp, int32_t *info, const Maps &maps, int32_t *viol = nullptr,
                    int jbeg = 0, int jend = -1) {
    // viol != nullptr: no-interchange panels (k_getrf_nopiv for large single matrices): same schedule, and no row permutations at all
    const int mn = (jend >= 0 && jend < (m < n ? m : n)) ? jend : (m < n ? m : n);  // one past the last pivot factored here
    if (jbeg >= mn) return FM_OK;
    const bool overlap = A.batch == 1 && getenv("FLOMAT_LU_NO_LOOKAHEAD") == nullptr;
    int csize = 1;
    FM_TRY(panel_factor(A, jbeg, m - jbeg, mn - jbeg < NB ? mn - jbeg : NB, ipiv, sp, info, maps, &csize, viol));
    for (int j0 = jbeg; j0 < mn; j0 += NB) {
        const int jb = std::min(NB, mn - j0), j1 = j0 + jb;
        // panel j0 is complete: its interchanges go to its own and to the trailing columns now; the columns to its left (L)
        // are off the critical chain and get all their interchanges in one pass at the end.  (csize == 0: the tall fallback
        // has already interchanged the panel's own columns.)
        if (viol) { /* identity */ }
        else if (csize == 0) FM_TRY(permute_rows(A, j1, n - j1, maps, j0, j1));
        else FM_TRY(permute_rows(A, j0, n - j0, maps, j0, j1));
        const int nr = n - j1;
        if (nr <= 0) break;
        FM_TRY(trsm_leaf<false>(jb, nr, A.at(j0, j0), A.lda, A.stride, A.at(j0, j1), A.lda, A.stride, A.batch));
        const int mrest = m - j1;
        if (mrest <= 0) break;
        const int jb2 = j1 < mn ? std::min(NB, mn - j1) : 0;
        if (jb2 > 0) {  // next panel's columns first, then its factorization runs beside the rest of the update
            FM_TRY(gemm_minus(mrest, jb2, jb, A.at(j1, j0), A.lda, A.stride, A.at(j0, j1), A.lda, A.stride, A.at(j1, j1), A.lda, A.stride, A.batch));
            FM_TRY(panel_factor(A, j1, mrest, jb2, ipiv, sp, info, maps, &csize, viol));
        }
        const int nb_cols = nr - jb2;
        if (nb_cols > 0) {
            const bool dep = overlap && jb2 > 0;
            FM_TRY(gemm_minus(mrest, nb_cols, jb, A.at(j1, j0), A.lda, A.stride, A.at(j0, j1 + jb2), A.lda, A.stride, A.at(j1, j1 + jb2), A.lda,
                              A.stride, A.batch, (dep && csize > 0) ? ctx().sm_count - csize : 0, dep && csize > 0));
        }
    }
    if (viol) return FM_OK;
    return permute_rows(A, jbeg, mn - jbeg, maps, jbeg, mn, true);  // L part: column c receives the interchanges of every panel right of it
}

// ---------------------------------------------------------------------------------------------
// Cholesky (dpotrf_, flomat.rkt:1794; flomat-cholesky(!) :1815-1826), lower form A = L L^T, blocked right-looking on a
// scratch copy W of A:   potf2(W_kk) -> W21 := W21 L11^-T (leaf TRSM, right side) -> W22 -= W21 W21^T (the DMMA GEMM on the
// whole trailing square with W21^T formed by the transpose kernel: 2x the flops of a SYRK, no new kernel) -> ... ; finally only
// the requested triangle is written back, so the other triangle of A stays untouched as LAPACK promises.
// ---------------------------------------------------------------------------------------------
// unblocked right-looking Cholesky of one nb x nb block (nb <= 128), the block distributed over the registers of one CTA:
// thread (ti, tc) of a 16 x 16 grid holds the elements (ti + 16 a, tc + 16 b), a, b < 8.  Per column: the owners of the column
// scale it by 1/sqrt(pivot) and publish it in shared memory, then every thread applies the rank-1 update to its 64 elements.
// info = global 1-based index of the first non-positive pivot (DPOTF2 stops there).
__global__ void __launch_bounds__(256) potf2_kernel(int nb, double *a, long long lda, int32_t *info, int j0) {
    __shared__ double s_col[NB];
    __shared__ double s_d;
    const int tid = threadIdx.x, ti = tid & 15, tc = tid >> 4;
    double e[8][8];
#pragma unroll
    for (int ba = 0; ba < 8; ++ba)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const int i = ti + 16 * ba, c = tc + 16 * bb;
            e[ba][bb] = (i < nb && c < nb && i >= c) ? a[c * lda + i] : 0.0;
        }
    if (tid == 0) s_d = e[0][0];
    __syncthreads();
    bool stop = false;
    // the column loop is split into an unrolled loop over the 16-column blocks (so that every register index below is a
    // compile-time constant) and a rolled loop over the 16 columns of a block
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
#pragma unroll 1
        for (int jt = 0; jt < 16; ++jt) {
            const int j = 16 * jb + jt;
            if (stop || j >= nb) break;
            const double d = s_d;
            if (!(d > 0.0)) {  // not positive definite (or NaN)
                if (tid == 0 && *info == 0) *info = j0 + j + 1;
                stop = true;
                break;
            }
            const double rd = sqrt(d);
            if (tc == jt) {  // this thread owns rows ti + 16 a of column j
#pragma unroll
                for (int ba = 0; ba < 8; ++ba) {
                    const int i = ti + 16 * ba;
                    if (i == j) e[ba][jb] = rd;
                    else if (i > j) e[ba][jb] = e[ba][jb] / rd;
                    if (i >= j) s_col[i] = e[ba][jb];
                }
            }
            __syncthreads();
            double li[8], lc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                li[q] = s_col[ti + 16 * q];
                lc[q] = s_col[tc + 16 * q];
            }
#pragma unroll
            for (int ba = 0; ba < 8; ++ba)
#pragma unroll
                for (int bb = 0; bb < 8; ++bb) {
                    if (bb < jb || ba < bb) continue;  // finished columns / blocks above the diagonal (compile time)
                    const int i = ti + 16 * ba, c = tc + 16 * bb;
                    if (c > j && i >= c && i < nb) e[ba][bb] -= li[ba] * lc[bb];
                }
            // the next pivot, already updated, goes to everybody
            if (jt < 15) {
                if (ti == jt + 1 && tc == jt + 1) s_d = e[jb][jb];
            } else if (jb < 7) {
                if (ti == 0 && tc == 0) s_d = e[jb + 1 < 8 ? jb + 1 : 7][jb + 1 < 8 ? jb + 1 : 7];
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int ba = 0; ba < 8; ++ba)
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
            const int i = ti + 16 * ba, c = tc + 16 * bb;
            if (i < nb && c < nb && i >= c) a[c * lda + i] = e[ba][bb];
        }
}
// dst triangle := src triangle (lower: i >= j); with `flip` dst(j, i) := src(i, j) (the upper factor U = L^T)
__global__ void __launch_bounds__(256) tri_copy_kernel(int n, const double *__restrict__ src, long long lds, double *dst, long long ldd, int flip) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int nt = (n + 31) / 32;
    for (long long t = blockIdx.x; t < (long long)nt * nt; t += gridDim.x) {
        const int ti = (int)(t % nt), tj = (int)(t / nt);
        if (ti < tj) continue;  // tiles strictly above the diagonal hold nothing of the lower triangle (uniform per CTA)
        const int i0 = ti * 32, j0 = tj * 32;
        if (!flip) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = i0 + tx, j = j0 + ty + 8 * r;
                if (i < n && j < n && i >= j) dst[j * ldd + i] = src[j * lds + i];
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = i0 + tx, j = j0 + ty + 8 * r;
                tile[ty + 8 * r][tx] = (i < n && j < n && i >= j) ? src[j * lds + i] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int j = j0 + tx, i = i0 + ty + 8 * r;  // dst(j, i) = src(i, j): row j, column i, j <= i
                if (i < n && j < n && i >= j) dst[i * ldd + j] = tile[tx][ty + 8 * r];
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LU WITHOUT interchanges, verified: the fast path of the solve inside expm (expm.rkt:199, `(mldivide den num)`).
// The Pade denominator den = sum_k c_k (-X)^k with ||X||_1 in (1/4, 1/2] (the reference's own scaling, expm.rkt:212-215) satisfies
// ||den - I||_1 <= 0.282: it is strictly diagonally dominant by columns, and for such matrices partial pivoting never interchanges
// (the Schur complements stay column diagonally dominant), i.e. dgetrf's answer IS the no-pivot factorization with ipiv = 1..n.
// Without the pivot search the panel is plain arithmetic -- no argmax chain of ~2300 cycles per column.  The claim is CHECKED, not
// assumed: every multiplier must satisfy |l_ij| <= 0.999 (|l_ij| <= 1 for all i > j is exactly "IDAMAX picks the diagonal at step j")
// and every pivot must be finite and nonzero; otherwise a violation flag is raised and the caller repeats the factorization with the
// pivoting kernels (expm.cu).  Blocked right-looking, NB = 128:
//   getf2_nopiv_kernel  the 128 x 128 diagonal block in the shared memory of one CTA, eliminated in block steps of 16 columns
//                       (DGETF2's arithmetic: multiplier = entry times the pivot's reciprocal); also writes U11^T (lower, non-unit)
//                       to a scratch block for the right-side solve below
//   L21 = A21 U11^-1    the leaf TRSM in its right-side mode on T = U11^T (as Cholesky uses it)
//   U12 = L11^-1 A12    the leaf TRSM (unit lower);   A22 -= L21 U12   the DMMA GEMM
// ---------------------------------------------------------------------------------------------
constexpr int NP_LD = NB + 2;                                                      // column stride of the block in shared memory
constexpr int NP_SMEM = (NB * NP_LD + 16 * NB) * (int)sizeof(double);              // the block + a row-major copy of the current U block row
__global__ void __launch_bounds__(256, 1) getf2_nopiv_kernel(int nb, double *a, long long lda, long long sa, double *ut, long long sut, int32_t *viol) {
    // The 128 x 128 block lives in shared memory: S(i, c) = S[c * NP_LD + i], padded with the identity beyond nb.  Eight block steps of
    // 16 columns: (1) unblocked elimination of the 16-wide block column over all rows below (thread = row, one barrier per column: the
    // sequential part; measured ~330 ns per column); (2) the U block row right of it by forward
    // substitution with the unit-lower 16 x 16 block (thread = column); (3) rank-16 update of the trailing square from registers
    // (thread (ty, tx) of a 16 x 16 grid owns the elements (ty + 16 a, tx + 16 b): 14 shared-memory loads per 49 FMAs).
    extern __shared__ double sm[];
    double *S = sm;
    double *Us = sm + NB * NP_LD;  // Us[t * NB + c] = U(c0 + t, c)
    a += blockIdx.x * sa;
    ut += blockIdx.x * sut;
    const int tid = threadIdx.x;
    {
        const int i = tid & (NB - 1), ch = tid >> 7;  // two columns per pass; 8 passes of loads in flight at a time
#pragma unroll 1
        for (int cb = 0; cb < NB; cb += 16) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = cb + 2 * q + ch;
                v[q] = (i < nb && c < nb) ? a[c * lda + i] : (i == c ? 1.0 : 0.0);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) S[(cb + 2 * q + ch) * NP_LD + i] = v[q];
        }
    }
    __syncthreads();
    bool bad = false;
    const int ty = tid & 15, tx = tid >> 4;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
        const int c0 = 16 * kb;
        if (c0 >= nb) break;  // (uniform)
        // ---- (1) the block column, unblocked: thread = row, one barrier per column ----
#pragma unroll 1
        for (int jt = 0; jt < 16; ++jt) {
            const int j = c0 + jt;
            if (j >= nb) break;  // (uniform)
            const double d = S[j * NP_LD + j];
            if (!(fabs(d) > 0.0) || !(fabs(d) < DBL_MAX)) bad = true;  // a zero / non-finite pivot is not a dominant one
            const bool use_rcp = fabs(d) >= DBL_MIN;                  // DGETF2: scale by the reciprocal unless the pivot is tiny
            const double rcp = 1.0 / d;
            if (tid < NB && tid > j) {
                const int row = tid;
                // all loads first (the compiler cannot reorder them across the stores into the same array), then the arithmetic
                // (only the columns still inside the block column: the step is bound by the shared-memory pipe -- 250 ns of its 330 ns
                //  per column were these loads and stores when all 15 were issued regardless)
                double own[15], piv[15];
                const int nq = 16 - jt;  // columns j+1 .. c0+15 are q = 1 .. nq-1
#pragma unroll
                for (int q = 1; q < 16; ++q) {
                    own[q - 1] = 0.0;
                    piv[q - 1] = 0.0;
                    if (q < nq) {
                        own[q - 1] = S[(j + q) * NP_LD + row];
                        piv[q - 1] = S[(j + q) * NP_LD + j];
                    }
                }
                const double x = S[j * NP_LD + row];
                const double l = use_rcp ? x * rcp : x / d;
                if (!(fabs(l) <= 0.999)) bad = true;  // |l| <= 1 everywhere is "IDAMAX picks the diagonal" (NaN fails the test too)
#pragma unroll
                for (int q = 1; q < 16; ++q)
                    if (q < nq) S[(j + q) * NP_LD + row] = own[q - 1] - l * piv[q - 1];
                S[j * NP_LD + row] = l;
            }
            __syncthreads();
        }
        const int r0 = c0 + 16;  // first row / column of the trailing square
        if (r0 >= NB) break;
        // ---- (2) U(c0 .. c0+15, c) for the columns right of the block: L11 u = a, unit lower ----
        if (tid < NB - r0) {
            const int c = r0 + tid;
            double u[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                double acc = S[c * NP_LD + c0 + t];
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q < t) acc -= S[(c0 + q) * NP_LD + c0 + t] * u[q];
                u[t] = acc;
            }
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                S[c * NP_LD + c0 + t] = u[t];
                Us[t * NB + c] = u[t];
            }
        }
        __syncthreads();
        // ---- (3) trailing square: S(i, c) -= sum_t L(i, c0 + t) U(c0 + t, c) ----
        constexpr int NA_MAX = 7;
        const int na = 7 - kb;  // 16-wide strips of the trailing square (compile time after unrolling)
        double acc[NA_MAX][NA_MAX];
#pragma unroll
        for (int x = 0; x < NA_MAX; ++x)
#pragma unroll
            for (int y = 0; y < NA_MAX; ++y)
                if (x < na && y < na) acc[x][y] = S[(r0 + tx + 16 * y) * NP_LD + r0 + ty + 16 * x];
#pragma unroll 4
        for (int t = 0; t < 16; ++t) {
            double lf[NA_MAX], uf[NA_MAX];
#pragma unroll
            for (int x = 0; x < NA_MAX; ++x)
                if (x < na) {
                    lf[x] = S[(c0 + t) * NP_LD + r0 + ty + 16 * x];
                    uf[x] = Us[t * NB + r0 + tx + 16 * x];
                }
#pragma unroll
            for (int x = 0; x < NA_MAX; ++x)
#pragma unroll
                for (int y = 0; y < NA_MAX; ++y)
                    if (x < na && y < na) acc[x][y] -= lf[x] * uf[y];
        }
#pragma unroll
        for (int x = 0; x < NA_MAX; ++x)
#pragma unroll
            for (int y = 0; y < NA_MAX; ++y)
                if (x < na && y < na) S[(r0 + tx + 16 * y) * NP_LD + r0 + ty + 16 * x] = acc[x][y];
        __syncthreads();
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx & (NB - 1), c = idx >> 7;
        if (i < nb && c < nb) a[c * lda + i] = S[c * NP_LD + i];
    }
    // T = U11^T (lower, non-unit) for the right-side solve: T(r, k) = U(k, r), column-major with leading dimension NB; consecutive
    // threads walk a column of T, i.e. a ROW of U
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx & (NB - 1), k = idx >> 7;
        if (r < nb && k <= r) ut[(long long)k * NB + r] = S[r * NP_LD + k];
    }
    if (__syncthreads_or(bad) && tid == 0) atomicOr(viol, 1);
}

// ipiv := 1..n (no interchanges) for every item; info := 0
__global__ void iota_pivots_kernel(int n, int32_t *ipiv, long long sp, int32_t *info) {
    int32_t *p = ipiv + blockIdx.y * sp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i + 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) info[blockIdx.y] = 0;
}

// idx[i] = original row that LAPACK's interchange sequence leaves at position i (sequential: one thread)
__global__ void perm_forward_kernel(int n, const int32_t *ipiv, int32_t *idx) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int i = 0; i < n; ++i) idx[i] = i;
    for (int i = 0; i < n; ++i) {
        const int p = ipiv[i] - 1;
        if (p != i && p >= 0 && p < n) { const int t = idx[i]; idx[i] = idx[p]; idx[p] = t; }
    }
}
// same, with ipiv and idx staged in shared memory (8n bytes): the sequential replay runs at shared-memory instead of global-memory
// latency (n = 8192: 0.78 ms -> < 0.1 ms)
__global__ void __launch_bounds__(1024) perm_forward_smem_kernel(int n, const int32_t *ipiv, int32_t *idx) {
    extern __shared__ int32_t pf[];
    int32_t *sp = pf, *si = pf + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sp[i] = ipiv[i] - 1; si[i] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) {
            const int p = sp[i];
            if (p != i && p >= 0 && p < n) { const int t = si[i]; si[i] = si[p]; si[p] = t; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) idx[i] = si[i];
}
// dst(:, idx[j]) := src(:, j)
__global__ void __launch_bounds__(256) scatter_cols_kernel(int n, const double *__restrict__ src, long long lds, double *dst, long long ldd,
                                                            const int32_t *idx) {
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        const long long dj = idx[j];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[dj * ldd + i] = src[(long long)j * lds + i];
    }
}

// Z := L^-1 for the unit lower triangle T (w x w), Z initialised to the identity: Z11 = L11^-1, Z22 = L22^-1 recursively,
// Z21 = -L22^-1 (L21 Z11); Z12 stays zero.  Half the work of a forward sweep over a dense identity.
// Z := inverse of the unit lower triangle at T (Z holds the identity on entry), bottom-up and level by level so that the many
// small independent pieces of a level are ONE batched launch each: all 128 x 128 diagonal blocks first, then pairs of inverted
// w x w blocks are merged ([Z11 0; Z21 Z22] with Z21 = -L22^-1 (L21 Z11)) for w = 128, 256, ... with batch = number of pairs and
// item stride 2w (ld + 1).  A ragged last pair (n not a multiple of 2w) is one extra unbatched merge per level.
// (The recursive top-down form ran the same pieces one after the other: 230 small GEMMs and 290 leaves for n = 8192.)
int rinv_merge(int w1, int w2, const double *T, int ldt, double *Z, int ldz, long long stT, long long stZ, int batch) {
    GemmEpilogue ep;
    ep.alpha = -1.0;
    ep.beta = 0.0;
    FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, w2, w1, w1, T + w1, ldt, stT, Z, ldz, stZ, Z + w1, ldz, stZ, batch, ep));  // Z21 = -L21 Z11
    return rtrsm<false>(w2, w1, T + (size_t)w1 * ldt + w1, ldt, stT, Z + w1, ldz, stZ, batch);                        // Z21 := L22^-1 Z21
}
int rinv_unit_lower(int n, const double *T, int ldt, double *Z, int ldz) {
    if (n <= 0) return FM_OK;
    const int nfull0 = n / NB, rem0 = n % NB;
    if (nfull0 > 0) FM_TRY(trsm_leaf<false>(NB, NB, T, ldt, (long long)NB * (ldt + 1), Z, ldz, (long long)NB * (ldz + 1), nfull0));
    if (rem0 > 0)
        FM_TRY(trsm_leaf<false>(rem0, rem0, T + (size_t)nfull0 * NB * (ldt + 1), ldt, 0, Z + (size_t)nfull0 * NB * (ldz + 1), ldz, 0, 1));
    for (long long w = NB; w < n; w *= 2) {
        const int nblocks = (int)((n + w - 1) / w), npairs = nblocks / 2;  // pairs of level-w blocks whose second block exists
        if (npairs == 0) break;
        const long long last_second = std::min<long long>(w, n - (2LL * npairs - 1) * w);
        const int nfull = last_second == w ? npairs : npairs - 1;
        const long long stT = 2 * w * ((long long)ldt + 1), stZ = 2 * w * ((long long)ldz + 1);
        if (nfull > 0) FM_TRY(rinv_merge((int)w, (int)w, T, ldt, Z, ldz, stT, stZ, nfull));
        if (nfull < npairs) FM_TRY(rinv_merge((int)w, (int)last_second, T + (size_t)nfull * stT, ldt, Z + (size_t)nfull * stZ, ldz, 0, 0, 1));
    }
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
    const int npanels = ceil_div(mn, NB);
    Maps maps;
    maps.item_stride = (long long)npanels * MAPN;
    maps.base = g_maps_getrf().get((size_t)maps.item_stride * batch);
    if (!maps.base) return FM_ERR_NOMEM;
    static const bool recursive = getenv("FLOMAT_LU_RECURSIVE") != nullptr;  // developer switch: DGETRF2-style recursion
    if (!recursive) return getrf_lookahead(A, m, n, ipiv, sp, info, maps);
    FM_TRY(rgetrf(A, m, 0, mn, ipiv, sp, info, maps));
    if (n > mn) {  // wide matrix: U12 = L11^-1 (P A12)
        FM_TRY(permute_rows(A, mn, n - mn, maps, 0, mn));
        FM_TRY(rtrsm<false>(mn, n - mn, a, lda, sa, A.at(0, mn), lda, sa, batch));
    }
    return FM_OK;
}

// LU of a square matrix whose columns ARRIVE in chunks (the host-operand dgesv_ / dgetrf_: the upload of a chunk runs beside the
// arithmetic of the chunks before it) and whose finished parts LEAVE early.  Chunks are [bounds[c], bounds[c+1]) (multiples of 128
// except the last bound, n).  Step c works on a window of two chunks, c (to be factored) and c+1 (just arrived):
//     arrive(n0, n1)           -- the caller makes the columns of chunk c+1 valid on the compute stream
//     catch-up of chunk c+1    -- against the chunks LEFT of c, in one piece: the interchanges of their panels, U(0:c0, chunk c+1) =
//                                 L(0:c0, 0:c0)^-1 A(0:c0, chunk c+1), A(c0:, chunk c+1) -= L(c0:, 0:c0) U
//     upper_done(n0, n1, c0)   -- rows [0, c0) of columns [n0, n1) are FINAL (the caller may start to download them)
//     getrf_lookahead          -- the panels of chunk c, right-looking over the window [c0, n1): the update of chunk c+1 by chunk c's
//                                 panels sits in the look-ahead slots, beside the next panel's factorisation
//     the interchanges of chunk c to the columns left of c0
//     done(c0, c1, lim)        -- rows [c0, c1) of columns [0, lim = n1) are FINAL (later interchanges move rows >= c1 only); after the
//                                 last chunk: rows [c0, n)
// The same arithmetic as k_getrf up to the summation order of the Schur updates (one k = c0 product per chunk instead of c0/128
// rank-128 updates), so a pivot can differ from k_getrf's only where two candidates agree to rounding.  What it costs: the panel
// chain is hidden behind the update of a two-chunk window instead of the whole trailing matrix, and the catch-ups are not hidden at
// all (device-resident n = 8192: tools/chunked_lu_time.py) -- less than the transfers it hides.
int k_getrf_streamed(int n, double *a, int lda, int32_t *ipiv, int32_t *info, const std::vector<int> &bounds,
                     const std::function<int(int, int)> &arrive, const std::function<int(int, int, int)> &upper_done,
                     const std::function<int(int, int, int)> &done) {
    if (n < 0 || lda < (n > 1 ? n : 1) || bounds.size() < 2 || bounds.front() != 0 || bounds.back() != n)
        return set_error(FM_ERR_ARG, "getrf_streamed: bad dimensions");
    for (size_t c = 0; c + 1 < bounds.size(); ++c)
        if (bounds[c + 1] <= bounds[c] || bounds[c] % NB != 0) return set_error(FM_ERR_ARG, "getrf_streamed: chunk bounds must increase in multiples of %d", NB);
    zero_info_kernel<<<1, 256, 0, ctx().stream>>>(info, 1);
    FM_LAUNCH_CHECK();
    Mat A{a, lda, 0, 1};
    Maps maps;
    maps.item_stride = (long long)ceil_div(n, NB) * MAPN;
    maps.base = g_maps_getrf().get((size_t)maps.item_stride);
    if (!maps.base) return FM_ERR_NOMEM;
    const size_t nchunks = bounds.size() - 1;
    FM_TRY(arrive(bounds[0], bounds[1]));
    for (size_t c = 0; c < nchunks; ++c) {
        const int c0 = bounds[c], c1 = bounds[c + 1];
        int lim = c1;
        if (c + 1 < nchunks) {
            const int n0 = c1, n1 = bounds[c + 2], w = n1 - n0;
            lim = n1;
            FM_TRY(arrive(n0, n1));
            if (c0 > 0) {
                FM_TRY(permute_rows(A, n0, w, maps, 0, c0));
                FM_TRY(rtrsm<false>(c0, w, a, lda, 0, A.at(0, n0), lda, 0, 1));
                FM_TRY(upper_done(n0, n1, c0));
                FM_TRY(gemm_minus(n - c0, w, c0, A.at(c0, 0), lda, 0, A.at(0, n0), lda, 0, A.at(c0, n0), lda, 0, 1));
            }
        }
        FM_TRY(getrf_lookahead(A, n, lim, ipiv, 0, info, maps, nullptr, c0, c1));
        if (c0 > 0) FM_TRY(permute_rows(A, 0, c0, maps, c0, c1));
        FM_TRY(done(c0, c1, lim));
    }
    return FM_OK;
}

// a := L\\U of a without interchanges (square, batched), ipiv := 1..n, info := 0; *viol (device int32, zeroed by the caller) is raised
// when the factorization is NOT what partial pivoting would have produced (see getf2_nopiv_kernel) -- the caller must then discard it.
int k_getrf_nopiv(int n, double *a, int lda, long long sa, int32_t *ipiv, long long sp, int32_t *info, int32_t *viol, int batch) {
    if (n < 0 || lda < (n > 1 ? n : 1) || batch < 0) return set_error(FM_ERR_ARG, "getrf_nopiv: bad dimensions");
    if (batch == 0) return FM_OK;
    if (batch > 65535) return set_error(FM_ERR_ARG, "getrf_nopiv: at most 65535 items per call");
    iota_pivots_kernel<<<dim3(n > 0 ? std::min(ceil_div(n, 256), 64) : 1, batch), 256, 0, ctx().stream>>>(n, ipiv, sp, info);
    FM_LAUNCH_CHECK();
    if (n == 0) return FM_OK;
    // Large single matrices: the look-ahead schedule of the pivoting LU (panel k+1 beside the trailing update of panel k) with the
    // no-interchange variant of the cluster panel kernel.  Small and batched ones: one CTA per diagonal block, below.
    // (measured, ms: n = 2048 2.24 cluster / 1.85 diagonal block; 4096 4.99 / 4.96; 8192 17.2 / 19.0; the pivoting LU: 3.15, 7.1, 21.5)
    static const int cluster_min_n = getenv("FLOMAT_NOPIV_CLUSTER_MIN_N") ? atoi(getenv("FLOMAT_NOPIV_CLUSTER_MIN_N")) : 6144;
    if (batch == 1 && n >= cluster_min_n && n <= 16384) {
        Mat A{a, lda, sa, 1};
        const int npanels = ceil_div(n, NB);
        Maps maps;
        maps.item_stride = (long long)npanels * MAPN;
        maps.base = g_maps_getrf().get((size_t)maps.item_stride);
        if (!maps.base) return FM_ERR_NOMEM;
        return getrf_lookahead(A, n, n, ipiv, sp, info, maps, viol);
    }
    static PerDevice<bool> attr_pd;  // function attributes are per device
    if (!attr_pd()) {
        FM_CUDA(cudaFuncSetAttribute(getf2_nopiv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NP_SMEM));
        attr_pd() = true;
    }
    double *ut = static_cast<double *>(fm_alloc_bytes((size_t)NB * NB * batch * sizeof(double)));  // U11^T of the current step, per item
    if (!ut) return FM_ERR_NOMEM;
    int rc = FM_OK;
    for (int j0 = 0; j0 < n && rc == FM_OK; j0 += NB) {
        const int jb = std::min(NB, n - j0), j1 = j0 + jb, rest = n - j1;
        double *akk = a + (size_t)j0 * lda + j0;
        getf2_nopiv_kernel<<<batch, 256, NP_SMEM, ctx().stream>>>(jb, akk, lda, sa, ut, (long long)NB * NB, viol);
        rc = cudaGetLastError() == cudaSuccess ? FM_OK : set_error(FM_ERR_CUDA, "getf2_nopiv launch failed");
        count_launch();
        if (rest <= 0 || rc != FM_OK) break;
        // L21 = A21 U11^-1; the kernel checks the multipliers it produces (|l| <= 0.999) and raises *viol otherwise
        rc = trsm_leaf_cols_launch<false, true>(jb, rest, ut, NB, (long long)NB * NB, akk + jb, lda, sa, batch, viol);
        if (rc == FM_OK) rc = trsm_leaf<false>(jb, rest, akk, lda, sa, a + (size_t)j1 * lda + j0, lda, sa, batch);  // U12 = L11^-1 A12
        if (rc == FM_OK)
            rc = gemm_minus(rest, rest, jb, akk + jb, lda, sa, a + (size_t)j1 * lda + j0, lda, sa, a + (size_t)j1 * lda + j1, lda, sa, batch);
    }
    fm_free(ut);
    return rc;
}

int k_getrs(int n, int nrhs, const double *lu, int lda, long long sa, const int32_t *ipiv, long long sp, double *b, int ldb, long long sb,
            int batch) {
    if (n < 0 || nrhs < 0 || lda < (n > 1 ? n : 1) || ldb < (n > 1 ? n : 1) || batch < 0) return set_error(FM_ERR_ARG, "getrs: bad dimensions");
    if (n == 0 || nrhs == 0 || batch == 0) return FM_OK;
    if (ipiv == nullptr) {  // the factorization had no interchanges (k_getrf_nopiv): P = I, just the two triangular sweeps
        FM_TRY(rtrsm<false>(n, nrhs, lu, lda, sa, b, ldb, sb, batch));
        return rtrsm<true>(n, nrhs, lu, lda, sa, b, ldb, sb, batch);
    }
    Mat B{b, ldb, sb, batch};
    const int nblocks = ceil_div(n, NB);
    Maps maps;
    maps.item_stride = (long long)nblocks * MAPN;
    maps.base = g_maps_getrs().get((size_t)maps.item_stride * batch);
    if (!maps.base) return FM_ERR_NOMEM;
    ipiv_to_maps_kernel<<<dim3(nblocks, batch), 32, 0, ctx().stream>>>(n, ipiv, sp, maps.base, maps.item_stride, 0, 0);
    FM_LAUNCH_CHECK();
    FM_TRY(permute_rows(B, 0, nrhs, maps, 0, n));                          // P B   (DLASWP)
    FM_TRY(rtrsm<false>(n, nrhs, lu, lda, sa, b, ldb, sb, batch));          // L y = P B
    return rtrsm<true>(n, nrhs, lu, lda, sa, b, ldb, sb, batch);            // U x = y
}

// dgetri: a (its LU factors) := inverse.  A = P^T L U  =>  inv(A) = U^-1 L^-1 P: W = U^-1 (L^-1) with the structured inverse of
// L (lower triangular, n^3/3 instead of n^3 flops for that half), then the columns of W go to their places: 4/3 n^3 flops on
// top of the factorization, LAPACK's count, all of it in GEMMs and leaf solves.
int k_getri(int n, double *a, int lda, const int32_t *ipiv) {
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "getri: bad dimensions");
    if (n == 0) return FM_OK;
    const int ldx = n + (n & 1);
    double *x = static_cast<double *>(fm_alloc_bytes((size_t)ldx * n * sizeof(double) + (size_t)n * sizeof(int32_t) + 16));
    if (!x) return FM_ERR_NOMEM;
    int32_t *idx = reinterpret_cast<int32_t *>(x + (size_t)ldx * n);
    int rc = k_eye(x, ldx, n, n, 0);
    if (rc == FM_OK) rc = rinv_unit_lower(n, a, lda, x, ldx);
    if (rc == FM_OK) rc = rtrsm<true>(n, n, a, lda, 0, x, ldx, 0, 1);
    if (rc == FM_OK) {
        if ((size_t)n * 8 <= (size_t)ctx().max_smem_optin) {
            static PerDevice<bool> attr_pd;
            bool &attr = attr_pd();
            if (!attr) {
                cudaFuncSetAttribute(perm_forward_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx().max_smem_optin);
                attr = true;
            }
            perm_forward_smem_kernel<<<1, 1024, (size_t)n * 8, ctx().stream>>>(n, ipiv, idx);
        } else {
            perm_forward_kernel<<<1, 32, 0, ctx().stream>>>(n, ipiv, idx);
        }
        dim3 grid(ceil_div(n, 256) < 32 ? ceil_div(n, 256) : 32, n < 4096 ? n : 4096);
        scatter_cols_kernel<<<grid, 256, 0, ctx().stream>>>(n, x, ldx, a, lda, idx);
        rc = cudaGetLastError() == cudaSuccess ? FM_OK : set_error(FM_ERR_CUDA, "getri: launch failed");
        count_launch(2);
    }
    fm_free(x);
    return rc;
}

// dpotrf: uplo 'L' (A = L L^T, lower triangle overwritten) or 'U' (A = U^T U, upper triangle overwritten)
int k_potrf(char uplo, int n, double *a, int lda, int32_t *info) {
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "potrf: bad dimensions");
    const bool upper = (uplo == 'U' || uplo == 'u');
    if (!upper && uplo != 'L' && uplo != 'l') return set_error(FM_ERR_ARG, "potrf: uplo must be 'U' or 'L'");
    zero_info_kernel<<<1, 32, 0, ctx().stream>>>(info, 1);
    FM_LAUNCH_CHECK();
    if (n == 0) return FM_OK;
    const int ldw = n + (n & 1);
    double *w = static_cast<double *>(fm_alloc_bytes(((size_t)ldw * n + (size_t)ldw * NB) * sizeof(double)));
    if (!w) return FM_ERR_NOMEM;
    double *wt = w + (size_t)ldw * n;  // W21^T of the current step: NB x rows, leading dimension NB
    int rc = upper ? k_transpose(n, n, a, lda, w, ldw) : k_copy2d(w, ldw, a, lda, n, n);  // the lower triangle of W holds the data
    for (int j0 = 0; rc == FM_OK && j0 < n; j0 += NB) {
        const int jb = std::min(NB, n - j0), j1 = j0 + jb, rest = n - j1;
        double *wkk = w + (size_t)j0 * ldw + j0;
        potf2_kernel<<<1, 256, 0, ctx().stream>>>(jb, wkk, ldw, info, j0);
        rc = cudaGetLastError() == cudaSuccess ? FM_OK : set_error(FM_ERR_CUDA, "potf2 launch failed");
        count_launch();
        if (rest <= 0 || rc != FM_OK) break;
        double *w21 = wkk + jb;  // rest x jb
        if (trsm_variant() == 0) rc = trsm_leaf_cols_launch<false, true>(jb, rest, wkk, ldw, 0, w21, ldw, 0, 1);
        else if (!trsm_use_fma()) rc = trsm_leaf_dmma_launch<false, true>(jb, rest, wkk, ldw, 0, w21, ldw, 0, 1);
        else if (rest >= 64 * (ctx().sm_count / 2)) rc = trsm_leaf_launch<false, 2, true>(jb, rest, wkk, ldw, 0, w21, ldw, 0, 1);
        else rc = trsm_leaf_launch<false, 1, true>(jb, rest, wkk, ldw, 0, w21, ldw, 0, 1);
        if (rc == FM_OK) rc = k_transpose(rest, jb, w21, ldw, wt, NB);
        // A22 -= L21 L21^T on the lower triangle only: column strips of 1024, each updated from its diagonal down (the strictly upper
        // part of W is never read: 56 % of the flops of the full square at n = 8192, no SYRK variant of the GEMM kernel needed); n = 8192: 23.7 -> 22.5 ms
        const int STRIP = rest > 2048 ? 8 * NB : rest;  // small trailing squares: one launch (the extra launches cost more than the flops)
        for (int c0 = 0; rc == FM_OK && c0 < rest; c0 += STRIP) {
            const int cw = std::min(STRIP, rest - c0);
            rc = gemm_minus(rest - c0, cw, jb, w21 + c0, ldw, 0, wt + (size_t)c0 * NB, NB, 0, w + (size_t)(j1 + c0) * ldw + j1 + c0, ldw, 0, 1);
        }
    }
    if (rc == FM_OK) {
        const long long tiles = (long long)ceil_div(n, 32) * ceil_div(n, 32);
        const long long cap = (long long)ctx().sm_count * 16;
        tri_copy_kernel<<<(int)(tiles < cap ? tiles : cap), 256, 0, ctx().stream>>>(n, w, ldw, a, lda, upper ? 1 : 0);
        rc = cudaGetLastError() == cudaSuccess ? FM_OK : set_error(FM_ERR_CUDA, "potrf: triangle copy launch failed");
        count_launch();
    }
    fm_free(w);
    return rc;
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

// LAPACK's singularity test of DGETRI / DTRTRI: the 1-based index of the first exactly-zero U(i,i), 0 if there is none.
// (Not inferred from the determinant: an overflowed partial product times the zero pivot is NaN, not 0.)
__global__ void __launch_bounds__(256) first_zero_diag_kernel(int n, const double *lu, long long lda, int32_t *out) {
    __shared__ int s_min;
    if (threadIdx.x == 0) s_min = 0x7fffffff;
    __syncthreads();
    int mine = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += 256)
        if (lu[i * (lda + 1)] == 0.0) { mine = i + 1; break; }  // indices grow with i: the first hit is this thread's minimum
    if (mine != 0x7fffffff) atomicMin(&s_min, mine);
    __syncthreads();
    if (threadIdx.x == 0) *out = s_min == 0x7fffffff ? 0 : s_min;
}

int k_first_zero_diag(int n, const double *lu, int lda, int32_t *d_out) {
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "getri: bad dimensions");
    first_zero_diag_kernel<<<1, 256, 0, ctx().stream>>>(n, lu, lda, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_det_device(int n, const double *lu, int lda, const int32_t *ipiv, double *d_out) {
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "det: bad dimensions");
    det_kernel<<<1, 256, 0, ctx().stream>>>(n, lu, lda, ipiv, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

}  // namespace fm
