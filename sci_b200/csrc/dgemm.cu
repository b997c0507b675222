// dgemm.cu -- C := alpha*op(A)*op(B) + beta*C (+ diag*I), FP64, column-major: the kernel behind `times`
// (flomat.rkt:3202 -> flomat* :905 -> cblas_dgemm :888), every product inside expm (expm.rkt:167-210) and
// the trailing updates of the blocked LU.
//
// B200 design (sm_100a).  tcgen05/UMMA has no f64 kind, so FP64 tensor math is the warp-level DMMA
// (`mma.sync.aligned.m8n8k4.f64` -> SASS DMMA.8x8x4), whose measured issue peak on B200 is 37.0 TFLOP/s
// (tools/peak_fp64.cu) -- that is the roofline denominator of this kernel.
//
//   * persistent kernel, one CTA per SM, static round-robin over 128x128 output tiles, tiles rasterised in
//     bands of 16 tile-rows so that the 148 concurrently running tiles share A/B panels through the 126 MB L2;
//   * warp-specialised: warp 8 is the TMA producer (one elected lane), warps 0-7 are DMMA consumers with
//     64x32 warp tiles (64 accumulator doubles per thread, all in registers);
//   * operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a ring of STAGES smem buffers
//     guarded by full/empty mbarriers; A (MN-major in memory) arrives as 16-row boxes, B (K-major) as one box;
//   * fragment loads are conflict-free LDS.128: the m8n8k4 row/column/k indices are permuted (see a_offsets /
//     b_offsets) so every thread reads 16-byte pairs and the 128B swizzle spreads a quarter-warp over all banks;
//   * epilogue applies alpha/beta/diag from registers and writes 16-byte pairs straight to C (and, for the
//     column-sharded multi-GPU product, to every peer GPU's C over NVLink -- the all-gather is fused here).
//
// Operands that TMA cannot describe (odd lda, 8-byte-aligned views, transposed A/B) take dgemm_generic_kernel,
// a plain smem-tiled DMMA kernel: same arithmetic, no fast path.  There is no CPU or library fallback.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace fm {
namespace {

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, double &x, double &y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct EpiParams {
    double alpha, beta, diag;
    int has_diag;
    int diag_col0;  // the diagonal term goes to the elements (row, col) with row == col + diag_col0 (C is a column block of a larger matrix)
    int npeers;
    double *peer[7];
    const int32_t *active_until;  // optional per-item round limit (batched squaring in expm)
    int round;
    unsigned zero;  // always 0; opaque to the compiler (see the stage-release token in the consumer loop)
    int pdl_wait;   // launched as a programmatic dependent of the previous kernel: wait for it before exiting
    int pdl_start;  // split-K chains of small products: launched as a programmatic dependent, lets ITS dependent (the reduction) become
                    // resident at once and waits for the previous kernel only after its own prologue (launch, barrier init, descriptor fetch)
    int ksplit;     // > 1: the k range is cut into ksplit slices, slice s of tile t writes its raw partial product to C + s * kstride
    long long kstride;
    int *tile_counter;  // non-null: tiles are handed out dynamically (atomic counter) instead of round-robin; see the producer loop
    int *counter_to_clear;  // a counter slot no kernel in flight uses: zeroed for a later launch
    int stagger_ns;  // two CTAs per SM: the second wave of CTAs starts this much later, so that one CTA's epilogue overlaps the other's main loop
    int copier;      // npeers > 0: the three spare warps of the producer warpgroup forward every finished tile to the peers (see below)
    int *spread;     // two CTAs per SM, static schedule, more tiles than CTAs: zeroed counters {low, high, per-SM arrivals[]}.  The CTAs renumber
                     // themselves so that the ids that receive the tiles of the last, partial round (the lowest) sit on distinct SMs
};

__device__ __forceinline__ void decode_tile(int tile, int tiles_m, int tiles_n, int &tm, int &tn, int &z) {
    constexpr int G = 16;  // band height in tiles
    const int per = tiles_m * tiles_n;
    z = tile / per;
    const int r = tile - z * per;
    const int band = r / (G * tiles_n);
    const int rr = r - band * (G * tiles_n);
    const int gsz = min(G, tiles_m - band * G);
    tn = rr / gsz;
    tm = band * G + (rr - tn * gsz);
}

// ---------------------------------------------------------------------------------------------
// TMA + DMMA persistent kernel (NN)
// ---------------------------------------------------------------------------------------------
template <int BM_, int BN_, int BK_, int WGM_, int WGN_, int STAGES_, int CTAS_PER_SM_ = 1>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, WGM = WGM_, WGN = WGN_, STAGES = STAGES_, CTAS_PER_SM = CTAS_PER_SM_;
    static constexpr int NW = WGM * WGN;            // consumer warps
    // Registers are allocated to a CTA in units of 4 warps, so the producer gets a whole warpgroup (3 of its warps exit
    // at once) and hands its registers to the consumers with setmaxnreg, the canonical warp-specialised layout.
    static constexpr int THREADS = (NW + 4) * 32;
    // 384 threads, 1 CTA/SM: 128*40 + 256*232 <= 65536;   256 threads, 2 CTAs/SM: 2*(128*40 + 128*216) = 65536
    static constexpr int REG_PRODUCER = 40, REG_CONSUMER = CTAS_PER_SM_ == 1 ? 232 : 216;
    static constexpr int WTM = BM / WGM, WTN = BN / WGN;
    static constexpr int MT = WTM / 8, NT = WTN / 8;  // m8n8 tiles per warp
    static constexpr int R = WTM / 16;                // 16-byte A loads per thread per k
    static constexpr int GPB = 8 / R;                 // fragment row-groups (g) per 16-row box
    static constexpr int A_BOX_BYTES = BK * 128;      // one box: BK rows (k) x 16 doubles (m)
    static constexpr int A_STAGE = (BM / 16) * A_BOX_BYTES;
    static constexpr int B_BOX_BYTES = BN * 128;      // one box: BN rows (n) x 16 doubles (k)
    static constexpr int B_STAGE = (BK / 16) * B_BOX_BYTES;
    static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2 * STAGES * 8 + STAGES * 4 /*tile ids*/ +
                                      96 /*copier: 4 mbarriers, 2 tile ids, a post counter per consumer warp (padded)*/;
    static_assert(BK % 16 == 0 && WTM % 16 == 0 && WTN % 8 == 0 && R >= 1 && R <= 4, "tile shape");
};

template <class Cfg, bool COPIER = false>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::CTAS_PER_SM)
dgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, double *__restrict__ C,
                 long long ldc, long long sc, int M, int N, int K, int tiles_m, int tiles_n, int total_tiles, EpiParams ep,
                 int vec_ok) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, NW = Cfg::NW, STAGES = Cfg::STAGES;
    constexpr int MT = Cfg::MT, NT = Cfg::NT, R = Cfg::R, GPB = Cfg::GPB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;  // full[s] then empty[s]
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tile_slot = bar_base + 16u * STAGES;  // dynamic scheduling: the tile whose FIRST k-step sits in stage s
    const bool dynamic = ep.tile_counter != nullptr;
    // Fused all-gather (column-sharded multi-GPU product): the consumers store a finished tile to the LOCAL C only and post its id in
    // a two-slot mailbox; the three spare warps of the producer warpgroup read the tile back (L2 hits) and store it to every peer's
    // C over NVLink while the consumers are already in the next tile's main loop.  (With the consumers issuing the 7 x 32 remote
    // stores themselves the DMMA pipe idled while the stores drained: 33.5 instead of 35.9 TFLOP/s per GPU at 8 GPUs.)
    // (constant offsets from smem_base, so the addresses cost no registers across the main loop)
    constexpr uint32_t COPY_BAR_OFF = (uint32_t)((STAGES * Cfg::STAGE_BYTES + 16 * STAGES + 4 * STAGES + 7) & ~7);
    const uint32_t copy_bar = smem_base + COPY_BAR_OFF;  // full[0], full[1], empty[0], empty[1]
    const uint32_t copy_tile = smem_base + COPY_BAR_OFF + 32u;
    auto copy_full = [&](int s) { return copy_bar + 8u * s; };
    auto copy_empty = [&](int s) { return copy_bar + 16u + 8u * s; };
    constexpr bool copier = COPIER;  // a separate instantiation: the plain product keeps its register allocation (168, no spills)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int s_cta;  // this CTA's position in the static tile schedule
    if (threadIdx.x == 0) {
        int id = (int)blockIdx.x;
        if (Cfg::CTAS_PER_SM == 2 && ep.spread != nullptr) {
            // first CTA to arrive on its SM: next id from the low end; second: next id from the high end.  Dense whatever the placement,
            // and the low ids -- which walk one tile more when the tile count is not a multiple of the grid -- are alone on their SMs
            // in that last round, where a CTA without a neighbour runs ~1.8x faster.
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int arrival = atomicAdd(ep.spread + 2 + smid, 1);
            id = arrival == 0 ? atomicAdd(ep.spread, 1) : (int)gridDim.x - 1 - atomicAdd(ep.spread + 1, 1);
        }
        s_cta = id;
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), NW); }
        for (int s = 0; s < 2; ++s) { mbar_init(copy_full(s), NW); mbar_init(copy_empty(s), 3); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (ep.pdl_start) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");  // everything the previous kernels wrote is visible from here on
    }
    const int KT = (K + BK - 1) / BK;

    if (warp >= NW) {
        // ===================== TMA producer warpgroup =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REG_PRODUCER));
        if (warp == NW && lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            // Dynamic scheduling (look-ahead LU: the trailing update shares the machine with the panel kernel, whose 16 SMs become
            // free half-way): the producer draws the next tile from a global counter and tells the consumers through a per-stage
            // slot written before the first k-step of the tile is armed; a negative id ends the consumers.  CTAs that only become
            // resident when the panel exits simply join in.
            if (dynamic && blockIdx.x == 0) *ep.counter_to_clear = 0;
            int tile = dynamic ? atomicAdd(ep.tile_counter, 1) : s_cta;
            for (;;) {
                if (tile >= total_tiles) {
                    if (dynamic) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        asm volatile("st.shared.s32 [%0], %1;" ::"r"(tile_slot + 4u * stage), "r"(-1) : "memory");
                        mbar_arrive(full_bar(stage));
                    }
                    break;
                }
                const int next = dynamic ? 0 : tile + (int)gridDim.x;
                int tm, tn, z;
                const int ks = tile % ep.ksplit;
                decode_tile(tile / ep.ksplit, tiles_m, tiles_n, tm, tn, z);
                if (ep.active_until && ep.round >= ep.active_until[z]) { tile = next; continue; }  // (never combined with dynamic)
                const int m0 = tm * BM, n0 = tn * BN;
                const int kt_per = (KT + ep.ksplit - 1) / ep.ksplit, kt_end = min(KT, (ks + 1) * kt_per);
                for (int kt = ks * kt_per; kt < kt_end; ++kt) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    if (dynamic && kt == ks * kt_per) asm volatile("st.shared.s32 [%0], %1;" ::"r"(tile_slot + 4u * stage), "r"(tile) : "memory");
                    const uint32_t fb = full_bar(stage);
                    mbar_expect_tx(fb, Cfg::STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_STAGE;
#pragma unroll
                    for (int b = 0; b < BM / 16; ++b) tma_load_3d(sa + b * Cfg::A_BOX_BYTES, &tmA, m0 + 16 * b, kt * BK, z, fb);
#pragma unroll
                    for (int kb = 0; kb < BK / 16; ++kb) tma_load_3d(sb + kb * Cfg::B_BOX_BYTES, &tmB, kt * BK + 16 * kb, n0, z, fb);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tile = dynamic ? atomicAdd(ep.tile_counter, 1) : next;
            }
        } else if (COPIER && warp > NW) {
            // ===================== tile forwarding to the peer GPUs =====================
            const int ct = threadIdx.x - (NW + 1) * 32;  // 0..95
            constexpr int CH = BM / 2;                    // 16-byte chunks per tile column
            int slot = 0;
            uint32_t cph = 0;
            for (;;) {
                mbar_wait(copy_full(slot), cph);
                int tile;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tile) : "r"(copy_tile + 4u * slot) : "memory");
                if (tile < 0) break;
                int tm, tn, z;
                decode_tile(tile, tiles_m, tiles_n, tm, tn, z);
                const long long zoff = z * sc;
                const int m0 = tm * BM, n0 = tn * BN;
                if (vec_ok && (M & 1) == 0) {
#pragma unroll 2
                    for (int idx = ct; idx < BN * CH; idx += 96) {
                        const int n = n0 + idx / CH, m = m0 + 2 * (idx % CH);
                        if (n < N && m < M) {
                            const long long off = zoff + (long long)n * ldc + m;
                            const double2 v = __ldcg(reinterpret_cast<const double2 *>(C + off));
#pragma unroll
                            for (int r = 0; r < 7; ++r)
                                if (r < ep.npeers) *reinterpret_cast<double2 *>(ep.peer[r] + off) = v;
                        }
                    }
                } else {
                    for (int idx = ct; idx < BN * BM; idx += 96) {
                        const int n = n0 + idx / BM, m = m0 + idx % BM;
                        if (n < N && m < M) {
                            const long long off = zoff + (long long)n * ldc + m;
                            const double v = __ldcg(C + off);
#pragma unroll
                            for (int r = 0; r < 7; ++r)
                                if (r < ep.npeers) ep.peer[r][off] = v;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(copy_empty(slot));
                slot ^= 1;
                if (slot == 0) cph ^= 1u;
            }
        }
        return;
    }

    // ===================== DMMA consumers =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REG_CONSUMER));
    if (Cfg::CTAS_PER_SM == 2 && ep.stagger_ns > 0 && blockIdx.x >= (gridDim.x + 1) / 2) {
        // Equal tiles keep the two CTAs of an SM in lock step: both in their epilogue (C read-modify-write) at once, DMMA pipe idle.
        // A fixed head start for the first wave of CTAs is kept for the whole launch, since every tile takes the same time.
        for (int t = 0; t < ep.stagger_ns; t += 1000) __nanosleep(1000);
    }
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % Cfg::WGM, wn = warp / Cfg::WGM;
    const int wm0 = wm * Cfg::WTM, wn0 = wn * Cfg::WTN;
    // A fragment rows: thread-group g owns, inside 16-row box (g / GPB), the 16-byte chunks c*GPB + g%GPB (c < R):
    // local rows 16*(g/GPB) + 2*(c*GPB + g%GPB) + {0,1}  -> m-tiles 2c, 2c+1.  k index for MMA j: 2t + (j&1) + 8*(j>>1).
    const int a_box = (wm0 >> 4) + g / GPB;
    const int a_chunk0 = g % GPB;
    uint32_t a_off[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int kr = 2 * t + e;  // row inside the box for j = e (j = e + 2 adds 8 rows = 1024 B)
        a_off[e] = a_box * Cfg::A_BOX_BYTES + kr * 128 + ((a_chunk0 ^ kr) << 4);
    }
    // B fragment columns: n-tile ni column  wn0 + 8*ni + perm(g), perm(g) = 4*(g&1) + (g>>1); k pairs: chunk t and t+4.
    const int permg = ((g & 1) << 2) | (g >> 1);
    const uint32_t b_off = (wn0 + permg) * 128 + ((t ^ permg) << 4);

    // what the consumers themselves store to: with the copier warps on duty, the local C only
    const int npeers_cons = copier ? 0 : ep.npeers;
    // every consumer warp keeps its post count in shared memory (touched once per tile; a register would be live across the main
    // loop, which has none to spare): mailbox slot of the next post = count & 1, its phase = (count >> 1) & 1
    const uint32_t post_cnt = copy_tile + 8u + 4u * warp;
    if (copier && lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(post_cnt), "r"(0u) : "memory");
    auto post = [&](int id) {  // all of this warp's stores of tile `id` are issued: release them to the copier warps
        __syncwarp();
        if (lane == 0) {
            unsigned nposts;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nposts) : "r"(post_cnt) : "memory");
            const int cslot = nposts & 1u;
            mbar_wait(copy_empty(cslot), ((nposts >> 1) & 1u) ^ 1u);  // the copier is done with the tile that used this slot two posts ago
            if (warp == 0) asm volatile("st.shared.s32 [%0], %1;" ::"r"(copy_tile + 4u * cslot), "r"(id) : "memory");
            mbar_arrive(copy_full(cslot));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(post_cnt), "r"(nposts + 1u) : "memory");
        }
    };
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = s_cta;; tile += gridDim.x) {
        if (dynamic) {  // the id of the tile whose first k-step is (or will be) in this stage
            mbar_wait(full_bar(stage), phase);
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(tile) : "r"(tile_slot + 4u * stage) : "memory");
            if (tile < 0) break;
        } else if (tile >= total_tiles) break;
        int tm, tn, z;
        const int ks = tile % ep.ksplit;
        decode_tile(tile / ep.ksplit, tiles_m, tiles_n, tm, tn, z);
        if (ep.active_until && ep.round >= ep.active_until[z]) continue;
        const int kt_per = (KT + ep.ksplit - 1) / ep.ksplit, kt_begin = ks * kt_per, kt_end = min(KT, (ks + 1) * kt_per);
        double acc[MT][NT][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        if (ep.beta != 0.0) {
            // rank-k updates (short k): the epilogue's C loads would pay the full DRAM latency twice per tile (two batches of loads),
            // a quarter of a k = 128 tile's time.  Pull this warp's WTM x WTN block of C into L2 now, one 128-byte line per lane and pass.
            const double *cw = C + z * sc + ks * ep.kstride + (long long)(tn * BN + wn0) * ldc + tm * BM + wm0;
#pragma unroll
            for (int i = 0; i < Cfg::WTN * (Cfg::WTM / 16) / 32; ++i) {
                const int line = lane + 32 * i, colw = line / (Cfg::WTM / 16), seg = line % (Cfg::WTM / 16);
                if (tn * BN + wn0 + colw < N && tm * BM + wm0 + 16 * seg < M)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(cw + (long long)colw * ldc + 16 * seg));
            }
        }

        for (int kt = kt_begin; kt < kt_end; ++kt) {
            mbar_wait(full_bar(stage), phase);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + Cfg::A_STAGE;
            unsigned tok = 0;  // folds in one word of every LDS result of this stage (see the release below)
#pragma unroll
            for (int kb = 0; kb < BK / 16; ++kb) {
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    double bf[NT][2];
#pragma unroll
                    for (int ni = 0; ni < NT; ++ni)
                    {
                        lds128((sb + kb * Cfg::B_BOX_BYTES + b_off + ni * 1024) ^ (jp << 6), bf[ni][0], bf[ni][1]);
                        tok ^= (unsigned)__double2loint(bf[ni][1]);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double af[MT];
#pragma unroll
                        for (int c = 0; c < R; ++c)
                        {
                            lds128((sa + a_off[e] + kb * 2048 + jp * 1024) ^ ((c * GPB) << 4), af[2 * c], af[2 * c + 1]);
                            tok ^= (unsigned)__double2loint(af[2 * c + 1]);
                        }
#pragma unroll
                        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                            for (int ni = 0; ni < NT; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni][e]);
                    }
                }
            }
            // Release the stage only after every LDS of the stage has RETURNED: the arrive's address depends (through a runtime
            // zero mask) on one word of each LDS result, so the scoreboard holds it back.  Without this ptxas hoists the arrive
            // above the last DMMAs, and when the LSU is congested (NVLink stores of other warps in the fused multi-GPU epilogue)
            // the TMA refill overtakes in-flight LDS: rare wrong A fragments.  Found with 2 GPUs, see DESIGN.md section 6.
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(stage) + (tok & ep.zero));
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }

        // ---------------- epilogue: registers -> global ----------------
        const int m_base = tm * BM + wm0 + 16 * (g / GPB) + 2 * a_chunk0;
        const int n_base = tn * BN + wn0;
        double *Cz = C + z * sc + ks * ep.kstride;
        const long long zoff = z * sc;
        if (vec_ok && npeers_cons == 0 && tm * BM + BM <= M && tn * BN + BN <= N) {
            // interior tile, the common case: no bounds, peer or per-element diagonal tests -- ~6 instructions per 16-byte store
            // instead of ~45 (ncu source view, batched 256^3 products: the generic epilogue was 1/3 of the DMMA count in
            // issued instructions and 12% of the consumer warps' time, which two CTAs per SM only partly hide).
            double *p0 = Cz + (long long)(n_base + t) * ldc + m_base;
            const bool dg = ep.has_diag && tm * BM < tn * BN + ep.diag_col0 + BN && tn * BN + ep.diag_col0 < tm * BM + BM;
            if (ep.beta == 0.0 && !dg) {
#pragma unroll
                for (int ni = 0; ni < NT; ++ni)
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        double *q = p0 + (long long)(8 * ni + 4 * x) * ldc;
#pragma unroll
                        for (int c = 0; c < R; ++c)
                            *reinterpret_cast<double2 *>(q + 2 * c * GPB) = make_double2(ep.alpha * acc[2 * c][ni][x], ep.alpha * acc[2 * c + 1][ni][x]);
                    }
                if (copier) post(tile);
                continue;
            }
            const int dm = m_base - (n_base + t + ep.diag_col0);  // element (c, ni, x) sits on the diagonal iff dm + 2*c*GPB == 8*ni + 4*x
            const double dgv = dg ? ep.diag : 0.0;
            if (ep.beta != 0.0) {
                double *q = p0;
#pragma unroll
                for (int ni = 0; ni < NT; ni += 2, q += 16 * ldc) {
                    // the C loads of two n-tiles are all in flight before the first dependent store; with load/use/store interleaved
                    // per element the stores fence the next load (possible aliasing) and every load pays a full memory latency
                    double2 old[2][2][R];
#pragma unroll
                    for (int d = 0; d < 2; ++d)
#pragma unroll
                        for (int x = 0; x < 2; ++x)
#pragma unroll
                            for (int c = 0; c < R; ++c)
                                if (ni + d < NT) old[d][x][c] = *reinterpret_cast<const double2 *>(q + (long long)(8 * d + 4 * x) * ldc + 2 * c * GPB);
#pragma unroll
                    for (int d = 0; d < 2; ++d)
#pragma unroll
                        for (int x = 0; x < 2; ++x)
#pragma unroll
                            for (int c = 0; c < R; ++c) {
                                if (ni + d >= NT) continue;
                                // the products depend on the loaded value, so ptxas cannot hoist all 64 of them above the loads (spills)
                                double v0 = fma(ep.alpha, acc[2 * c][ni + d][x], ep.beta * old[d][x][c].x);
                                double v1 = fma(ep.alpha, acc[2 * c + 1][ni + d][x], ep.beta * old[d][x][c].y);
                                const int off = 8 * (ni + d) + 4 * x - 2 * c * GPB;
                                if (dm == off) v0 += dgv;
                                if (dm + 1 == off) v1 += dgv;
                                *reinterpret_cast<double2 *>(q + (long long)(8 * d + 4 * x) * ldc + 2 * c * GPB) = make_double2(v0, v1);
                            }
                }
                if (copier) post(tile);
                continue;
            }
            double *q = p0;
#pragma unroll
            for (int ni = 0; ni < NT; ++ni, q += 8 * ldc)
#pragma unroll
                for (int x = 0; x < 2; ++x)
#pragma unroll
                    for (int c = 0; c < R; ++c) {
                        const int off = 8 * ni + 4 * x - 2 * c * GPB;
                        const double v0 = fma(ep.alpha, acc[2 * c][ni][x], dm == off ? dgv : 0.0);
                        const double v1 = fma(ep.alpha, acc[2 * c + 1][ni][x], dm + 1 == off ? dgv : 0.0);
                        *reinterpret_cast<double2 *>(q + (long long)(4 * x) * ldc + 2 * c * GPB) = make_double2(v0, v1);
                    }
            if (copier) post(tile);
            continue;
        }
        if (vec_ok && (M & 1) == 0) {
            // edge tiles and the fused multi-GPU store (16-byte pairs everywhere): when beta != 0 all C loads of a batch of two n-tiles are issued before
            // the first dependent FMA/store -- with load/use/store interleaved per element the stores fence the next load
            // (possible aliasing) and every load pays a full memory latency: ~27 us per tile for rank-k updates.
#pragma unroll
            for (int nb2 = 0; nb2 < NT; nb2 += 2) {
                double2 old[2][2][R];
#pragma unroll
                for (int d = 0; d < 2; ++d)
#pragma unroll
                    for (int x = 0; x < 2; ++x)
#pragma unroll
                        for (int c = 0; c < R; ++c) {
                            const int n = n_base + 8 * (nb2 + d) + (x ? 4 + t : t), m = m_base + 2 * c * GPB;
                            old[d][x][c] = (ep.beta != 0.0 && nb2 + d < NT && n < N && m < M)
                                               ? *reinterpret_cast<const double2 *>(Cz + (long long)n * ldc + m)
                                               : make_double2(0.0, 0.0);
                        }
#pragma unroll
                for (int d = 0; d < 2; ++d)
#pragma unroll
                    for (int x = 0; x < 2; ++x)
#pragma unroll
                        for (int c = 0; c < R; ++c) {
                            const int ni = nb2 + d;
                            const int n = n_base + 8 * ni + (x ? 4 + t : t), m = m_base + 2 * c * GPB;
                            if (ni >= NT || n >= N || m >= M) continue;
                            double v0 = ep.alpha * acc[2 * c][ni][x];
                            double v1 = ep.alpha * acc[2 * c + 1][ni][x];
                            if (ep.beta != 0.0) {
                                v0 += ep.beta * old[d][x][c].x;
                                v1 += ep.beta * old[d][x][c].y;
                            }
                            if (ep.has_diag) {
                                if (m == n + ep.diag_col0) v0 += ep.diag;
                                if (m + 1 == n + ep.diag_col0) v1 += ep.diag;
                            }
                            *reinterpret_cast<double2 *>(Cz + (long long)n * ldc + m) = make_double2(v0, v1);
#pragma unroll
                            for (int r = 0; r < 7; ++r)
                                if (r < npeers_cons)
                                    *reinterpret_cast<double2 *>(ep.peer[r] + zoff + (long long)n * ldc + m) = make_double2(v0, v1);
                        }
            }
            if (copier) post(tile);
            continue;
        }
#pragma unroll
        for (int ni = 0; ni < NT; ++ni) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int n = n_base + 8 * ni + (x ? 4 + t : t);
                if (n >= N) continue;
#pragma unroll
                for (int c = 0; c < R; ++c) {
                    const int m = m_base + 2 * c * GPB;
                    if (m >= M) continue;
                    double v0 = ep.alpha * acc[2 * c][ni][x];
                    double v1 = ep.alpha * acc[2 * c + 1][ni][x];
                    double *p = Cz + (long long)n * ldc + m;
                    const bool pair = vec_ok && (m + 1 < M);
                    if (ep.beta != 0.0) {
                        if (pair) {
                            const double2 o = *reinterpret_cast<const double2 *>(p);
                            v0 += ep.beta * o.x;
                            v1 += ep.beta * o.y;
                        } else {
                            v0 += ep.beta * p[0];
                            if (m + 1 < M) v1 += ep.beta * p[1];
                        }
                    }
                    if (ep.has_diag) {
                        if (m == n + ep.diag_col0) v0 += ep.diag;
                        if (m + 1 == n + ep.diag_col0) v1 += ep.diag;
                    }
                    if (pair) {
                        *reinterpret_cast<double2 *>(p) = make_double2(v0, v1);
#pragma unroll
                        for (int r = 0; r < 7; ++r)
                            if (r < npeers_cons)
                                *reinterpret_cast<double2 *>(ep.peer[r] + zoff + (long long)n * ldc + m) = make_double2(v0, v1);
                    } else {
                        p[0] = v0;
                        if (m + 1 < M) p[1] = v1;
#pragma unroll
                        for (int r = 0; r < 7; ++r)
                            if (r < npeers_cons) {
                                double *q = ep.peer[r] + zoff + (long long)n * ldc + m;
                                q[0] = v0;
                                if (m + 1 < M) q[1] = v1;
                            }
                    }
                }
            }
        }
        if (copier) post(tile);
    }
    if (copier) post(-1);  // end of this CTA's tiles
    // Look-ahead LU launches this kernel as a programmatic dependent of the panel kernel so that both run concurrently.
    // Nothing here reads the panel's output, but kernels launched after this one rely on stream order for BOTH: by waiting
    // for the primary grid before exiting, this grid's completion implies the panel's.
    if (ep.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Generic DMMA kernel: any transposition, alignment and leading dimension; 64x64x16 tiles, 4 warps of 32x32.
// ---------------------------------------------------------------------------------------------
constexpr int GB = 64, GK = 16, GPAD = 4;

__global__ void __launch_bounds__(128)
dgemm_generic_kernel(int ta, int tb, int M, int N, int K, const double *__restrict__ A, long long lda, long long sa_,
                     const double *__restrict__ B, long long ldb, long long sb_, double *C, long long ldc, long long sc, EpiParams ep) {
    __shared__ double As[GK][GB + GPAD];  // As[k][m]
    __shared__ double Bs[GK][GB + GPAD];  // Bs[k][n]
    const int z = blockIdx.z;
    if (ep.active_until && ep.round >= ep.active_until[z]) return;
    A += z * sa_;
    B += z * sb_;
    C += z * sc;
    const int m0 = blockIdx.x * GB, n0 = blockIdx.y * GB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int k0 = 0; k0 < K; k0 += GK) {
        // cooperative loads with bounds checks; index order follows the contiguous direction of each operand
        for (int idx = tid; idx < GB * GK; idx += 128) {
            int mm, kk;
            if (!ta) { mm = idx % GB; kk = idx / GB; } else { kk = idx % GK; mm = idx / GK; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < M && gk < K) v = ta ? A[(long long)gm * lda + gk] : A[(long long)gk * lda + gm];
            As[kk][mm] = v;
        }
        for (int idx = tid; idx < GB * GK; idx += 128) {
            int nn, kk;
            if (!tb) { kk = idx % GK; nn = idx / GK; } else { nn = idx % GB; kk = idx / GB; }
            const int gn = n0 + nn, gk = k0 + kk;
            double v = 0.0;
            if (gn < N && gk < K) v = tb ? B[(long long)gk * ldb + gn] : B[(long long)gn * ldb + gk];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < GK / 4; ++j) {
            double af[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) af[mi] = As[4 * j + t][wm0 + 8 * mi + g];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) bf[ni] = Bs[4 * j + t][wn0 + 8 * ni + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int m = m0 + wm0 + 8 * mi + g, n = n0 + wn0 + 8 * ni + 2 * t + x;
                if (m < M && n < N) {
                    double *p = C + (long long)n * ldc + m;
                    double v = ep.alpha * acc[mi][ni][x];
                    if (ep.beta != 0.0) v += ep.beta * *p;
                    if (ep.has_diag && m == n + ep.diag_col0) v += ep.diag;
                    *p = v;
#pragma unroll
                    for (int r = 0; r < 7; ++r)
                        if (r < ep.npeers) ep.peer[r][z * sc + (long long)n * ldc + m] = v;
                }
            }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map over a column-major (rows x cols) x batch operand: dim0 = rows (contiguous), dim1 = cols, dim2 = batch
int make_map(CUtensorMap *map, const double *base, long long rows, long long cols, long long ld, long long stride, int batch,
             int box_rows, int box_cols) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(FM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)cols, (cuuint64_t)batch};
    long long s2 = batch > 1 ? stride : ld * (cols > 0 ? cols : 1);
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8ull, (cuuint64_t)s2 * 8ull};
    cuuint32_t box[3] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(FM_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FM_OK;
}

// split-K second pass: C := alpha * sum_s W_s + beta * C (+ diag on the diagonal), fixed summation order
__global__ void __launch_bounds__(256) splitk_reduce_kernel(int m, int n, int ksplit, const double *__restrict__ w, long long ldw, long long wz,
                                                             long long kstride, double *c, long long ldc, long long sc, double alpha, double beta,
                                                             double diag, int has_diag, int diag_col0) {
    // launched as a programmatic dependent of the product that fills w: resident early, starts when the product has completed
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int z = blockIdx.z;
    for (int j = blockIdx.y; j < n; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
            double acc = 0.0;
            for (int s = 0; s < ksplit; ++s) acc += w[s * kstride + z * wz + j * ldw + i];
            double v = alpha * acc;
            double *p = c + z * sc + j * ldc + i;
            if (beta != 0.0) v += beta * *p;
            if (has_diag && i == j + diag_col0) v += diag;
            *p = v;
        }
}

using CfgBig = GemmCfg<128, 128, 16, 2, 4, 6>;
// 64x128 tiles, 4 consumer warps, two CTAs per SM: twice the tiles for problems that cannot fill 148 SMs with 128x128
// tiles, and with two independent CTAs on an SM one CTA's epilogue (C read-modify-write) overlaps the other's main loop
// -- the case of the rank-128/256 updates inside LU and the triangular solves.
using CfgHalf = GemmCfg<64, 128, 16, 1, 4, 4, 2>;

bool g_spread = getenv("FLOMAT_GEMM_NO_SPREAD") == nullptr;  // developer switch: no renumbering of the CTAs for a partial last round

template <class Cfg, bool COPIER = false>
int launch_tma(int m, int n, int k, const double *a, int lda, long long sa, const double *b, int ldb, long long sb, double *c, int ldc,
               long long sc, int batch, const EpiParams &ep, int max_sms) {
    static PerDevice<bool> attr_set;  // function attributes are per device
    if (!attr_set()) {
        FM_CUDA(cudaFuncSetAttribute(dgemm_tma_kernel<Cfg, COPIER>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set() = true;
    }
    CUtensorMap tmA, tmB;
    FM_TRY(make_map(&tmA, a, m, k, lda, sa, batch, 16, Cfg::BK));
    FM_TRY(make_map(&tmB, b, k, n, ldb, sb, batch, 16, Cfg::BN));
    const int tiles_m = ceil_div(m, Cfg::BM), tiles_n = ceil_div(n, Cfg::BN);
    const long long total = (long long)tiles_m * tiles_n * batch * ep.ksplit;
    if (total > 0x7fffffffLL) return set_error(FM_ERR_ARG, "dgemm: too many tiles");
    int sms = ctx().sm_count;
    if (max_sms > 0 && max_sms < sms && !ep.tile_counter) sms = max_sms;  // dynamic tiles: full grid, late CTAs join when SMs free up
    const long long slots = (long long)sms * Cfg::CTAS_PER_SM;
    const int grid = (int)(total < slots ? total : slots);
    const int vec_ok = ((reinterpret_cast<uintptr_t>(c) & 15u) == 0 && (ldc % 2) == 0 && (sc % 2) == 0) ? 1 : 0;
    int vec_all = vec_ok;
    for (int r = 0; r < ep.npeers; ++r)
        if (reinterpret_cast<uintptr_t>(ep.peer[r]) & 15u) vec_all = 0;
    EpiParams epl = ep;
    if (total < 4LL * grid) epl.stagger_ns = 0;  // the head start only pays when every CTA walks several tiles
    epl.spread = nullptr;
    if (Cfg::CTAS_PER_SM == 2 && g_spread && !ep.tile_counter && !ep.pdl_wait && !ep.pdl_start && total > grid && total % grid != 0 &&
        2 * (total % grid) <= grid) {
        // a partial last round of at most one tile per SM: see the renumbering at the top of the kernel.  (Not for launches that
        // must follow their predecessor directly as programmatic dependents: the memset would come between the two.)
        static PerDevice<int *> buf_pd;
        int *&buf = buf_pd();
        constexpr int WORDS = 2 + 1024;  // %smid stays far below 1024
        if (!buf) FM_CUDA(cudaMalloc(&buf, WORDS * sizeof(int)));
        FM_CUDA(cudaMemsetAsync(buf, 0, WORDS * sizeof(int), ctx().stream));
        epl.spread = buf;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = ctx().stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (ep.pdl_wait || ep.pdl_start) ? 1 : 0;
    FM_CUDA(cudaLaunchKernelEx(&cfg, dgemm_tma_kernel<Cfg, COPIER>, tmA, tmB, c, (long long)ldc, sc, m, n, k, tiles_m, tiles_n, (int)total, epl, vec_all));
    count_launch();
    return FM_OK;
}

bool g_half_enabled = getenv("FLOMAT_GEMM_NO_HALF") == nullptr;  // developer switches for A/B timing
bool g_splitk_enabled = getenv("FLOMAT_GEMM_NO_SPLITK") == nullptr;
bool g_stagger_all = getenv("FLOMAT_GEMM_STAGGER_ALL") != nullptr;
bool g_dynamic_enabled = getenv("FLOMAT_GEMM_NO_DYNAMIC") == nullptr;
// shortest k slice of a split-K product.  64 against 128 (round 1): 256^3 19.1 -> 14.8 us, 384^3 19.3 -> 15.1, 512^3 20.2 -> 19.6 (tools/small_gemm_time.py)
int g_splitk_min_k = getenv("FLOMAT_GEMM_SPLITK_MIN_K") && atoi(getenv("FLOMAT_GEMM_SPLITK_MIN_K")) >= 16 ? atoi(getenv("FLOMAT_GEMM_SPLITK_MIN_K")) : 64;
bool g_tail_rule = getenv("FLOMAT_GEMM_NO_TAIL_RULE") == nullptr;
bool g_force_half = getenv("FLOMAT_GEMM_FORCE_HALF") != nullptr;  // developer switch: 64x128 tiles, two CTAs per SM, for every shape
bool g_pdl_chain = getenv("FLOMAT_GEMM_NO_PDL_CHAIN") == nullptr;
bool g_copier_enabled = getenv("FLOMAT_GEMM_NO_COPIER") == nullptr;  // off: the consumer warps store to the peers themselves (round-1 epilogue)
int g_copier_min_peers = getenv("FLOMAT_GEMM_COPIER_MIN_PEERS") ? atoi(getenv("FLOMAT_GEMM_COPIER_MIN_PEERS")) : 3;
int g_stagger_ns = getenv("FLOMAT_GEMM_STAGGER_NS") ? atoi(getenv("FLOMAT_GEMM_STAGGER_NS")) : 6000;  // measured best for k = 128..256

}  // namespace

// Which kernel configuration a TMA-path product takes: 0 = 128x128 tiles (one CTA per SM), 1 = 64x128 tiles (two CTAs per SM);
// *ksplit > 1: 64x128 tiles with the k range cut into that many slices + the reduction kernel.  Pure host arithmetic (exported as
// fm_dgemm_tile_plan for the CPU tests).
//   * fewer than 2 x SMs tiles of 128x128, or k <= 256 (rank-k updates, triangular sweeps): 64x128 -- twice the tiles, two independent
//     CTAs per SM (one's epilogue under the other's main loop);
//   * otherwise wave quantisation decides (round 2, tools/gemm_wave_shapes.py).  In units of one 128x128 tile on one SM: 128x128 tiles
//     cost ceil(tiles / SMs); 64x128 tiles run two to an SM, a full round of 2 x SMs of them costs 1.01, and a last round of at most
//     ONE tile per SM costs 0.56 -- its CTAs have no neighbour (the kernel renumbers the CTAs so that this holds, see `spread`) and
//     run ~1.8x faster.  3.46 waves: 4 against 3.59 (30.8 -> 34.2 TFLOP/s measured), 2.16: 3 against 2.58 (25.5 -> 29.6), 4.11: 5
//     against 4.60; 2.7 / 3.89 / 6.92 / 27.7 waves stay with 128x128 tiles;
//   * at most SMs tiles of 64x128 and k >= 256 (n <= ~1100): one tile per CTA leaves most SMs idle while each warp walks k/16 steps
//     alone -- cut k into slices of at least 64 (one CTA each, raw partial products to a workspace) and reduce in a second pass.
int gemm_tile_plan(int m, int n, int k, int batch, int sms, bool splitk_allowed, int *ksplit_out) {
    const long long sms_ = sms > 0 ? sms : 148;
    const long long big_tiles = (long long)ceil_div(m, 128) * ceil_div(n, 128) * batch;
    const long long half_tiles = (long long)ceil_div(m, 64) * ceil_div(n, 128) * batch;
    const long long h_full = half_tiles / (2 * sms_), h_rest = half_tiles % (2 * sms_);
    const double t_big = (double)((big_tiles + sms_ - 1) / sms_);
    const double t_half = 1.01 * (double)h_full + (h_rest == 0 ? 0.0 : (h_rest <= sms_ && g_spread ? 0.56 : 1.01));
    const bool short_tail = g_tail_rule && t_half < t_big;
    const bool half = ((big_tiles < 2 * sms_ || k <= 256 || short_tail) && g_half_enabled) || g_force_half;
    int ksplit = 1;
    if (half && g_splitk_enabled && splitk_allowed && half_tiles <= sms_ && k >= 256) {
        ksplit = (int)((2 * sms_) / half_tiles);
        if (ksplit > k / g_splitk_min_k) ksplit = k / g_splitk_min_k;
        if (ksplit > 16) ksplit = 16;
        if (ksplit < 2) ksplit = 1;
    }
    if (ksplit_out) *ksplit_out = ksplit;
    return half ? 1 : 0;
}

int k_dgemm(int transA, int transB, int m, int n, int k, const double *a, int lda, long long sa, const double *b, int ldb, long long sb,
            double *c, int ldc, long long sc, int batch, const GemmEpilogue &epi) {
    const bool ta = (transA == FM_TRANS), tb = (transB == FM_TRANS);
    if ((transA != FM_TRANS && transA != FM_NO_TRANS) || (transB != FM_TRANS && transB != FM_NO_TRANS))
        return set_error(FM_ERR_ARG, "dgemm: trans flags must be 111 or 112");
    if (m < 0 || n < 0 || k < 0 || batch < 0) return set_error(FM_ERR_ARG, "dgemm: negative dimension");
    const int rows_a = ta ? k : m, rows_b = tb ? n : k;
    if (lda < (rows_a > 1 ? rows_a : 1) || ldb < (rows_b > 1 ? rows_b : 1) || ldc < (m > 1 ? m : 1))
        return set_error(FM_ERR_ARG, "dgemm: leading dimension too small (lda=%d ldb=%d ldc=%d)", lda, ldb, ldc);
    if (epi.npeers > 7) return set_error(FM_ERR_ARG, "dgemm: at most 7 peers");
    if (m == 0 || n == 0 || batch == 0) return FM_OK;
    EpiParams ep;
    ep.alpha = epi.alpha;
    ep.beta = epi.beta;
    ep.diag = epi.diag;
    ep.has_diag = epi.has_diag ? 1 : 0;
    ep.diag_col0 = epi.diag_col0;
    ep.npeers = epi.npeers;
    for (int r = 0; r < 7; ++r) ep.peer[r] = r < epi.npeers ? epi.peers[r] : nullptr;
    ep.active_until = epi.active_until;
    ep.round = epi.round;
    ep.zero = 0u;
    ep.pdl_wait = epi.pdl_dependent ? 1 : 0;
    ep.pdl_start = 0;
    ep.spread = nullptr;
    ep.tile_counter = nullptr;
    ep.counter_to_clear = nullptr;
    ep.ksplit = 1;
    ep.kstride = 0;
    ep.stagger_ns = ((epi.beta != 0.0 || g_stagger_all) && k <= 512) ? g_stagger_ns : 0;
    // forwarding by the spare warps pays from three peers on (measured, n = 16384: 8 GPUs 33.3 -> 34.9 TFLOP/s per GPU; with ONE peer the
    // consumers' own stores are 0.4 % faster: profiles/r02_colshard_copier_ab.jsonl)
    ep.copier = (epi.npeers >= g_copier_min_peers && g_copier_enabled) ? 1 : 0;

    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const bool tma_ok = !ta && !tb && k > 0 && al16(a) && al16(b) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                        (batch == 1 || (sa % 2 == 0 && sb % 2 == 0)) && encode_fn() != nullptr;
    if (tma_ok) {
        int ksplit = 1;
        const bool half = gemm_tile_plan(m, n, k, batch, ctx().sm_count, epi.npeers == 0 && !epi.active_until, &ksplit) != 0;
        {
            if (ksplit >= 2) {
                const long long ldw = m + (m & 1), wz = ldw * n, kstride = wz * batch;
                double *w = static_cast<double *>(scratch((size_t)kstride * ksplit * sizeof(double)));
                if (!w) return FM_ERR_NOMEM;
                EpiParams e2 = ep;
                e2.alpha = 1.0; e2.beta = 0.0; e2.has_diag = 0; e2.diag = 0.0;
                e2.ksplit = ksplit; e2.kstride = kstride;
                // product -> reduction -> next product ... are chained as programmatic dependents (FLOMAT_GEMM_NO_PDL_CHAIN: plain launches):
                // each kernel's launch and prologue overlap the tail of the one before, and each waits for it before touching memory
                e2.pdl_start = g_pdl_chain ? 1 : 0;
                FM_TRY((launch_tma<CfgHalf>(m, n, k, a, lda, sa, b, ldb, sb, w, (int)ldw, wz, batch, e2, epi.max_sms)));
                dim3 grid(ceil_div(m, 256), n < 1024 ? n : 1024, batch);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid;
                cfg.blockDim = dim3(256, 1, 1);
                cfg.stream = ctx().stream;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at;
                cfg.numAttrs = g_pdl_chain ? 1 : 0;
                FM_CUDA(cudaLaunchKernelEx(&cfg, splitk_reduce_kernel, m, n, ksplit, (const double *)w, ldw, wz, kstride, c, (long long)ldc, sc, ep.alpha,
                                           ep.beta, ep.diag, ep.has_diag, ep.diag_col0));
                count_launch();
                return FM_OK;
            }
        }
        if (epi.dynamic_tiles && g_dynamic_enabled && !epi.active_until) {
            // a ring of counters: this launch draws from slot i and clears slot i + 32 for a launch far in the future (at most a
            // couple of GEMMs are ever in flight at once)
            struct Counters { int *dev = nullptr; unsigned seq = 0; };
            static PerDevice<Counters> counters;
            Counters &cn = counters();
            if (!cn.dev) {
                FM_CUDA(cudaMalloc(&cn.dev, 64 * sizeof(int)));
                FM_CUDA(cudaMemsetAsync(cn.dev, 0, 64 * sizeof(int), ctx().stream));
            }
            ep.tile_counter = cn.dev + (cn.seq % 64);
            ep.counter_to_clear = cn.dev + ((cn.seq + 32) % 64);
            ++cn.seq;
        }
        if (ep.copier) {  // the fused all-gather: the instantiation whose spare producer-group warps forward the tiles to the peers
            if (half) return launch_tma<CfgHalf, true>(m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch, ep, epi.max_sms);
            return launch_tma<CfgBig, true>(m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch, ep, epi.max_sms);
        }
        if (half) return launch_tma<CfgHalf>(m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch, ep, epi.max_sms);
        return launch_tma<CfgBig>(m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, batch, ep, epi.max_sms);
    }

    // Operands the tensor maps cannot describe (transposed, 8-byte aligned views, odd leading dimensions): for products that are
    // worth it, bring the offending operand into an aligned, untransposed workspace (16 B/element of extra traffic against
    // 2k flops/element) and take the TMA path; small products stay on the generic kernel.
    if (batch == 1 && k > 0 && encode_fn() != nullptr && (double)m * n * k >= 16777216.0 && epi.npeers == 0) {
        const bool fix_a = ta || !al16(a) || (lda % 2 != 0), fix_b = tb || !al16(b) || (ldb % 2 != 0);
        const int ldwa = m + (m & 1), ldwb = k + (k & 1);
        double *wa = nullptr, *wb = nullptr;
        int rc = FM_OK;
        if (fix_a) {
            wa = static_cast<double *>(fm_alloc_bytes((size_t)ldwa * k * sizeof(double)));
            if (!wa) return FM_ERR_NOMEM;
            rc = ta ? k_transpose(k, m, a, lda, wa, ldwa) : k_copy2d(wa, ldwa, a, lda, m, k);
        }
        if (rc == FM_OK && fix_b) {
            wb = static_cast<double *>(fm_alloc_bytes((size_t)ldwb * n * sizeof(double)));
            if (!wb) rc = FM_ERR_NOMEM;
            else rc = tb ? k_transpose(n, k, b, ldb, wb, ldwb) : k_copy2d(wb, ldwb, b, ldb, k, n);
        }
        if (rc == FM_OK)
            rc = k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, n, k, fix_a ? wa : a, fix_a ? ldwa : lda, 0, fix_b ? wb : b, fix_b ? ldwb : ldb, 0, c, ldc, sc, 1, epi);
        if (wa) fm_free(wa);  // cached blocks are reused in stream order
        if (wb) fm_free(wb);
        return rc;
    }

    dim3 grid(ceil_div(m, GB), ceil_div(n, GB), batch);
    if (grid.y > 65535 || grid.z > 65535) return set_error(FM_ERR_ARG, "dgemm(generic): grid too large");
    dgemm_generic_kernel<<<grid, 128, 0, ctx().stream>>>(ta ? 1 : 0, tb ? 1 : 0, m, n, k, a, lda, sa, b, ldb, sb, c, ldc, sc, ep);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

}  // namespace fm
