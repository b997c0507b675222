// mg.cu -- multi-GPU entry points for ONE host process and ONE calling thread, which is how the reference calls its BLAS
// (`flomat*` -> cblas_dgemm, flomat.rkt:905-914 / :888, one Racket OS thread).  fm_mg_init(ndev) starts one worker thread
// per device (each bound to its device's Context) and enables peer access between all pairs; a call hands one closure to every
// worker and returns when all of them have queued (device operands) or finished (host operands) their share.
//
//   fm_mg_dgemm        C := alpha A B + beta C, sharded by column blocks of B and C: A replicated, block j on device j.
//     host operands    every device uploads ITS column slice of A over its own PCIe link, the slices are all-gathered over NVLink
//                      (peer copies on the upload streams), then block j of B / C runs through the same upload | product | download
//                      column-block pipeline as fm_dgemm_host: 8 PCIe links carry 1/8 of the bytes each, A crosses PCIe once.
//     device operands  (all on one "home" device) the other devices pull their slice of A from home, all-gather the rest among
//                      themselves, pull their block of B, and the GEMM epilogue itself stores every finished tile of C into the
//                      home device's C over NVLink (the fused peer-store epilogue of dgemm.cu, one peer) -- no gather pass.
//                      Completion is event based: the home stream waits for one event per device; the call does not block.
//   fm_mg_expm_batched  independent items in contiguous blocks per device, no exchange (host or home-device operands).
//
// LU / solve / inv / det stay single-GPU (north_star: "LU stays single-GPU").  No arithmetic happens on the host.
#include "common.cuh"
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fm {
namespace {

constexpr int MG_MAX = 8;  // the peer-store epilogue addresses at most 7 peers; one NVSwitch domain has 8 GPUs

// Persistent worker threads, one per device.  run(f) executes f(rank) on every worker at once and returns the first failure.
// agree(rc) is a barrier among the workers INSIDE a job that also spreads a failure: every worker leaves with the same verdict, so
// nobody waits at a later barrier for a worker that has bailed out.
class Workers {
  public:
    int start(int n, const int *devices) {
        n_ = n;
        fail_.assign(n, 0);
        msg_.assign(n, std::string());
        for (int r = 0; r < n; ++r) std::thread([this, r, d = devices[r]] { loop(r, d); }).detach();  // leaked on purpose (see xfer.cu)
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [this] { return booted_ == n_; });
        for (int r = 0; r < n; ++r)
            if (fail_[r]) return set_error(fail_[r], "fm_mg_init: device of worker %d: %s", r, msg_[r].c_str());
        return FM_OK;
    }
    int run(const std::function<int(int)> &f) {
        std::unique_lock<std::mutex> lk(mu_);
        job_ = &f;
        left_ = n_;
        arrived_ = 0;
        verdict_ = FM_OK;
        std::fill(fail_.begin(), fail_.end(), 0);
        ++gen_;
        cv_job_.notify_all();
        cv_done_.wait(lk, [this] { return left_ == 0; });
        job_ = nullptr;
        for (int r = 0; r < n_; ++r)
            if (fail_[r]) return set_error(fail_[r], "%s", msg_[r].c_str());
        return FM_OK;
    }
    int agree(int rc) {
        std::unique_lock<std::mutex> lk(mu_);
        if (rc != FM_OK && verdict_ == FM_OK) verdict_ = rc;
        const unsigned long long my = bar_gen_;
        if (++arrived_ == n_) {
            arrived_ = 0;
            last_verdict_ = verdict_;
            ++bar_gen_;
            cv_bar_.notify_all();
        } else {
            cv_bar_.wait(lk, [&] { return bar_gen_ != my; });
        }
        return last_verdict_;
    }

  private:
    void loop(int r, int device) {
        int rc = bind_device(device);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (rc != FM_OK) { fail_[r] = rc; msg_[r] = error_text(); }
            if (++booted_ == n_) cv_done_.notify_all();
        }
        unsigned long long seen = 0;
        for (;;) {
            const std::function<int(int)> *f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_job_.wait(lk, [&] { return gen_ != seen && job_ != nullptr; });
                seen = gen_;
                f = job_;
            }
            rc = (*f)(r);
            std::lock_guard<std::mutex> lk(mu_);
            if (rc != FM_OK) { fail_[r] = rc; msg_[r] = error_text(); }
            if (--left_ == 0) cv_done_.notify_all();
        }
    }
    int n_ = 0, booted_ = 0, left_ = 0, arrived_ = 0, verdict_ = FM_OK, last_verdict_ = FM_OK;
    unsigned long long gen_ = 0, bar_gen_ = 0;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_, cv_bar_;
    const std::function<int(int)> *job_ = nullptr;
    std::vector<int> fail_;
    std::vector<std::string> msg_;
};

struct Group {
    int ndev = 0;
    int dev[MG_MAX] = {};
    Workers workers;
    cudaEvent_t ev_a[MG_MAX] = {};     // device r: its own slice of A is in place
    cudaEvent_t ev_pull[MG_MAX] = {};  // device r: has finished reading its peers' slices
    cudaEvent_t ev_done[MG_MAX] = {};  // device r: its share of the call is complete
    cudaEvent_t ev_home = nullptr;     // home stream: everything queued before the call
    int ev_home_dev = -1;
    double *da[MG_MAX] = {};           // per call: each device's copy of A
    cudaEvent_t ev_bar[MG_MAX][2] = {};  // fm_mg_expm: device r has reached group barrier k (slot k & 1)
    double *ws[MG_MAX][6] = {};          // fm_mg_expm: every device's six workspaces
};
Group *g_group = nullptr;

int rank_of_device(const Group &g, int device) {
    for (int r = 0; r < g.ndev; ++r)
        if (g.dev[r] == device) return r;
    return -1;
}

// column block of rank r: whole 128-column tile columns, remainder to the last ranks first (same rule as sci_b200/multigpu.py)
void column_block(int n, int world, int r, int *j0, int *nloc) {
    const int tiles = (n + 127) / 128, base = tiles / world, extra = tiles % world;
    const int t0 = r * base + std::min(r, extra), t1 = t0 + base + (r < extra ? 1 : 0);
    const int a = std::min(t0 * 128, n), b = std::min(t1 * 128, n);
    *j0 = a;
    *nloc = b - a;
}
void item_block(int batch, int world, int r, int *z0, int *cnt) {
    const int base = batch / world, extra = batch % world;
    *z0 = r * base + std::min(r, extra);
    *cnt = base + (r < extra ? 1 : 0);
}

#define MG_CUDA(rc, call)                                                                                                     \
    do {                                                                                                                      \
        if ((rc) == FM_OK) {                                                                                                  \
            cudaError_t e__ = (call);                                                                                         \
            if (e__ != cudaSuccess) (rc) = set_error(FM_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
        }                                                                                                                     \
    } while (0)

// dst (on this device) := src (on device `src_dev`), rows x cols doubles; one flat peer copy when both are dense
int pull2d(double *dst, size_t ldd, const double *src, size_t lds, int src_dev, size_t rows, size_t cols, cudaStream_t st) {
    if (rows == 0 || cols == 0) return FM_OK;
    if (ldd == rows && lds == rows) {
        FM_CUDA(cudaMemcpyPeerAsync(dst, ctx().device, src, src_dev, rows * cols * 8, st));
    } else {
        FM_CUDA(cudaMemcpy2DAsync(dst, ldd * 8, src, lds * 8, rows * 8, cols, cudaMemcpyDefault, st));
    }
    return FM_OK;
}

}  // namespace
}  // namespace fm

using namespace fm;

extern "C" {

int fm_mg_init(int ndev) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_error(FM_ERR_NOGPU, "fm_mg_init: no CUDA device available; libflomat_cuda has no CPU fallback");
    }
    if (ndev <= 0) ndev = count < MG_MAX ? count : MG_MAX;
    if (ndev > count || ndev > MG_MAX) return set_error(FM_ERR_ARG, "fm_mg_init: %d devices requested, %d visible, at most %d per group", ndev, count, MG_MAX);
    if (g_group) {
        if (g_group->ndev == ndev) return FM_OK;
        return set_error(FM_ERR_ARG, "fm_mg_init: a group of %d devices already exists (one group per process)", g_group->ndev);
    }
    FM_TRY(ensure_init());
    // the calling thread's device is rank 0 (operands that live on a device are expected there or on any other device of the group)
    Group *g = new Group;
    g->ndev = ndev;
    const int first = ctx().device;
    for (int r = 0, d = first; r < ndev; ++r, d = (d + 1) % count) g->dev[r] = d;
    FM_TRY(g->workers.start(ndev, g->dev));
    const int rc = g->workers.run([g](int r) -> int {
        int rc = FM_OK;
        for (int p = 0; p < g->ndev && rc == FM_OK; ++p) {
            if (p == r) continue;
            int can = 0;
            MG_CUDA(rc, cudaDeviceCanAccessPeer(&can, g->dev[r], g->dev[p]));
            if (rc == FM_OK && !can) rc = set_error(FM_ERR_CUDA, "fm_mg_init: device %d cannot access device %d as a peer", g->dev[r], g->dev[p]);
            if (rc == FM_OK) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[p], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
                else if (e != cudaSuccess) rc = set_error(FM_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", g->dev[r], g->dev[p], cudaGetErrorString(e));
            }
        }
        MG_CUDA(rc, cudaEventCreateWithFlags(&g->ev_a[r], cudaEventDisableTiming));
        MG_CUDA(rc, cudaEventCreateWithFlags(&g->ev_pull[r], cudaEventDisableTiming));
        MG_CUDA(rc, cudaEventCreateWithFlags(&g->ev_done[r], cudaEventDisableTiming));
        MG_CUDA(rc, cudaEventCreateWithFlags(&g->ev_bar[r][0], cudaEventDisableTiming));
        MG_CUDA(rc, cudaEventCreateWithFlags(&g->ev_bar[r][1], cudaEventDisableTiming));
        return rc;
    });
    if (rc != FM_OK) return rc;  // (the workers stay parked; a later fm_mg_init may try again with the same count)
    g_group = g;
    return FM_OK;
}

int fm_mg_device_count(void) { return g_group ? g_group->ndev : 0; }

// the partition fm_mg_dgemm / fm_mg_expm_batched use (pure host arithmetic, callable without a GPU): block of rank r of `world`
int fm_mg_column_block(int n, int world, int r, int *first, int *count) {
    if (n < 0 || world <= 0 || r < 0 || r >= world || !first || !count) return set_error(FM_ERR_ARG, "fm_mg_column_block: bad arguments");
    column_block(n, world, r, first, count);
    return FM_OK;
}
int fm_mg_item_block(int batch, int world, int r, int *first, int *count) {
    if (batch < 0 || world <= 0 || r < 0 || r >= world || !first || !count) return set_error(FM_ERR_ARG, "fm_mg_item_block: bad arguments");
    item_block(batch, world, r, first, count);
    return FM_OK;
}

int fm_mg_dgemm(int transA, int transB, int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta,
                double *c, int ldc) {
    if (!g_group) FM_TRY(fm_mg_init(0));
    Group &g = *g_group;
    if (transA != FM_NO_TRANS || transB != FM_NO_TRANS)
        return set_error(FM_ERR_ARG, "fm_mg_dgemm: only the untransposed product is sharded (what `times` issues, flomat.rkt:905-914); use fm_dgemm");
    if (m < 0 || n < 0 || k < 0) return set_error(FM_ERR_ARG, "fm_mg_dgemm: negative dimension");
    if (lda < (m > 1 ? m : 1) || ldb < (k > 1 ? k : 1) || ldc < (m > 1 ? m : 1)) return set_error(FM_ERR_ARG, "fm_mg_dgemm: leading dimension too small");
    if (m == 0 || n == 0) return FM_OK;
    int da_dev = -1, db_dev = -1, dc_dev = -1;
    const bool a_on = is_device_pointer(a, &da_dev), b_on = is_device_pointer(b, &db_dev), c_on = is_device_pointer(c, &dc_dev);
    if (k == 0 || g.ndev == 1 || n < 256) {  // nothing to shard: the single-device entry points do the right thing
        if (!a_on && !b_on && !c_on) return fm_dgemm_host(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 0);
        if (a_on && b_on && c_on) return fm_dgemm(transA, transB, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
        return set_error(FM_ERR_ARG, "fm_mg_dgemm: operands must be all host or all device memory");
    }

    if (!a_on && !b_on && !c_on) {
        // ---------------- host operands: 1/ndev of the PCIe traffic per link, A all-gathered over NVLink ----------------
        const int ldda = (m + (m & 1)) < 2 ? 2 : m + (m & 1);
        return g.workers.run([&](int r) -> int {
            int rc = FM_OK;
            int j0, nloc, k0, kloc;
            column_block(n, g.ndev, r, &j0, &nloc);
            column_block(k, g.ndev, r, &k0, &kloc);  // the slice of A's columns this device fetches from the host
            cudaStream_t up = nullptr;
            rc = xfer_streams(&up, nullptr);
            g.da[r] = rc == FM_OK ? static_cast<double *>(fm_alloc_bytes((size_t)ldda * k * 8)) : nullptr;
            if (rc == FM_OK && !g.da[r]) rc = FM_ERR_NOMEM;
            // (a recycled block may still be in use by kernels queued earlier on this device)
            MG_CUDA(rc, cudaEventRecord(g.ev_done[r], ctx().stream));
            MG_CUDA(rc, cudaStreamWaitEvent(up, g.ev_done[r], 0));
            if (rc == FM_OK && kloc > 0) rc = copy_h2d(g.da[r] + (size_t)k0 * ldda, ldda, a + (size_t)k0 * lda, lda, m, kloc, up);
            MG_CUDA(rc, cudaEventRecord(g.ev_a[r], up));
            rc = g.workers.agree(rc);  // every slice is on its way and every g.da[] is known
            if (rc != FM_OK) { if (g.da[r]) fm_free(g.da[r]); return rc; }
            for (int i = 1; i < g.ndev && rc == FM_OK; ++i) {  // start with the right-hand neighbour: no two devices read the same peer at once
                const int p = (r + i) % g.ndev;
                int pk0, pkloc;
                column_block(k, g.ndev, p, &pk0, &pkloc);
                if (pkloc == 0) continue;
                MG_CUDA(rc, cudaStreamWaitEvent(up, g.ev_a[p], 0));
                MG_CUDA(rc, cudaMemcpyPeerAsync(g.da[r] + (size_t)pk0 * ldda, g.dev[r], g.da[p] + (size_t)pk0 * ldda, g.dev[p], (size_t)ldda * pkloc * 8, up));
            }
            MG_CUDA(rc, cudaEventRecord(g.ev_pull[r], up));
            if (rc == FM_OK && nloc > 0)
                rc = dgemm_host_pipeline(m, nloc, k, alpha, nullptr, 0, g.da[r], ldda, b + (size_t)j0 * ldb, ldb, beta, c + (size_t)j0 * ldc, ldc, 0);
            MG_CUDA(rc, cudaStreamSynchronize(up));
            rc = g.workers.agree(rc);  // nobody frees its copy of A while a peer may still be reading from it
            fm_free(g.da[r]);
            g.da[r] = nullptr;
            return rc;
        });
    }
    if (!(a_on && b_on && c_on)) return set_error(FM_ERR_ARG, "fm_mg_dgemm: operands must be all host or all device memory");
    if (da_dev != db_dev || da_dev != dc_dev) return set_error(FM_ERR_ARG, "fm_mg_dgemm: device operands must live on ONE device (A on %d, B on %d, C on %d)", da_dev, db_dev, dc_dev);
    const int home = rank_of_device(g, da_dev);
    if (home < 0) return set_error(FM_ERR_ARG, "fm_mg_dgemm: the operands' device %d is not part of the group", da_dev);

    // ---------------- device operands on `home` ----------------
    // everything queued so far on the home device's library stream (the producers of A, B, C) comes first
    FM_TRY(ensure_init());
    const int caller_dev = ctx().device;
    if (caller_dev != g.dev[home]) return set_error(FM_ERR_ARG, "fm_mg_dgemm: call from the thread bound to the operands' device (%d), not %d", g.dev[home], caller_dev);
    if (!g.ev_home || g.ev_home_dev != caller_dev) {
        if (g.ev_home) cudaEventDestroy(g.ev_home);
        FM_CUDA(cudaEventCreateWithFlags(&g.ev_home, cudaEventDisableTiming));
        g.ev_home_dev = caller_dev;
    }
    FM_CUDA(cudaEventRecord(g.ev_home, ctx().stream));
    const int ldda = (m + (m & 1)) < 2 ? 2 : m + (m & 1);
    const int rc = g.workers.run([&](int r) -> int {
        int rc = FM_OK;
        cudaStream_t st = ctx().stream;
        int j0, nloc, k0, kloc;
        column_block(n, g.ndev, r, &j0, &nloc);
        column_block(k, g.ndev, r, &k0, &kloc);
        const bool is_home = (r == home);
        double *dB = nullptr, *dC = nullptr;
        const int lddb = (k + (k & 1)) < 2 ? 2 : k + (k & 1);
        g.da[r] = nullptr;
        MG_CUDA(rc, cudaStreamWaitEvent(st, g.ev_home, 0));
        if (!is_home && rc == FM_OK) {
            g.da[r] = static_cast<double *>(fm_alloc_bytes((size_t)ldda * k * 8));
            if (!g.da[r]) rc = FM_ERR_NOMEM;
            if (rc == FM_OK && kloc > 0) rc = pull2d(g.da[r] + (size_t)k0 * ldda, ldda, a + (size_t)k0 * lda, lda, g.dev[home], m, kloc, st);
        }
        MG_CUDA(rc, cudaEventRecord(g.ev_a[r], st));
        rc = g.workers.agree(rc);
        if (rc != FM_OK) { if (g.da[r]) fm_free(g.da[r]); return rc; }
        if (!is_home) {
            for (int i = 1; i < g.ndev && rc == FM_OK; ++i) {
                const int p = (r + i) % g.ndev;
                int pk0, pkloc;
                column_block(k, g.ndev, p, &pk0, &pkloc);
                if (pkloc == 0) continue;
                if (p == home) {  // the home device holds all of A: its own slice comes from the original
                    rc = pull2d(g.da[r] + (size_t)pk0 * ldda, ldda, a + (size_t)pk0 * lda, lda, g.dev[home], m, pkloc, st);
                } else {
                    MG_CUDA(rc, cudaStreamWaitEvent(st, g.ev_a[p], 0));
                    MG_CUDA(rc, cudaMemcpyPeerAsync(g.da[r] + (size_t)pk0 * ldda, g.dev[r], g.da[p] + (size_t)pk0 * ldda, g.dev[p], (size_t)ldda * pkloc * 8, st));
                }
            }
            if (rc == FM_OK && nloc > 0) {
                dB = static_cast<double *>(fm_alloc_bytes((size_t)lddb * nloc * 8));
                dC = static_cast<double *>(fm_alloc_bytes((size_t)ldc * nloc * 8));  // same leading dimension as the home C: the peer store reuses the offsets
                if (!dB || !dC) rc = FM_ERR_NOMEM;
                if (rc == FM_OK) rc = pull2d(dB, lddb, b + (size_t)j0 * ldb, ldb, g.dev[home], k, nloc, st);
                if (rc == FM_OK && beta != 0.0) rc = pull2d(dC, ldc, c + (size_t)j0 * ldc, ldc, g.dev[home], m, nloc, st);
            }
        }
        MG_CUDA(rc, cudaEventRecord(g.ev_pull[r], st));
        rc = g.workers.agree(rc);
        if (rc == FM_OK) {
            // a device's copy of A is recycled (stream order) only after every peer has finished reading from it
            for (int p = 0; p < g.ndev; ++p)
                if (p != r) MG_CUDA(rc, cudaStreamWaitEvent(st, g.ev_pull[p], 0));
        }
        if (rc == FM_OK && nloc > 0) {
            GemmEpilogue ep;
            ep.alpha = alpha;
            ep.beta = beta;
            if (is_home) {
                rc = k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, nloc, k, a, lda, 0, b + (size_t)j0 * ldb, ldb, 0, c + (size_t)j0 * ldc, ldc, 0, 1, ep);
            } else {
                double *peer[1] = {c + (size_t)j0 * ldc};  // every finished tile goes straight into the home device's C over NVLink
                ep.peers = peer;
                ep.npeers = 1;
                rc = k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, nloc, k, g.da[r], ldda, 0, dB, lddb, 0, dC, ldc, 0, 1, ep);
            }
        }
        MG_CUDA(rc, cudaEventRecord(g.ev_done[r], st));
        if (g.da[r]) fm_free(g.da[r]);
        if (dB) fm_free(dB);
        if (dC) fm_free(dC);
        g.da[r] = nullptr;
        return rc;
    });
    if (rc != FM_OK) return rc;
    for (int r = 0; r < g.ndev; ++r)
        if (r != home) FM_CUDA(cudaStreamWaitEvent(ctx().stream, g.ev_done[r], 0));
    return FM_OK;
}

int fm_mg_expm_batched(int n, const double *a, int lda, long long stride_a, double *out, int ldo, long long stride_o, int batch, int mode,
                       double *s_out) {
    if (!g_group) FM_TRY(fm_mg_init(0));
    Group &g = *g_group;
    if (n < 0 || batch < 0 || lda < (n > 1 ? n : 1) || ldo < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "fm_mg_expm_batched: bad dimensions");
    if (n == 0 || batch == 0) return FM_OK;
    if (stride_a < (long long)lda * n || stride_o < (long long)ldo * n) return set_error(FM_ERR_ARG, "fm_mg_expm_batched: items overlap (stride < ld * n)");
    int a_dev = -1, o_dev = -1;
    const bool a_on = is_device_pointer(a, &a_dev), o_on = is_device_pointer(out, &o_dev);
    if (a_on != o_on) return set_error(FM_ERR_ARG, "fm_mg_expm_batched: input and output must be both host or both device memory");
    int home = -1;
    if (a_on) {
        if (a_dev != o_dev) return set_error(FM_ERR_ARG, "fm_mg_expm_batched: device operands must live on one device");
        home = rank_of_device(g, a_dev);
        if (home < 0) return set_error(FM_ERR_ARG, "fm_mg_expm_batched: the operands' device %d is not part of the group", a_dev);
        FM_TRY(ensure_init());
        if (ctx().device != a_dev) return set_error(FM_ERR_ARG, "fm_mg_expm_batched: call from the thread bound to the operands' device (%d)", a_dev);
        if (!g.ev_home || g.ev_home_dev != a_dev) {
            if (g.ev_home) cudaEventDestroy(g.ev_home);
            FM_CUDA(cudaEventCreateWithFlags(&g.ev_home, cudaEventDisableTiming));
            g.ev_home_dev = a_dev;
        }
        FM_CUDA(cudaEventRecord(g.ev_home, ctx().stream));
    }
    const int rc = g.workers.run([&](int r) -> int {
        int z0, cnt;
        item_block(batch, g.ndev, r, &z0, &cnt);
        if (cnt == 0) return FM_OK;
        int rc = FM_OK;
        cudaStream_t st = ctx().stream;
        const double *src = a + (size_t)z0 * stride_a;
        double *dst = out + (size_t)z0 * stride_o;
        double *s_r = s_out ? s_out + z0 : nullptr;
        if (a_on && r == home) {
            rc = k_expm_batched(n, src, lda, stride_a, dst, ldo, stride_o, cnt, mode, s_r);
            MG_CUDA(rc, cudaEventRecord(g.ev_done[r], st));
            return rc;
        }
        // a dense local copy of the items and of the results: n x n each, item stride n*n
        const size_t per = (size_t)n * n;
        double *x = static_cast<double *>(fm_alloc_bytes(per * cnt * 8)), *e = static_cast<double *>(fm_alloc_bytes(per * cnt * 8));
        if (!x || !e) rc = FM_ERR_NOMEM;
        if (a_on) MG_CUDA(rc, cudaStreamWaitEvent(st, g.ev_home, 0));
        if (rc == FM_OK) {
            if (lda == n) {  // every item is one dense run of n*n doubles: one strided copy for the whole block
                if (a_on) rc = pull2d(x, per, src, (size_t)stride_a, a_dev, per, cnt, st);
                else rc = copy_h2d(x, per, src, (size_t)stride_a, per, cnt, st);
            } else {
                for (int z = 0; z < cnt && rc == FM_OK; ++z) {
                    if (a_on) rc = pull2d(x + z * per, n, src + (size_t)z * stride_a, lda, a_dev, n, n, st);
                    else rc = copy_h2d(x + z * per, n, src + (size_t)z * stride_a, lda, n, n, st);
                }
            }
        }
        if (rc == FM_OK) rc = k_expm_batched(n, x, n, (long long)per, e, n, (long long)per, cnt, mode, s_r);
        if (rc == FM_OK) {
            if (a_on) {  // push the results into the home device's buffer (UVA: a plain 2-D copy between peers)
                if (ldo == n) MG_CUDA(rc, cudaMemcpy2DAsync(dst, (size_t)stride_o * 8, e, per * 8, per * 8, cnt, cudaMemcpyDefault, st));
                else
                    for (int z = 0; z < cnt && rc == FM_OK; ++z)
                        MG_CUDA(rc, cudaMemcpy2DAsync(dst + (size_t)z * stride_o, (size_t)ldo * 8, e + z * per, (size_t)n * 8, (size_t)n * 8, n, cudaMemcpyDefault, st));
            } else {
                if (ldo == n) rc = copy_d2h(dst, (size_t)stride_o, e, per, per, cnt, st);
                else
                    for (int z = 0; z < cnt && rc == FM_OK; ++z) rc = copy_d2h(dst + (size_t)z * stride_o, ldo, e + z * per, n, n, n, st);
                MG_CUDA(rc, cudaStreamSynchronize(st));
            }
        }
        MG_CUDA(rc, cudaEventRecord(g.ev_done[r], st));
        if (x) fm_free(x);
        if (e) fm_free(e);
        return rc;
    });
    if (rc != FM_OK) return rc;
    if (a_on)
        for (int r = 0; r < g.ndev; ++r)
            if (r != home) FM_CUDA(cudaStreamWaitEvent(ctx().stream, g.ev_done[r], 0));
    return FM_OK;
}

// One matrix exponential over the whole group ("single large expm": every product column-sharded and all-gathered, the
// solve's right-hand sides sharded, the LU replicated): see expm.cu k_expm_sharded.  a and out: both host memory, or both on the
// calling thread's device.  Blocks until the result is in place.  *info = LAPACK info of the Pade denominator's LU (0 in practice).
int fm_mg_expm(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *info) {
    if (!g_group) FM_TRY(fm_mg_init(0));
    Group &g = *g_group;
    if (n < 0 || lda < (n > 1 ? n : 1) || ldo < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "fm_mg_expm: bad dimensions");
    if (info) *info = 0;
    if (n == 0) return FM_OK;
    int a_dev = -1, o_dev = -1;
    const bool a_on = is_device_pointer(a, &a_dev), o_on = is_device_pointer(out, &o_dev);
    if (a_on != o_on) return set_error(FM_ERR_ARG, "fm_mg_expm: input and output must be both host or both device memory");
    FM_TRY(ensure_init());
    if (a_on && (a_dev != o_dev || ctx().device != a_dev))
        return set_error(FM_ERR_ARG, "fm_mg_expm: device operands must live on the calling thread's device (%d)", ctx().device);
    const int home = rank_of_device(g, ctx().device);
    if (home < 0) return set_error(FM_ERR_ARG, "fm_mg_expm: the calling thread's device %d is not part of the group", ctx().device);
    auto one_device = [&]() -> int {
        if (a_on) {
            FM_TRY(fm_expm(n, a, lda, out, ldo, mode, s_out, info));
            return fm_sync();
        }
        double *da = static_cast<double *>(fm_alloc_bytes((size_t)n * n * 8));
        if (!da) return FM_ERR_NOMEM;
        int rc = fm_upload(da, n, a, lda, n, n);
        if (rc == FM_OK) rc = fm_expm(n, da, n, out, ldo, mode, s_out, info);  // (host result buffer: downloads behind the last squaring)
        if (rc == FM_OK) rc = fm_sync();
        fm_free(da);
        return rc;
    };
    if (g.ndev == 1 || n < 128 * g.ndev) return one_device();  // nothing to shard (a column block is at least one 128-column tile column)
    if (a_on) {
        if (!g.ev_home || g.ev_home_dev != a_dev) {
            if (g.ev_home) cudaEventDestroy(g.ev_home);
            FM_CUDA(cudaEventCreateWithFlags(&g.ev_home, cudaEventDisableTiming));
            g.ev_home_dev = a_dev;
        }
        FM_CUDA(cudaEventRecord(g.ev_home, ctx().stream));
    }
    const int lddx = n + (n & 1);
    int32_t flags[MG_MAX][2] = {};
    double s_of[MG_MAX] = {};
    const int rc = g.workers.run([&](int r) -> int {
        int rc = FM_OK;
        cudaStream_t st = ctx().stream;
        int j0, nloc;
        column_block(n, g.ndev, r, &j0, &nloc);
        unsigned nbar = 0;
        auto barrier = [&](int rc_in) -> int {
            int rc = rc_in;
            const int slot = (int)(nbar++ & 1u);
            MG_CUDA(rc, cudaEventRecord(g.ev_bar[r][slot], st));
            rc = g.workers.agree(rc);  // every record of this barrier precedes every wait on it; two slots: nobody re-records an event a peer has yet to wait on
            if (rc != FM_OK) return rc;
            for (int p = 0; p < g.ndev; ++p)
                if (p != r) MG_CUDA(rc, cudaStreamWaitEvent(st, g.ev_bar[p][slot], 0));
            return rc;
        };
        // ---- a full copy of A on every device ----
        const bool is_home = a_on && r == home;
        g.da[r] = is_home ? nullptr : static_cast<double *>(fm_alloc_bytes((size_t)lddx * n * 8));
        if (!is_home && !g.da[r]) rc = FM_ERR_NOMEM;
        if (rc == FM_OK) rc = k_expm_workspace(n, g.ws[r]);
        if (a_on) MG_CUDA(rc, cudaStreamWaitEvent(st, g.ev_home, 0));
        // own column slice first (from the host over this device's PCIe link, or from the home device), the rest from the peers
        if (rc == FM_OK && !is_home && nloc > 0) {
            if (a_on) rc = pull2d(g.da[r] + (size_t)j0 * lddx, lddx, a + (size_t)j0 * lda, lda, a_dev, n, nloc, st);
            else rc = copy_h2d(g.da[r] + (size_t)j0 * lddx, lddx, a + (size_t)j0 * lda, lda, n, nloc, st);
        }
        rc = barrier(rc);  // every slice is in place and every g.da[] / g.ws[] is known
        if (rc != FM_OK) { if (g.da[r]) fm_free(g.da[r]); g.da[r] = nullptr; return rc; }
        if (!is_home) {
            for (int i = 1; i < g.ndev && rc == FM_OK; ++i) {
                const int p = (r + i) % g.ndev;
                int pj0, pn;
                column_block(n, g.ndev, p, &pj0, &pn);
                if (pn == 0) continue;
                if (a_on && p == home) rc = pull2d(g.da[r] + (size_t)pj0 * lddx, lddx, a + (size_t)pj0 * lda, lda, a_dev, n, pn, st);
                else MG_CUDA(rc, cudaMemcpyPeerAsync(g.da[r] + (size_t)pj0 * lddx, g.dev[r], g.da[p] + (size_t)pj0 * lddx, g.dev[p], (size_t)lddx * pn * 8, st));
            }
        }
        ExpmShard sh;
        sh.rank = r;
        sh.world = g.ndev;
        sh.j0 = j0;
        sh.nloc = nloc;
        sh.device = g.dev;
        sh.ws = g.ws;
        sh.barrier = barrier;
        // (k_expm_sharded opens with a barrier after the norm: nobody frees or overwrites its copy of A before every peer has read it)
        double *dout = nullptr;  // device result: straight into `out` on the home device; host result: every device returns its column block
        if (a_on) dout = is_home ? out : nullptr;
        const double *src = is_home ? a : g.da[r];
        const int lds = is_home ? lda : lddx;
        int where = -1;
        if (rc != FM_OK) { barrier(rc); }
        else rc = k_expm_sharded(n, src, lds, dout, ldo, mode, &s_of[r], flags[r], sh, &where);
        // host result: every device delivers its own column block over its own PCIe link
        if (rc == FM_OK && !a_on && nloc > 0) rc = copy_d2h(out + (size_t)j0 * ldo, ldo, g.ws[r][where] + (size_t)j0 * lddx, lddx, n, nloc, st);
        MG_CUDA(rc, cudaStreamSynchronize(st));
        if (g.da[r]) fm_free(g.da[r]);
        g.da[r] = nullptr;
        return rc;
    });
    if (rc != FM_OK) return rc;
    if (s_out) *s_out = s_of[home];
    bool redo = false;
    for (int r = 0; r < g.ndev; ++r) redo = redo || flags[r][0] != 0 || flags[r][1] != 0;
    // singular or not column-dominant denominator: the one-device pipeline knows what LAPACK / the reference do then (and the input
    // is untouched: the sharded pass only wrote workspaces and `out`)
    if (redo) return one_device();
    return FM_OK;
}

int fm_mg_sync(void) {
    if (!g_group) return fm_sync();
    Group &g = *g_group;
    const int rc = g.workers.run([&](int) -> int {
        FM_CUDA(cudaStreamSynchronize(ctx().stream));
        return FM_OK;
    });
    if (rc != FM_OK) return rc;
    return fm_sync();
}

}  // extern "C"
