// xfer.cu -- host <-> HBM copies for PAGEABLE host memory, the kind the reference hands over: flomat allocates every matrix with
// the garbage collector's `malloc` (flomat.rkt:437-441), so the operands of cblas_dgemm / dgesv_ / dgetrf_ / dgetri_ / dlange_
// arrive unpinned.  cudaMemcpy from pageable memory is staged by the driver on one thread (measured here: ~12.5 GB/s, 120 ms for
// the 1.5 GB of an n = 8192 product against 31 ms of arithmetic).  This file does the staging itself: a ring of pinned bounce
// buffers, filled (or drained) by a small pool of host threads while the DMA of the previous buffer is in flight, so the copy
// runs at the rate several cores can read memory instead of one.  Pinned (or registered) host memory skips all of this and goes
// straight to cudaMemcpy2DAsync.  No arithmetic happens here.
#include "common.cuh"
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace fm {
namespace {

constexpr int MAX_BOUNCE = 8;
// ring geometry; FLOMAT_CUDA_BOUNCE_MB / FLOMAT_CUDA_BOUNCE_COUNT are developer switches for sweeps (tools/host_gemm.py)
const size_t BOUNCE_BYTES = (size_t)(getenv("FLOMAT_CUDA_BOUNCE_MB") && atoi(getenv("FLOMAT_CUDA_BOUNCE_MB")) > 0 ? atoi(getenv("FLOMAT_CUDA_BOUNCE_MB")) : 32) << 20;
const int NBOUNCE = [] {
    const int c = getenv("FLOMAT_CUDA_BOUNCE_COUNT") ? atoi(getenv("FLOMAT_CUDA_BOUNCE_COUNT")) : 4;
    return c < 2 ? 2 : (c > MAX_BOUNCE ? MAX_BOUNCE : c);
}();
constexpr size_t MIN_STAGED_BYTES = (size_t)4 << 20;  // below this the driver's own path is as good

// persistent workers; the calling thread takes part as worker 0.  Leaked on purpose: the library may be unloaded at exit while a
// static destructor would have to join threads that are parked on the condition variable.
class CopyPool {
  public:
    explicit CopyPool(int nthreads) : n_(nthreads < 1 ? 1 : nthreads) {
        for (int i = 1; i < n_; ++i) std::thread([this] { worker(); }).detach();
    }
    int size() const { return n_; }
    // runs f(part) for part = 0 .. parts-1 and returns when all are done
    void run(int parts, const std::function<void(int)> &f) {
        if (parts <= 0) return;
        if (n_ == 1 || parts == 1) {
            for (int p = 0; p < parts; ++p) f(p);
            return;
        }
        std::lock_guard<std::mutex> one_job(run_mu_);  // one job slot: callers on different devices (fm_mg_* worker threads) take turns
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &f;
            parts_ = parts;
            next_.store(0);
            left_ = parts;
            ++gen_;
        }
        cv_.notify_all();
        drain(f, parts);
        std::unique_lock<std::mutex> lk(mu_);
        // also wait for workers that picked this job up but found nothing left: they must not touch next_ once the next job is set
        done_.wait(lk, [this] { return left_ == 0 && active_ == 0; });
        job_ = nullptr;
    }

  private:
    void drain(const std::function<void(int)> &f, int parts) {
        int mine = 0;
        for (;;) {
            const int p = next_.fetch_add(1);
            if (p >= parts) break;
            f(p);
            ++mine;
        }
        if (mine) {
            std::lock_guard<std::mutex> lk(mu_);
            left_ -= mine;
            if (left_ == 0) done_.notify_all();
        }
    }
    void worker() {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *f;
            int parts;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen && job_ != nullptr; });
                seen = gen_;
                f = job_;
                parts = parts_;
                ++active_;
            }
            drain(*f, parts);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--active_ == 0) done_.notify_all();
            }
        }
    }
    const int n_;
    std::mutex run_mu_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    int parts_ = 0, left_ = 0, active_ = 0;
    std::atomic<int> next_{0};
    unsigned long long gen_ = 0;
};

CopyPool &pool() {
    static CopyPool *p = [] {
        int t = 0;
        if (const char *e = getenv("FLOMAT_CUDA_COPY_THREADS")) t = atoi(e);
        if (t <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            t = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
        }
        return new CopyPool(t > 16 ? 16 : t);
    }();
    return *p;
}

struct Bounce {  // one ring per device: its events belong to that device and its buffers may be in flight on that device's streams
    double *buf[MAX_BOUNCE] = {};
    cudaEvent_t ev[MAX_BOUNCE] = {};
    bool busy[MAX_BOUNCE] = {};
    int next = 0;  // keeps rotating across calls: back-to-back uploads of small blocks do not all wait on buffer 0
    bool ready = false;
};
PerDevice<Bounce> g_bounce;

int bounce_init() {
    Bounce &g_b = g_bounce();
    if (g_b.ready) return FM_OK;
    for (int i = 0; i < NBOUNCE; ++i) {
        if (!g_b.buf[i]) FM_CUDA(cudaMallocHost(&g_b.buf[i], BOUNCE_BYTES));
        if (!g_b.ev[i]) FM_CUDA(cudaEventCreateWithFlags(&g_b.ev[i], cudaEventDisableTiming));
        g_b.busy[i] = false;
    }
    g_b.ready = true;
    return FM_OK;
}

// columns [c0, c0 + nc) of a host matrix <-> a packed m x nc bounce buffer, split over the pool
void host_pack(double *packed, const double *host, size_t ldh, size_t m, size_t c0, size_t nc, bool to_packed) {
    CopyPool &p = pool();
    if (ldh == m) {  // one contiguous range: equal byte slices
        const size_t total = m * nc * sizeof(double);
        char *a = reinterpret_cast<char *>(packed);
        char *h = reinterpret_cast<char *>(const_cast<double *>(host) + c0 * ldh);
        const int parts = p.size();
        const size_t slice = ((total + parts - 1) / parts + 4095) & ~(size_t)4095;
        p.run(parts, [&](int i) {
            const size_t lo = (size_t)i * slice;
            if (lo >= total) return;
            const size_t len = total - lo < slice ? total - lo : slice;
            if (to_packed) memcpy(a + lo, h + lo, len);
            else memcpy(h + lo, a + lo, len);
        });
        return;
    }
    const int parts = (int)(nc < (size_t)(4 * p.size()) ? nc : (size_t)(4 * p.size()));
    p.run(parts, [&](int i) {
        const size_t lo = nc * i / parts, hi = nc * (i + 1) / parts;
        for (size_t j = lo; j < hi; ++j) {
            double *h = const_cast<double *>(host) + (c0 + j) * ldh;
            if (to_packed) memcpy(packed + j * m, h, m * sizeof(double));
            else memcpy(h, packed + j * m, m * sizeof(double));
        }
    });
}

bool use_staging(const void *host, size_t m, size_t n) {
    static const bool off = getenv("FLOMAT_CUDA_NO_STAGING") != nullptr;  // developer switch: the driver's pageable path
    if (off || m * n * sizeof(double) < MIN_STAGED_BYTES || m * sizeof(double) > BOUNCE_BYTES) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

}  // namespace

// dev(m x n, ldd) := host(m x n, ldh), queued on `st`.  Pinned host memory: asynchronous.  Pageable: returns once the host matrix
// has been read (the last DMA may still be in flight on `st`).
int copy_h2d(double *dev, size_t ldd, const double *host, size_t ldh, size_t m, size_t n, cudaStream_t st) {
    if (m == 0 || n == 0) return FM_OK;
    if (!use_staging(host, m, n)) {
        FM_CUDA(cudaMemcpy2DAsync(dev, ldd * 8, host, ldh * 8, m * 8, n, cudaMemcpyHostToDevice, st));
        return FM_OK;
    }
    FM_TRY(bounce_init());
    Bounce &g_b = g_bounce();
    const size_t per = BOUNCE_BYTES / (m * sizeof(double));
    int b = g_b.next;
    for (size_t c0 = 0; c0 < n; c0 += per, b = (b + 1) % NBOUNCE, g_b.next = b) {
        const size_t nc = n - c0 < per ? n - c0 : per;
        if (g_b.busy[b]) FM_CUDA(cudaEventSynchronize(g_b.ev[b]));
        host_pack(g_b.buf[b], host, ldh, m, c0, nc, true);
        FM_CUDA(cudaMemcpy2DAsync(dev + c0 * ldd, ldd * 8, g_b.buf[b], m * 8, m * 8, nc, cudaMemcpyHostToDevice, st));
        FM_CUDA(cudaEventRecord(g_b.ev[b], st));
        g_b.busy[b] = true;
    }
    return FM_OK;
}

// host(m x n, ldh) := dev(m x n, ldd) after everything queued on `st`.  Pinned host memory: asynchronous (valid after the stream
// has been synchronised).  Pageable: returns when the data is in place.
int copy_d2h(double *host, size_t ldh, const double *dev, size_t ldd, size_t m, size_t n, cudaStream_t st) {
    if (m == 0 || n == 0) return FM_OK;
    if (!use_staging(host, m, n)) {
        FM_CUDA(cudaMemcpy2DAsync(host, ldh * 8, dev, ldd * 8, m * 8, n, cudaMemcpyDeviceToHost, st));
        return FM_OK;
    }
    FM_TRY(bounce_init());
    Bounce &g_b = g_bounce();
    const size_t per = BOUNCE_BYTES / (m * sizeof(double));
    const size_t nchunks = (n + per - 1) / per;
    auto issue = [&](size_t i) -> int {
        const int b = (int)(i % NBOUNCE);
        const size_t c0 = i * per, nc = n - c0 < per ? n - c0 : per;
        if (g_b.busy[b]) FM_CUDA(cudaEventSynchronize(g_b.ev[b]));  // an earlier upload out of this buffer
        FM_CUDA(cudaMemcpy2DAsync(g_b.buf[b], m * 8, dev + c0 * ldd, ldd * 8, m * 8, nc, cudaMemcpyDeviceToHost, st));
        FM_CUDA(cudaEventRecord(g_b.ev[b], st));
        g_b.busy[b] = true;
        return FM_OK;
    };
    for (size_t i = 0; i < nchunks && i < (size_t)NBOUNCE; ++i) FM_TRY(issue(i));
    for (size_t i = 0; i < nchunks; ++i) {
        const int b = (int)(i % NBOUNCE);
        const size_t c0 = i * per, nc = n - c0 < per ? n - c0 : per;
        FM_CUDA(cudaEventSynchronize(g_b.ev[b]));
        g_b.busy[b] = false;
        host_pack(g_b.buf[b], host, ldh, m, c0, nc, false);
        if (i + NBOUNCE < nchunks) FM_TRY(issue(i + NBOUNCE));
    }
    return FM_OK;
}

}  // namespace fm
