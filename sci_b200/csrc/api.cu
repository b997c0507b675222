// api.cu -- the C ABI of libflomat_cuda (include/flomat_cuda.h): context, memory, host<->device sync points and the
// compat tier that carries the exact symbol names/signatures flomat.rkt binds (cblas_dgemm :855, cblas_daxpy :751,
// cblas_dcopy :483, cblas_dscal :1014, dlange_ :1317, dgetrf_ :1505, dgesv_ :2112, dgetri_ :1754).
// Host operands of the compat tier are staged through HBM inside the call; device operands run in place.
// Nothing here computes on the CPU: every arithmetic result comes from a kernel in this library.
#include "common.cuh"
#include <map>
#include <unordered_map>
#include <cstdlib>
#include <vector>
#include <utility>
#include <cmath>
#include <mutex>

namespace fm {

static thread_local char g_err[512] = "";
static Context g_ctxs[FM_MAX_DEVICES];
static thread_local int t_dev = -1;  // the device this host thread is bound to
static int g_default_dev = -1;       // the first device any thread bound: the home of threads that never called fm_init
static std::mutex g_init_mu;

Context &ctx() { return g_ctxs[t_dev >= 0 ? t_dev : (g_default_dev >= 0 ? g_default_dev : 0)]; }
int bound_device() { return t_dev; }

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (getenv("FLOMAT_CUDA_DEBUG")) fprintf(stderr, "[libflomat_cuda] error %d: %s\n", code, g_err);
    return code;
}

static int init_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(FM_ERR_NOGPU, "no CUDA device available (%s); libflomat_cuda has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count || device >= FM_MAX_DEVICES)
        return set_error(FM_ERR_ARG, "fm_init: device %d out of range (0..%d)", device, (count < FM_MAX_DEVICES ? count : FM_MAX_DEVICES) - 1);
    FM_CUDA(cudaSetDevice(device));  // (the CUDA current device is per host thread as well)
    std::lock_guard<std::mutex> lk(g_init_mu);
    Context &c = g_ctxs[device];
    if (!c.ready) {
        cudaDeviceProp prop;
        FM_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return set_error(FM_ERR_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        c.device = device;
        c.sm_count = prop.multiProcessorCount;
        c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
        c.scratch = nullptr;
        c.scratch_bytes = 0;
        FM_CUDA(cudaMalloc(&c.d_scalar, 64 * sizeof(double)));
        FM_CUDA(cudaMallocHost(&c.h_scalar, 64 * sizeof(double)));
        c.ready = true;
    }
    t_dev = device;
    if (g_default_dev < 0) g_default_dev = device;
    return FM_OK;
}

int bind_device(int device) { return init_device(device); }
unsigned long long launches_of_device(int d) { return (d >= 0 && d < FM_MAX_DEVICES) ? g_ctxs[d].launches : 0ull; }
const char *error_text() { return g_err; }

int ensure_init() {
    if (t_dev >= 0) return FM_OK;  // bound threads have a ready context
    if (g_default_dev >= 0) return init_device(g_default_dev);
    const char *dev = getenv("FLOMAT_CUDA_DEVICE");
    return init_device(dev ? atoi(dev) : 0);
}

void *scratch(size_t bytes) {
    Context &c = ctx();
    if (bytes <= c.scratch_bytes) return c.scratch;
    if (c.scratch) cudaFree(c.scratch);  // implicit device sync: nothing in flight still uses it
    c.scratch = nullptr;
    c.scratch_bytes = 0;
    size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
    if (cudaMalloc(&c.scratch, want) != cudaSuccess) {
        cudaGetLastError();
        set_error(FM_ERR_NOMEM, "scratch allocation of %zu bytes failed", want);
        return nullptr;
    }
    c.scratch_bytes = want;
    return c.scratch;
}

// ---- pointer classification + staging for the compat tier --------------------------------------
bool is_device_pointer(const void *p, int *device_out) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (device_out) *device_out = at.device;
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
static bool is_device_ptr(const void *p) { return is_device_pointer(p, nullptr); }

// A host matrix staged into HBM for the duration of one compat call.
struct Staged {
    double *dev = nullptr;
    int ld = 0;
    bool owned = false;
    double *host = nullptr;
    int ldh = 0, m = 0, n = 0;
    int in(const double *p, int ldp, int rows, int cols, bool copy_in) {
        m = rows; n = cols;
        if (is_device_ptr(p) || rows == 0 || cols == 0) { dev = const_cast<double *>(p); ld = ldp; owned = false; return FM_OK; }
        host = const_cast<double *>(p); ldh = ldp;
        ld = rows + (rows & 1);  // even leading dimension keeps the TMA/128-bit paths available
        if (ld < 2) ld = 2;
        if (!(dev = static_cast<double *>(fm_alloc_bytes((size_t)ld * cols * sizeof(double))))) {
            cudaGetLastError();
            return set_error(FM_ERR_NOMEM, "staging allocation of %zu bytes failed", (size_t)ld * cols * sizeof(double));
        }
        owned = true;
        if (copy_in)
            FM_TRY(copy_h2d(dev, ld, host, ldh, rows, cols, ctx().stream));
        return FM_OK;
    }
    int out() {
        if (!owned) return FM_OK;
        FM_TRY(copy_d2h(host, ldh, dev, ld, m, n, ctx().stream));
        return FM_OK;
    }
    ~Staged() { if (owned && dev) { cudaStreamSynchronize(ctx().stream); fm_free(dev); } }
};

static void report(const char *who, int rc) {
    if (rc != FM_OK) fprintf(stderr, "[libflomat_cuda] %s failed (%d): %s\n", who, rc, g_err);
}

}  // namespace fm

using namespace fm;

extern "C" {

// =================================== library / device ===================================
int fm_init(int device) { return init_device(device); }
int fm_device_count(void) { int c = 0; if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; } return c; }
int fm_set_stream(void *s) {
    FM_TRY(ensure_init());
    if (ctx().stream != static_cast<cudaStream_t>(s)) cudaStreamSynchronize(ctx().stream);  // cached blocks are reused in stream order
    ctx().stream = static_cast<cudaStream_t>(s);
    return FM_OK;
}
int fm_transfer_sync(void);
int fm_sync(void) {
    FM_TRY(ensure_init());
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    return fm_transfer_sync();
}
const char *fm_last_error(void) { return g_err; }
const char *fm_version(void) { return "libflomat_cuda 0.1 (sm_100a)"; }
uint64_t fm_kernel_launches(void) {  // all devices of the process (fm_mg_* launch from worker threads on every device)
    uint64_t total = 0;
    for (int d = 0; d < FM_MAX_DEVICES; ++d) total += fm::launches_of_device(d);
    return total;
}

// =================================== memory ===================================
// fm_alloc / fm_free sit where the reference has GC'd `malloc` (flomat.rkt:437-441): cheap and called for every
// intermediate matrix.  cudaMalloc/cudaFree cost milliseconds at these sizes and synchronise the device, so freed blocks
// are kept in an exact-size cache and handed out again; reuse is safe without a sync because all work of the library is
// ordered on one stream.  FLOMAT_CUDA_CACHE_MB bounds the cache (default 64 GiB); fm_trim_cache() empties it.
namespace {
// Asynchronous host<->device transfers run on their own stream so that an upload can overlap the kernels queued before it and a
// download the kernels queued after it (fm_upload_async / fm_download_async / fm_transfer_sync).
struct Xfer {
    cudaStream_t stream = nullptr;
    cudaStream_t down = nullptr;  // device->host chunks of fm_dgemm_host: a second copy engine beside the uploads on `stream`
    cudaEvent_t last = nullptr, last_down = nullptr, tmp = nullptr;  // last upload / last download issued; scratch
    bool dirty = false;  // transfers issued since the last fm_transfer_sync
};
PerDevice<Xfer> g_xfer_pd;
Xfer &xfer() { return g_xfer_pd(); }
struct Block {
    size_t bytes = 0;
    cudaEvent_t freed = nullptr;  // recorded on the compute stream when the block was last given back (only once transfers are in use)
    bool fresh = true;            // straight from cudaMalloc: no queued work can refer to it
};
struct BlockCache {
    std::multimap<size_t, std::pair<void *, Block>> free_blocks;
    std::unordered_map<void *, Block> live;
    size_t cached_bytes = 0;
    size_t limit() {
        static size_t lim = [] {
            const char *e = getenv("FLOMAT_CUDA_CACHE_MB");
            return (e ? (size_t)atoll(e) : (size_t)65536) << 20;
        }();
        return lim;
    }
    void trim() {
        if (free_blocks.empty()) return;
        cudaStreamSynchronize(ctx().stream);
        if (xfer().stream) cudaStreamSynchronize(xfer().stream);
        if (xfer().down) cudaStreamSynchronize(xfer().down);
        for (auto &kv : free_blocks) {
            if (kv.second.second.freed) cudaEventDestroy(kv.second.second.freed);
            cudaFree(kv.second.first);
        }
        free_blocks.clear();
        cached_bytes = 0;
    }
    Block *find(const void *p) {  // the live block containing p (views point into the middle of a block)
        auto it = live.find(const_cast<void *>(p));
        if (it != live.end()) return &it->second;
        for (auto &kv : live)
            if (p >= kv.first && p < static_cast<const char *>(kv.first) + kv.second.bytes) return &kv.second;
        return nullptr;
    }
};
PerDevice<BlockCache> g_cache_pd;
BlockCache &cache() { return g_cache_pd(); }
int xfer_init() {
    if (xfer().stream) return FM_OK;
    FM_CUDA(cudaStreamCreateWithFlags(&xfer().stream, cudaStreamNonBlocking));
    FM_CUDA(cudaStreamCreateWithFlags(&xfer().down, cudaStreamNonBlocking));
    FM_CUDA(cudaEventCreateWithFlags(&xfer().last, cudaEventDisableTiming));
    FM_CUDA(cudaEventCreateWithFlags(&xfer().last_down, cudaEventDisableTiming));
    FM_CUDA(cudaEventCreateWithFlags(&xfer().tmp, cudaEventDisableTiming));
    return FM_OK;
}
}  // namespace

void *fm_alloc_bytes(size_t bytes) {
    if (ensure_init() != FM_OK) return nullptr;
    if (bytes == 0) bytes = 16;
    bytes = (bytes + 511) & ~(size_t)511;
    auto it = cache().free_blocks.find(bytes);
    if (it != cache().free_blocks.end()) {
        void *p = it->second.first;
        Block blk = it->second.second;
        cache().free_blocks.erase(it);
        cache().cached_bytes -= bytes;
        blk.fresh = false;
        cache().live[p] = blk;
        return p;
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        cache().trim();  // give the cached blocks back and try once more
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            set_error(FM_ERR_NOMEM, "cudaMalloc of %zu bytes failed", bytes);
            return nullptr;
        }
    }
    Block blk;
    blk.bytes = bytes;
    cache().live[p] = blk;
    return p;
}
double *fm_alloc(int m, int n) {
    if (m < 0 || n < 0) { set_error(FM_ERR_ARG, "fm_alloc: negative dimension"); return nullptr; }
    return static_cast<double *>(fm_alloc_bytes((size_t)m * (size_t)n * sizeof(double)));
}
int fm_free(void *dev) {
    if (!dev) return FM_OK;
    auto it = cache().live.find(dev);
    if (it == cache().live.end()) {
        // a block of another device's cache (freed from a thread bound elsewhere): give it straight back to the driver
        for (int d = 0; d < FM_MAX_DEVICES; ++d) {
            BlockCache &other = g_cache_pd.v[d];
            auto jt = other.live.find(dev);
            if (&other == &cache() || jt == other.live.end()) continue;
            if (jt->second.freed) cudaEventDestroy(jt->second.freed);
            other.live.erase(jt);
            break;
        }
        FM_CUDA(cudaFree(dev));  // not ours at all (or freed twice): plain cudaFree reports the error
        return FM_OK;
    }
    Block blk = it->second;
    cache().live.erase(it);
    if (xfer().stream) {
        // a transfer may still be reading or writing the block: whatever reuses it on the compute stream comes after the transfers
        if (xfer().dirty) {
            FM_CUDA(cudaStreamWaitEvent(ctx().stream, xfer().last, 0));
            FM_CUDA(cudaStreamWaitEvent(ctx().stream, xfer().last_down, 0));
        }
        // and a later asynchronous upload into the recycled block must wait for the kernels queued up to now
        if (!blk.freed) FM_CUDA(cudaEventCreateWithFlags(&blk.freed, cudaEventDisableTiming));
        FM_CUDA(cudaEventRecord(blk.freed, ctx().stream));
    }
    if (cache().cached_bytes + blk.bytes > cache().limit()) {
        if (blk.freed) cudaEventDestroy(blk.freed);
        FM_CUDA(cudaFree(dev));
        return FM_OK;
    }
    cache().free_blocks.emplace(blk.bytes, std::make_pair(dev, blk));
    cache().cached_bytes += blk.bytes;
    return FM_OK;
}
int fm_trim_cache(void) {
    FM_TRY(ensure_init());
    cache().trim();
    return FM_OK;
}
int fm_upload(double *dev, int ldd, const double *host, int ldh, int m, int n) {
    FM_TRY(ensure_init());
    if (m < 0 || n < 0 || ldd < m || ldh < m) return set_error(FM_ERR_ARG, "fm_upload: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    return copy_h2d(dev, ldd, host, ldh, m, n, ctx().stream);  // pageable host memory is staged by the library's own copy threads
}
int fm_download(double *host, int ldh, const double *dev, int ldd, int m, int n) {
    FM_TRY(ensure_init());
    if (m < 0 || n < 0 || ldd < m || ldh < m) return set_error(FM_ERR_ARG, "fm_download: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    FM_TRY(copy_d2h(host, ldh, dev, ldd, m, n, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    return FM_OK;
}
// Asynchronous transfers on the library's transfer stream.  Upload: the copy may run while kernels queued BEFORE the call are still
// executing (the destination must not be in use by them, other than through the allocator's recycling, which is tracked); everything
// queued on the compute stream AFTER the call sees the data.  Download: waits for the work queued before the call, then runs beside
// whatever is queued after it; the host buffer is valid after fm_transfer_sync (or fm_sync).  Pinned host memory is needed for a
// real overlap; pageable memory works but serialises inside the driver.
int fm_upload_async(double *dev, int ldd, const double *host, int ldh, int m, int n) {
    FM_TRY(ensure_init());
    if (m < 0 || n < 0 || ldd < m || ldh < m) return set_error(FM_ERR_ARG, "fm_upload_async: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    FM_TRY(xfer_init());
    Block *blk = cache().find(dev);
    if (blk && blk->freed) {
        FM_CUDA(cudaStreamWaitEvent(xfer().stream, blk->freed, 0));
    } else if (!blk || !blk->fresh) {  // unknown history: behind everything queued so far
        FM_CUDA(cudaEventRecord(xfer().tmp, ctx().stream));
        FM_CUDA(cudaStreamWaitEvent(xfer().stream, xfer().tmp, 0));
    }
    FM_CUDA(cudaMemcpy2DAsync(dev, (size_t)ldd * 8, host, (size_t)ldh * 8, (size_t)m * 8, n, cudaMemcpyHostToDevice, xfer().stream));
    FM_CUDA(cudaEventRecord(xfer().last, xfer().stream));
    FM_CUDA(cudaStreamWaitEvent(ctx().stream, xfer().last, 0));
    xfer().dirty = true;
    return FM_OK;
}
int fm_download_async(double *host, int ldh, const double *dev, int ldd, int m, int n) {
    FM_TRY(ensure_init());
    if (m < 0 || n < 0 || ldd < m || ldh < m) return set_error(FM_ERR_ARG, "fm_download_async: bad dimensions");
    if (m == 0 || n == 0) return FM_OK;
    FM_TRY(xfer_init());
    // on the DOWNLOAD stream: a download that waits for a kernel must not hold back the uploads queued after it (second copy engine)
    FM_CUDA(cudaEventRecord(xfer().tmp, ctx().stream));
    FM_CUDA(cudaStreamWaitEvent(xfer().down, xfer().tmp, 0));
    FM_CUDA(cudaMemcpy2DAsync(host, (size_t)ldh * 8, dev, (size_t)ldd * 8, (size_t)m * 8, n, cudaMemcpyDeviceToHost, xfer().down));
    FM_CUDA(cudaEventRecord(xfer().last_down, xfer().down));
    xfer().dirty = true;
    return FM_OK;
}
}  // extern "C"
namespace fm {
int download_behind_compute(double *host, size_t ldh, const double *dev, size_t ldd, size_t m, size_t n, bool last) {
    FM_TRY(xfer_init());
    FM_CUDA(cudaEventRecord(xfer().tmp, ctx().stream));
    FM_CUDA(cudaStreamWaitEvent(xfer().down, xfer().tmp, 0));
    FM_TRY(copy_d2h(host, ldh, dev, ldd, m, n, xfer().down));
    if (last) {
        FM_CUDA(cudaEventRecord(xfer().tmp, xfer().down));
        FM_CUDA(cudaStreamWaitEvent(ctx().stream, xfer().tmp, 0));
    }
    xfer().dirty = true;
    return FM_OK;
}
}  // namespace fm
extern "C" {
int fm_transfer_sync(void) {
    FM_TRY(ensure_init());
    if (xfer().stream) FM_CUDA(cudaStreamSynchronize(xfer().stream));
    if (xfer().down) FM_CUDA(cudaStreamSynchronize(xfer().down));
    xfer().dirty = false;
    return FM_OK;
}
int fm_get(const double *dev, size_t index, double *out) {
    FM_TRY(ensure_init());
    FM_CUDA(cudaMemcpyAsync(out, dev + index, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    return FM_OK;
}
int fm_set(double *dev, size_t index, double value) {
    FM_TRY(ensure_init());
    FM_CUDA(cudaMemcpyAsync(dev + index, &value, sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    return FM_OK;
}
int fm_upload_i32(int32_t *dev, const int32_t *host, size_t n) {
    FM_TRY(ensure_init());
    FM_CUDA(cudaMemcpyAsync(dev, host, n * 4, cudaMemcpyHostToDevice, ctx().stream));
    return FM_OK;
}
int fm_download_i32(int32_t *host, const int32_t *dev, size_t n) {
    FM_TRY(ensure_init());
    FM_CUDA(cudaMemcpyAsync(host, dev, n * 4, cudaMemcpyDeviceToHost, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    return FM_OK;
}

// =================================== elementwise ===================================
int fm_fill(double *a, int lda, int m, int n, double x) { FM_TRY(ensure_init()); return k_fill(a, lda, m, n, x); }
int fm_eye(double *a, int lda, int m, int n, int k) { FM_TRY(ensure_init()); return k_eye(a, lda, m, n, k); }
int fm_copy2d(double *dst, int ldd, const double *src, int lds, int m, int n) { FM_TRY(ensure_init()); return k_copy2d(dst, ldd, src, lds, m, n); }
int fm_ewise(int op, int m, int n, const double *a, int lda, double sa, const double *b, int ldb, double sb, double *c, int ldc) {
    FM_TRY(ensure_init());
    return k_ewise(op, m, n, a, lda, sa, b, ldb, sb, c, ldc);
}
int fm_axpy2d(int m, int n, double alpha, const double *a, int lda, double *b, int ldb) { FM_TRY(ensure_init()); return k_axpy2d(m, n, alpha, a, lda, b, ldb); }
int fm_scal2d(int m, int n, double s, double *a, int lda) { FM_TRY(ensure_init()); return k_scal2d(m, n, s, a, lda); }
int fm_addsub(int m, int n, const double *e, int lde, const double *o, int ldo, double *num, int ldn, double *den, int ldd) {
    FM_TRY(ensure_init());
    return k_addsub(m, n, e, lde, o, ldo, num, ldn, den, ldd);
}

int fm_norm(char type, int m, int n, const double *a, int lda, double *out) {
    FM_TRY(ensure_init());
    FM_TRY(k_norm_device(type, m, n, a, lda, ctx().d_scalar));
    FM_CUDA(cudaMemcpyAsync(ctx().h_scalar, ctx().d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = ctx().h_scalar[0];
    return FM_OK;
}

// =================================== gemm ===================================
int fm_dgemm(int transA, int transB, int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta,
             double *c, int ldc) {
    FM_TRY(ensure_init());
    GemmEpilogue ep;
    ep.alpha = alpha;
    ep.beta = beta;
    return k_dgemm(transA, transB, m, n, k, a, lda, 0, b, ldb, 0, c, ldc, 0, 1, ep);
}
// The column blocks fm_dgemm_host uses for an m x n result on `sms` SMs: width (a multiple of 128 columns) and count; the last block
// takes whatever is left (it may be up to half a block wider).  Pure host arithmetic, no device needed.
int fm_dgemm_host_plan(int m, int n, int sms, int *width, int *nblocks) {
    if (m <= 0 || n <= 0 || !width || !nblocks) return set_error(FM_ERR_ARG, "fm_dgemm_host_plan: bad arguments");
    if (sms <= 0) sms = 148;
    // about eight blocks: the narrowest one (shortest head and tail of the pipeline) whose 128 x 128 tiles fill >= 97 % of their
    // waves of the persistent GEMM, else the best there is
    const int tile_rows = (m + 127) / 128, nblk = (n + 127) / 128;
    int per = 1;
    const int lo = nblk / 8 > 1 ? nblk / 8 : 1, hi = nblk / 3 > lo ? nblk / 3 : lo;
    double best = -1.0;
    for (int cnd = lo; cnd <= hi; ++cnd) {
        const long long tiles = (long long)cnd * tile_rows, waves = (tiles + sms - 1) / sms;
        const double eff = (double)tiles / (double)(waves * sms);
        if (eff > best + 1e-9) { best = eff; per = cnd; }
        if (eff >= 0.97) { per = cnd; break; }
    }
    const int cw = per * 128;
    int nchunks = (n + cw - 1) / cw;
    if (nchunks > 1 && n - (nchunks - 1) * cw < cw / 2) --nchunks;  // a short remainder rides with the last block
    *width = cw;
    *nblocks = nchunks;
    return FM_OK;
}

// Which configuration of the DMMA GEMM an untransposed, aligned m x n x k product takes on `sms` SMs (dgemm.cu: gemm_tile_plan):
// *tile_rows = 128 or 64 (x 128 columns; 64: two CTAs per SM), *ksplit = number of k slices (1 = no split).  Pure host arithmetic.
int fm_dgemm_tile_plan(int m, int n, int k, int batch, int sms, int *tile_rows, int *ksplit) {
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0 || !tile_rows || !ksplit) return set_error(FM_ERR_ARG, "fm_dgemm_tile_plan: bad arguments");
    *tile_rows = gemm_tile_plan(m, n, k, batch, sms, true, ksplit) ? 64 : 128;
    return FM_OK;
}
// The column chunks in which dgesv_ / dgetrf_ move a HOST matrix of order n through the LU pipeline (getrf_host_pipeline): bounds[0] = 0
// < bounds[1] < ... < bounds[*nchunks] = n, every bound but the last a multiple of 128.  Pure host arithmetic.
int fm_lu_host_plan(int n, int *bounds, int max_bounds, int *nchunks) {
    if (n <= 0 || !bounds || !nchunks || max_bounds < 2) return set_error(FM_ERR_ARG, "fm_lu_host_plan: bad arguments");
    const std::vector<int> b = lu_host_bounds(n);
    if ((int)b.size() > max_bounds) return set_error(FM_ERR_ARG, "fm_lu_host_plan: %d bounds do not fit %d", (int)b.size(), max_bounds);
    for (size_t i = 0; i < b.size(); ++i) bounds[i] = b[i];
    *nchunks = (int)b.size() - 1;
    return FM_OK;
}

// C := alpha*A*B + beta*C with all three operands in HOST memory (what the unmodified reference passes to cblas_dgemm,
// flomat.rkt:888): instead of upload / compute / download one after the other, B and C travel in column blocks -- A goes up once,
// then block j of B is uploaded while the product of block j-1 runs and block j-2 of C comes down (uploads on the transfer stream,
// downloads on a second one: both copy engines and the SMs are busy at once).  Block widths are whole numbers of 128-column tile
// columns chosen so that a block's tiles fill complete waves of the persistent GEMM.  n = 8192 from pinned memory: upload + compute +
// download 60 ms one after the other, ~44 ms pipelined.  async != 0: return without waiting; C is valid after fm_transfer_sync().
int fm_dgemm_host(int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc,
                  int async) {
    FM_TRY(ensure_init());
    if (m < 0 || n < 0 || k < 0) return set_error(FM_ERR_ARG, "fm_dgemm_host: negative dimension");
    if (lda < (m > 1 ? m : 1) || ldb < (k > 1 ? k : 1) || ldc < (m > 1 ? m : 1)) return set_error(FM_ERR_ARG, "fm_dgemm_host: leading dimension too small");
    if (m == 0 || n == 0) return FM_OK;
    if (is_device_ptr(a) || is_device_ptr(b) || is_device_ptr(c)) return set_error(FM_ERR_ARG, "fm_dgemm_host: operands must be host memory (use fm_dgemm)");
    return dgemm_host_pipeline(m, n, k, alpha, a, lda, nullptr, 0, b, ldb, beta, c, ldc, async);
}
}  // extern "C"
namespace fm {
int xfer_streams(cudaStream_t *up, cudaStream_t *down) {
    FM_TRY(xfer_init());
    if (up) *up = xfer().stream;
    if (down) *down = xfer().down;
    xfer().dirty = true;  // the caller is about to use them: fm_free / fm_transfer_sync must take them into account
    return FM_OK;
}
// The column-block pipeline behind fm_dgemm_host.  a_dev != nullptr: A is already (or will be, by work queued on the UPLOAD stream)
// resident at a_dev with leading dimension ldda -- fm_mg_dgemm assembles it there over PCIe + NVLink -- and a_host is not read.
int dgemm_host_pipeline(int m, int n, int k, double alpha, const double *a, int lda, const double *a_dev, int ldda_in, const double *b, int ldb,
                        double beta, double *c, int ldc, int async) {
    FM_TRY(xfer_init());
    const int ldda = a_dev ? ldda_in : (m + (m & 1) < 2 ? 2 : m + (m & 1)), lddb = k + (k & 1) < 2 ? 2 : k + (k & 1);
    const int lddc = m + (m & 1) < 2 ? 2 : m + (m & 1);
    double *da = a_dev ? const_cast<double *>(a_dev) : (k > 0 ? static_cast<double *>(fm_alloc_bytes((size_t)ldda * k * 8)) : nullptr);
    double *db = k > 0 ? static_cast<double *>(fm_alloc_bytes((size_t)lddb * n * 8)) : nullptr;
    double *dc = static_cast<double *>(fm_alloc_bytes((size_t)lddc * n * 8));
    int rc = FM_OK;
    if ((k > 0 && (!da || !db)) || !dc) rc = FM_ERR_NOMEM;
    int cw = 0, nchunks = 0;
    fm_dgemm_host_plan(m, n, ctx().sm_count, &cw, &nchunks);
    static PerDevice<std::vector<cudaEvent_t>> ev_pd;  // [2j] block j uploaded, [2j+1] block j computed; events belong to their device
    std::vector<cudaEvent_t> &ev = ev_pd();
    while (rc == FM_OK && (int)ev.size() < 2 * nchunks) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "dgemm_host_pipeline: event creation failed");
        else ev.push_back(e);
    }
    auto chk = [&](cudaError_t e) { if (e != cudaSuccess && rc == FM_OK) rc = set_error(FM_ERR_CUDA, "fm_dgemm_host: %s", cudaGetErrorString(e)); };
    // developer switch: timestamps of every block's upload / product / download, printed after the call (tools/host_gemm_trace.py)
    static const bool trace_on = getenv("FLOMAT_HOST_GEMM_TRACE") != nullptr;
    const bool trace = trace_on && !async && rc == FM_OK;
    std::vector<cudaEvent_t> tr;  // [0] start, [1] A uploaded, then per block: uploaded, product started, product done, downloaded
    if (trace) {
        tr.resize(2 + 4 * (size_t)nchunks);
        for (auto &e : tr) cudaEventCreate(&e);
    }
    if (rc == FM_OK) {
        // the staging blocks may be recycled ones that kernels queued so far still use
        chk(cudaEventRecord(xfer().tmp, ctx().stream));
        chk(cudaStreamWaitEvent(xfer().stream, xfer().tmp, 0));
        if (trace) cudaEventRecord(tr[0], xfer().stream);
        if (k > 0 && rc == FM_OK && !a_dev) rc = copy_h2d(da, ldda, a, lda, m, k, xfer().stream);
        if (trace) cudaEventRecord(tr[1], xfer().stream);
        for (int j = 0; j < nchunks && rc == FM_OK; ++j) {
            const int c0 = j * cw, w = (j == nchunks - 1) ? n - c0 : cw;
            if (k > 0) rc = copy_h2d(db + (size_t)c0 * lddb, lddb, b + (size_t)c0 * ldb, ldb, k, w, xfer().stream);
            if (beta != 0.0 && rc == FM_OK) rc = copy_h2d(dc + (size_t)c0 * lddc, lddc, c + (size_t)c0 * ldc, ldc, m, w, xfer().stream);
            chk(cudaEventRecord(ev[2 * j], xfer().stream));
            chk(cudaStreamWaitEvent(ctx().stream, ev[2 * j], 0));
            if (rc != FM_OK) break;
            if (trace) { cudaEventRecord(tr[2 + 4 * j], xfer().stream); cudaEventRecord(tr[3 + 4 * j], ctx().stream); }
            GemmEpilogue ep;
            ep.alpha = alpha;
            ep.beta = beta;
            if (k > 0) rc = k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, w, k, da, ldda, 0, db + (size_t)c0 * lddb, lddb, 0, dc + (size_t)c0 * lddc, lddc, 0, 1, ep);
            else if (beta == 0.0) rc = k_fill(dc + (size_t)c0 * lddc, lddc, m, w, 0.0);  // BLAS: k = 0 leaves C := beta*C (beta = 0: no read of C)
            else rc = k_scal2d(m, w, beta, dc + (size_t)c0 * lddc, lddc);
            if (rc == FM_OK) chk(cudaEventRecord(ev[2 * j + 1], ctx().stream));
            if (trace) cudaEventRecord(tr[4 + 4 * j], ctx().stream);
        }
        // downloads are queued after every upload: a copy to pageable memory blocks the host until it has run
        for (int j = 0; j < nchunks && rc == FM_OK; ++j) {
            const int c0 = j * cw, w = (j == nchunks - 1) ? n - c0 : cw;
            chk(cudaStreamWaitEvent(xfer().down, ev[2 * j + 1], 0));
            if (rc == FM_OK) rc = copy_d2h(c + (size_t)c0 * ldc, ldc, dc + (size_t)c0 * lddc, lddc, m, w, xfer().down);
            if (trace) cudaEventRecord(tr[5 + 4 * j], xfer().down);
        }
        // whatever reuses the staging blocks (compute stream) comes after the last download
        chk(cudaEventRecord(xfer().tmp, xfer().down));
        chk(cudaStreamWaitEvent(ctx().stream, xfer().tmp, 0));
        chk(cudaEventRecord(xfer().last, xfer().stream));
        xfer().dirty = true;
    }
    if (rc != FM_OK || !async) {
        cudaStreamSynchronize(xfer().stream);
        cudaStreamSynchronize(ctx().stream);
        cudaStreamSynchronize(xfer().down);
    }
    if (trace) {
        auto at = [&](size_t i) { float ms = 0.f; cudaEventElapsedTime(&ms, tr[0], tr[i]); return ms; };
        fprintf(stderr, "[fm_dgemm_host] m=%d n=%d k=%d beta=%g: %d blocks of %d columns; A uploaded at %.2f ms\n", m, n, k, beta, nchunks, cw, at(1));
        fprintf(stderr, "[fm_dgemm_host] block  uploaded  product start  product done  downloaded   (ms after the first upload was queued)\n");
        for (int j = 0; j < nchunks; ++j)
            fprintf(stderr, "[fm_dgemm_host] %5d  %8.2f  %13.2f  %12.2f  %10.2f\n", j, at(2 + 4 * j), at(3 + 4 * j), at(4 + 4 * j), at(5 + 4 * j));
        for (auto &e : tr) cudaEventDestroy(e);
    }
    if (da && !a_dev) fm_free(da);
    if (db) fm_free(db);
    if (dc) fm_free(dc);
    return rc;
}

// dgesv_ / dgetrf_ on a HOST matrix (the reference's call: flomat.rkt:2140 / :1510 hand over GC memory).  Upload, factorisation and
// download used to run one after the other (n = 8192: 20 + 21.5 + 11 ms); here A goes up in column chunks on the transfer stream
// while k_getrf_streamed factors the chunks that have arrived, and every part of L\U comes down on the download stream as soon as
// no later step can touch it.  nrhs > 0: also solves for b (host or device) when the factorisation succeeded.  *hinfo = LAPACK info.
int lu_pipeline_min_n() {  // read at every call (tests lower it): below it upload, factorisation and download run one after the other
    if (getenv("FLOMAT_CUDA_NO_LU_PIPELINE")) return 0x7fffffff;
    const char *e = getenv("FLOMAT_CUDA_LU_PIPELINE_MIN_N");
    return e && atoi(e) >= 256 ? atoi(e) : 6144;
}
// chunk bounds of the pipeline: the first chunk is small (nothing overlaps its upload), the widths then double (wide chunks factor
// closer to the rate of the plain LU) up to `cap` (nothing overlaps the download of the last chunk's rows either)
std::vector<int> lu_chunk_bounds(int n, int first, int cap) {
    std::vector<int> b{0};
    for (int k = 0; b.back() < n; ++k) {
        long long w = (long long)first << (k > 1 ? k - 1 : 0);  // first, first, 2 first, 4 first, ...
        if (w > cap) w = cap;
        long long next = b.back() + w;
        if (next >= n || n - next < 128) next = n;
        b.push_back((int)next);
    }
    return b;
}
// the chunks getrf_host_pipeline uses for a matrix of order n: 1024, 1024, then doubling up to n/4 columns
std::vector<int> lu_host_bounds(int n) {
    int first = 1024, cap = (n / 4 / 128) * 128;  // two chunks (a window) arrive before the first panel can start; 512 .. 2048 measured alike
    if (cap < first) cap = first;
    if (const char *e = getenv("FLOMAT_CUDA_LU_PIPELINE_FIRST")) { const int v = atoi(e); if (v >= 128 && v % 128 == 0) { first = v; if (cap < first) cap = first; } }
    if (const char *e = getenv("FLOMAT_CUDA_LU_PIPELINE_CHUNK")) { const int v = atoi(e); if (v >= 128 && v % 128 == 0) first = cap = v; }  // developer switch: uniform chunks
    return lu_chunk_bounds(n, first, cap);
}
int getrf_host_pipeline(int n, double *a, int lda, int32_t *ipiv, int nrhs, double *b, int ldb, int32_t *hinfo) {
    FM_TRY(xfer_init());
    const int ldd = n + (n & 1);
    const std::vector<int> bounds = lu_host_bounds(n);
    const int nchunks = (int)bounds.size() - 1;
    auto chunk_of = [&](int c0) { int j = 0; while (bounds[j] != c0) ++j; return j; };
    int rc = FM_OK;
    double *da = static_cast<double *>(fm_alloc_bytes((size_t)ldd * n * 8));
    const bool ipiv_dev = is_device_ptr(ipiv);
    int32_t *d_ipiv = ipiv_dev ? ipiv : static_cast<int32_t *>(fm_alloc_bytes((size_t)n * 4 + 16));
    int32_t *d_info = ipiv_dev ? static_cast<int32_t *>(fm_alloc_bytes(16)) : (d_ipiv ? d_ipiv + n : nullptr);
    if (!da || !d_ipiv || !d_info) rc = set_error(FM_ERR_NOMEM, "dgesv_/dgetrf_: staging allocation failed (n=%d)", n);
    static PerDevice<std::vector<cudaEvent_t>> ev_pd;  // [3j] chunk j uploaded, [3j+1] its rows above the window final, [3j+2] its own rows final
    std::vector<int> upper_rows((size_t)nchunks, 0), done_cols((size_t)nchunks, 0);  // what the two events of chunk j cover
    std::vector<cudaEvent_t> &ev = ev_pd();
    while (rc == FM_OK && (int)ev.size() < 3 * nchunks) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "getrf_host_pipeline: event creation failed");
        else ev.push_back(e);
    }
    auto chk = [&](cudaError_t e) { if (e != cudaSuccess && rc == FM_OK) rc = set_error(FM_ERR_CUDA, "dgesv_/dgetrf_ (host pipeline): %s", cudaGetErrorString(e)); };
    // developer switch: per chunk, when its upload / factorisation / download finished (ms after the call started)
    static const bool trace = getenv("FLOMAT_HOST_LU_TRACE") != nullptr;
    std::vector<cudaEvent_t> tr;
    if (trace) {
        tr.resize(1 + 4 * (size_t)nchunks);
        for (auto &e : tr) cudaEventCreate(&e);
        cudaEventRecord(tr[0], ctx().stream);
        for (int j = 0; j < nchunks; ++j) cudaEventRecord(tr[2 + 4 * j], ctx().stream);  // chunk 0 has no U rows: keep the event valid
    }
    {
        Staged B;  // released (after a sync of the compute stream) before da is
        if (rc == FM_OK && nrhs > 0) rc = B.in(b, ldb, n, nrhs, true);
        if (rc == FM_OK) {
            chk(cudaEventRecord(xfer().tmp, ctx().stream));  // the staging blocks may be recycled ones that queued kernels still use
            chk(cudaStreamWaitEvent(xfer().stream, xfer().tmp, 0));
        }
        auto arrive = [&](int c0, int c1) -> int {
            const int j = chunk_of(c0);
            FM_TRY(copy_h2d(da + (size_t)c0 * ldd, ldd, a + (size_t)c0 * lda, lda, n, c1 - c0, xfer().stream));
            FM_CUDA(cudaEventRecord(ev[3 * j], xfer().stream));
            FM_CUDA(cudaStreamWaitEvent(ctx().stream, ev[3 * j], 0));
            if (trace) cudaEventRecord(tr[1 + 4 * j], xfer().stream);
            return FM_OK;
        };
        auto upper_done = [&](int c0, int, int rows) -> int {
            const int j = chunk_of(c0);
            upper_rows[j] = rows;
            FM_CUDA(cudaEventRecord(ev[3 * j + 1], ctx().stream));
            if (trace) cudaEventRecord(tr[2 + 4 * j], ctx().stream);
            return FM_OK;
        };
        auto done = [&](int c0, int, int lim) -> int {
            const int j = chunk_of(c0);
            done_cols[j] = lim;
            FM_CUDA(cudaEventRecord(ev[3 * j + 2], ctx().stream));
            if (trace) cudaEventRecord(tr[3 + 4 * j], ctx().stream);
            return FM_OK;
        };
        if (rc == FM_OK) rc = k_getrf_streamed(n, da, ldd, d_ipiv, d_info, bounds, arrive, upper_done, done);
        // The solve is queued before the factors' way back so that it runs beside the downloads.  With the right-hand sides in a staging
        // block that is safe when A turns out singular (LAPACK: b untouched): the block is simply not copied back.
        bool solved = false;
        if (rc == FM_OK && nrhs > 0 && B.owned) {
            rc = k_getrs(n, nrhs, da, ldd, 0, d_ipiv, 0, B.dev, B.ld, 0, 1);
            solved = true;
        }
        // downloads come after every upload is queued: a copy to pageable memory blocks this thread until it has run
        // in the order the events fire: step j - 1 releases the upper rows of chunk j, then (its panels done) the rows of chunk j - 1
        for (int j = 0; j <= nchunks && rc == FM_OK; ++j) {
            if (j < nchunks && upper_rows[j] > 0) {  // rows [0, upper_rows) of chunk j's columns
                const int c0 = bounds[j], c1 = bounds[j + 1];
                chk(cudaStreamWaitEvent(xfer().down, ev[3 * j + 1], 0));
                if (rc == FM_OK) rc = copy_d2h(a + (size_t)c0 * lda, lda, da + (size_t)c0 * ldd, ldd, upper_rows[j], c1 - c0, xfer().down);
            }
            if (j > 0) {  // rows of chunk j - 1 (the last chunk: down to row n) of the columns [0, done_cols)
                const int i = j - 1, r0 = bounds[i], r1 = i == nchunks - 1 ? n : bounds[i + 1];
                chk(cudaStreamWaitEvent(xfer().down, ev[3 * i + 2], 0));
                if (rc == FM_OK) rc = copy_d2h(a + r0, lda, da + r0, ldd, r1 - r0, done_cols[i], xfer().down);
                if (trace) cudaEventRecord(tr[4 + 4 * i], xfer().down);
            }
        }
        int32_t info = 0;
        if (rc == FM_OK) rc = fm_download_i32(&info, d_info, 1);  // synchronises the compute stream
        if (rc == FM_OK && nrhs > 0 && !solved && info == 0) rc = k_getrs(n, nrhs, da, ldd, 0, d_ipiv, 0, B.dev, B.ld, 0, 1);  // b in HBM: only now
        if (rc == FM_OK && nrhs > 0 && info == 0) rc = B.out();
        if (rc == FM_OK && !ipiv_dev) rc = fm_download_i32(ipiv, d_ipiv, (size_t)n);
        if (rc == FM_OK) *hinfo = info;
        chk(cudaEventRecord(xfer().last, xfer().stream));
        chk(cudaEventRecord(xfer().last_down, xfer().down));
        xfer().dirty = true;
        cudaStreamSynchronize(xfer().stream);
        cudaStreamSynchronize(xfer().down);  // pinned host memory: the copies above were asynchronous
        cudaStreamSynchronize(ctx().stream);
    }
    if (trace) {
        auto at = [&](size_t i) { float ms = 0.f; return cudaEventElapsedTime(&ms, tr[0], tr[i]) == cudaSuccess ? ms : -1.f; };
        fprintf(stderr, "[host LU pipeline] n=%d nrhs=%d: %d chunks\n[host LU pipeline] chunk  columns  uploaded  U rows final  factored  downloaded (ms)\n", n, nrhs, nchunks);
        for (int j = 0; j < nchunks; ++j)
            fprintf(stderr, "[host LU pipeline] %5d  %7d  %8.2f  %12.2f  %8.2f  %10.2f\n", j, bounds[j + 1] - bounds[j], at(1 + 4 * j), at(2 + 4 * j), at(3 + 4 * j), at(4 + 4 * j));
        for (auto &e : tr) cudaEventDestroy(e);
    }
    if (da) fm_free(da);
    if (ipiv_dev) { if (d_info) fm_free(d_info); } else if (d_ipiv) fm_free(d_ipiv);
    return rc;
}
}  // namespace fm
extern "C" {
int fm_dgemm_diag(int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc,
                  double diag) {
    FM_TRY(ensure_init());
    GemmEpilogue ep;
    ep.alpha = alpha;
    ep.beta = beta;
    ep.diag = diag;
    ep.has_diag = true;
    return k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, n, k, a, lda, 0, b, ldb, 0, c, ldc, 0, 1, ep);
}
int fm_dgemm_batched(int m, int n, int k, double alpha, const double *a, int lda, long long stride_a, const double *b, int ldb,
                     long long stride_b, double beta, double *c, int ldc, long long stride_c, double diag, int batch) {
    FM_TRY(ensure_init());
    GemmEpilogue ep;
    ep.alpha = alpha;
    ep.beta = beta;
    ep.diag = diag;
    ep.has_diag = (diag != 0.0);
    return k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, n, k, a, lda, stride_a, b, ldb, stride_b, c, ldc, stride_c, batch, ep);
}

// =================================== LU family ===================================
int fm_getrf(int m, int n, double *a, int lda, int32_t *ipiv, int32_t *info) {
    FM_TRY(ensure_init());
    return k_getrf(m, n, a, lda, 0, ipiv, 0, info, 1);
}
int fm_getrf_chunked(int n, double *a, int lda, int32_t *ipiv, int32_t *info, int chunk) {
    FM_TRY(ensure_init());
    if (!ipiv || !info) return set_error(FM_ERR_ARG, "fm_getrf_chunked: ipiv and info must be device pointers");
    if (chunk < 128 || chunk % 128 != 0) return set_error(FM_ERR_ARG, "fm_getrf_chunked: chunk must be a positive multiple of 128");
    auto nothing = [](int, int) { return FM_OK; };
    auto nothing3 = [](int, int, int) { return FM_OK; };
    return k_getrf_streamed(n, a, lda, ipiv, info, lu_chunk_bounds(n, chunk, chunk), nothing, nothing3, nothing3);
}
int fm_getrf_nopiv(int n, double *a, int lda, int32_t *ipiv, int32_t *info, int32_t *violated) {
    FM_TRY(ensure_init());
    if (!ipiv || !info || !violated) return set_error(FM_ERR_ARG, "fm_getrf_nopiv: ipiv, info and violated must be device pointers");
    FM_CUDA(cudaMemsetAsync(violated, 0, sizeof(int32_t), ctx().stream));
    return k_getrf_nopiv(n, a, lda, 0, ipiv, 0, info, violated, 1);
}
int fm_getrs(int n, int nrhs, const double *lu, int lda, const int32_t *ipiv, double *b, int ldb) {
    FM_TRY(ensure_init());
    return k_getrs(n, nrhs, lu, lda, 0, ipiv, 0, b, ldb, 0, 1);
}
// one device int32 to the host (syncs the library stream)
static int read_i32(const int32_t *dev, int32_t *out) {
    int32_t *h = reinterpret_cast<int32_t *>(ctx().h_scalar + 32);
    FM_CUDA(cudaMemcpyAsync(h, dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = *h;
    return FM_OK;
}
int fm_gesv(int n, int nrhs, double *a, int lda, int32_t *ipiv, double *b, int ldb, int32_t *info) {
    FM_TRY(ensure_init());
    if (!info) return set_error(FM_ERR_ARG, "fm_gesv: info is NULL");
    FM_TRY(k_getrf(n, n, a, lda, 0, ipiv, 0, info, 1));
    // LAPACK's dgesv solves only when the factorisation succeeded; with info > 0 B is left untouched, so a caller that ignores
    // info (flomat-solve-many!, flomat.rkt:2135-2142) gets its copy of B back.  One 4-byte readback decides.
    int32_t hinfo = 0;
    FM_TRY(read_i32(info, &hinfo));
    if (hinfo != 0) return FM_OK;
    return k_getrs(n, nrhs, a, lda, 0, ipiv, 0, b, ldb, 0, 1);
}
int fm_getri(int n, double *a, int lda, const int32_t *ipiv, int32_t *info) {
    FM_TRY(ensure_init());
    if (n < 0 || lda < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "fm_getri: bad dimensions");
    int32_t *d_info = info ? info : reinterpret_cast<int32_t *>(ctx().d_scalar + 16);
    if (n == 0) {
        if (info) FM_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), ctx().stream));
        return FM_OK;
    }
    // LAPACK's dgetri (through dtrtri) returns info = i for an exactly zero U(i,i) before any work: A stays the LU factors, which
    // is what `inv` of a singular matrix hands back in the reference (flomat-inverse!, flomat.rkt:1770-1779, info ignored)
    FM_TRY(k_first_zero_diag(n, a, lda, d_info));
    int32_t hinfo = 0;
    FM_TRY(read_i32(d_info, &hinfo));
    if (hinfo != 0) return FM_OK;
    return k_getri(n, a, lda, ipiv);
}
int fm_det(int n, const double *lu, int lda, const int32_t *ipiv, double *out) {
    FM_TRY(ensure_init());
    FM_TRY(k_det_device(n, lu, lda, ipiv, ctx().d_scalar));
    FM_CUDA(cudaMemcpyAsync(ctx().h_scalar, ctx().d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = ctx().h_scalar[0];
    return FM_OK;
}

// =================================== neighbours of the hot path (N2-N4) ===================================
int fm_transpose(int m, int n, const double *a, int lda, double *b, int ldb) {
    FM_TRY(ensure_init());
    return k_transpose(m, n, a, lda, b, ldb);
}
int fm_tri(int lower, int unit_diag, int m, int n, const double *a, int lda, double *b, int ldb) {
    FM_TRY(ensure_init());
    return k_tri(lower, unit_diag, m, n, a, lda, b, ldb);
}
int fm_pivots_to_matrix(int k, const int32_t *ipiv, double *p, int ldp) {
    FM_TRY(ensure_init());
    return k_pivots_to_matrix(k, ipiv, p, ldp);
}
int fm_map(int fn, int m, int n, const double *a, int lda, double *c, int ldc) {
    FM_TRY(ensure_init());
    if (fn < 5 || fn > 11) return set_error(FM_ERR_ARG, "fm_map: fn %d not in FM_MAP_SIN..FM_MAP_EXP", fn);
    return k_ewise(fn, m, n, a, lda, 0.0, nullptr, 0, 0.0, c, ldc);
}
int fm_gemv(int trans, int m, int n, double alpha, const double *a, int lda, const double *x, int incx, double beta, double *y, int incy) {
    FM_TRY(ensure_init());
    if (trans != FM_TRANS && trans != FM_NO_TRANS) return set_error(FM_ERR_ARG, "fm_gemv: trans must be 111 or 112");
    if (incx <= 0 || incy <= 0) return set_error(FM_ERR_ARG, "fm_gemv: increments must be positive");
    if (!x) return set_error(FM_ERR_ARG, "fm_gemv: x is NULL");
    return k_gemv(trans == FM_TRANS, m, n, alpha, a, lda, x, incx, beta, y, incy);
}
static int scalar_back(double *out) {
    FM_CUDA(cudaMemcpyAsync(ctx().h_scalar, ctx().d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    FM_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = ctx().h_scalar[0];
    return FM_OK;
}
int fm_dot(int n, const double *x, int incx, const double *y, int incy, double *out) {
    FM_TRY(ensure_init());
    if (incx < 0 || incy < 0) return set_error(FM_ERR_ARG, "fm_dot: negative increments are not used by flomat and not supported");
    FM_TRY(k_dot_device(n, x, incx, y, incy, ctx().d_scalar));
    return scalar_back(out);
}
int fm_nrm2(int n, const double *x, int incx, double *out) {
    FM_TRY(ensure_init());
    if (incx <= 0) { *out = 0.0; return FM_OK; }
    FM_TRY(k_nrm2_device(n, x, incx, ctx().d_scalar));
    return scalar_back(out);
}
int fm_iamax(int n, const double *x, int incx, int *out) {
    FM_TRY(ensure_init());
    *out = 0;
    if (n <= 0 || incx <= 0) return FM_OK;
    FM_TRY(k_iamax_device(n, x, incx, reinterpret_cast<long long *>(ctx().d_scalar)));
    double raw;
    FM_TRY(scalar_back(&raw));
    long long idx;
    memcpy(&idx, &raw, sizeof(idx));
    *out = (int)idx;
    return FM_OK;
}
int fm_swap(int n, double *x, int incx, double *y, int incy) {
    FM_TRY(ensure_init());
    if (incx < 0 || incy < 0) return set_error(FM_ERR_ARG, "fm_swap: negative increments are not used by flomat and not supported");
    return k_swap(n, x, incx, y, incy);
}
int fm_colsums(int m, int n, const double *a, int lda, double *out) {  // out[j] = sum_i A(i,j): A' * ones
    FM_TRY(ensure_init());
    return k_gemv(1, m, n, 1.0, a, lda, nullptr, 0, 0.0, out, 1);
}
int fm_rowsums(int m, int n, const double *a, int lda, double *out) {  // out[i] = sum_j A(i,j): A * ones
    FM_TRY(ensure_init());
    return k_gemv(0, m, n, 1.0, a, lda, nullptr, 0, 0.0, out, 1);
}
int fm_potrf(char uplo, int n, double *a, int lda, int32_t *info) {
    FM_TRY(ensure_init());
    if (!info) return set_error(FM_ERR_ARG, "fm_potrf: info is NULL");
    return k_potrf(uplo, n, a, lda, info);
}
int fm_larnv(int idist, int32_t *iseed, long long n, double *x) {
    FM_TRY(ensure_init());
    if (!iseed) return set_error(FM_ERR_ARG, "fm_larnv: iseed is NULL");
    return k_larnv(idist, iseed, n, x);
}

// =================================== expm ===================================
int fm_expm(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *info_host) {
    FM_TRY(ensure_init());
    // a HOST result buffer: the last squaring is cut into column blocks whose downloads overlap the blocks still to be computed;
    // the buffer is valid after fm_transfer_sync() / fm_sync() (pinned memory) or on return (pageable)
    if (n > 0 && out && !is_device_ptr(out)) return k_expm_host_out(n, a, lda, out, ldo, mode, s_out, info_host);
    return k_expm(n, a, lda, out, ldo, mode, s_out, info_host);
}
int fm_expm_batched(int n, const double *a, int lda, long long stride_a, double *out, int ldo, long long stride_o, int batch, int mode,
                    double *s_out) {
    FM_TRY(ensure_init());
    return k_expm_batched(n, a, lda, stride_a, out, ldo, stride_o, batch, mode, s_out);
}

// =================================== multi-GPU ===================================
int fm_ipc_export(void *dev, void *handle_out, size_t *offset_out) {
    FM_TRY(ensure_init());
    static_assert(sizeof(cudaIpcMemHandle_t) == FM_IPC_HANDLE_BYTES, "ipc handle size");
    // An IPC handle names the whole allocation; the importer needs the offset of `dev` inside it.
    typedef int (*RangeFn)(unsigned long long *, size_t *, unsigned long long);
    static RangeFn range = nullptr;
    if (!range) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FM_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) return set_error(FM_ERR_CUDA, "cuMemGetAddressRange entry point not available");
        range = reinterpret_cast<RangeFn>(p);
    }
    unsigned long long base = 0;
    size_t size = 0;
    const int r = range(&base, &size, (unsigned long long)reinterpret_cast<uintptr_t>(dev));
    if (r != 0) return set_error(FM_ERR_CUDA, "cuMemGetAddressRange failed with CUresult %d", r);
    cudaIpcMemHandle_t h;
    FM_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void *>(static_cast<uintptr_t>(base))));
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = static_cast<size_t>(reinterpret_cast<uintptr_t>(dev) - static_cast<uintptr_t>(base));
    return FM_OK;
}
static std::vector<std::pair<void *, void *>> g_ipc_open;  // (pointer handed out, mapping base)
int fm_ipc_open(const void *handle, size_t offset, void **dev_out) {
    FM_TRY(ensure_init());
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *base = nullptr;
    FM_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_out = static_cast<char *>(base) + offset;
    g_ipc_open.emplace_back(*dev_out, base);
    return FM_OK;
}
int fm_ipc_close(void *dev) {
    for (size_t i = 0; i < g_ipc_open.size(); ++i)
        if (g_ipc_open[i].first == dev) {
            void *base = g_ipc_open[i].second;
            g_ipc_open.erase(g_ipc_open.begin() + i);
            FM_CUDA(cudaIpcCloseMemHandle(base));
            return FM_OK;
        }
    return set_error(FM_ERR_ARG, "fm_ipc_close: pointer was not returned by fm_ipc_open");
}
int fm_dgemm_colshard(int m, int nloc, int k, const double *a, int lda, const double *bloc, int ldb, double *c, int ldc, int j0,
                      double *const *peer_c, int npeers) {
    FM_TRY(ensure_init());
    if (npeers < 0 || npeers > 7) return set_error(FM_ERR_ARG, "fm_dgemm_colshard: npeers must be 0..7");
    GemmEpilogue ep;
    ep.alpha = 1.0;
    ep.beta = 0.0;
    double *shifted[7];
    for (int r = 0; r < npeers; ++r) shifted[r] = peer_c[r] + (size_t)j0 * ldc;
    ep.peers = shifted;
    ep.npeers = npeers;
    return k_dgemm(FM_NO_TRANS, FM_NO_TRANS, m, nloc, k, a, lda, 0, bloc, ldb, 0, c + (size_t)j0 * ldc, ldc, 0, 1, ep);
}

// =================================== compat tier ===================================
void cblas_dgemm(int order, int transA, int transB, int M, int N, int K, double alpha, const double *A, int lda, const double *B, int ldb,
                 double beta, double *C, int ldc) {
    int rc = ensure_init();
    if (rc == FM_OK && order != FM_COL_MAJOR) rc = set_error(FM_ERR_ARG, "cblas_dgemm: only CblasColMajor (102) is supported, as flomat.rkt:888 passes");
    if (rc == FM_OK) {
        const bool ta = transA == FM_TRANS, tb = transB == FM_TRANS;
        // all operands on the host and a product large enough for the transfers to matter: column-block pipeline
        if (!ta && !tb && transA == FM_NO_TRANS && transB == FM_NO_TRANS && M > 0 && N >= 1024 && K > 0 && (double)M * N * K >= 1073741824.0 &&
            !is_device_ptr(A) && !is_device_ptr(B) && !is_device_ptr(C)) {
            rc = fm_dgemm_host(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 0);
            report("cblas_dgemm", rc);
            return;
        }
        Staged a, b, c;
        rc = a.in(A, lda, ta ? K : M, ta ? M : K, true);
        if (rc == FM_OK) rc = b.in(B, ldb, tb ? N : K, tb ? K : N, true);
        if (rc == FM_OK) rc = c.in(C, ldc, M, N, beta != 0.0);
        if (rc == FM_OK) {
            GemmEpilogue ep;
            ep.alpha = alpha;
            ep.beta = beta;
            rc = k_dgemm(transA, transB, M, N, K, a.dev, a.ld, 0, b.dev, b.ld, 0, c.dev, c.ld, 0, 1, ep);
        }
        if (rc == FM_OK) rc = c.out();
        if (rc == FM_OK && c.owned && cudaStreamSynchronize(ctx().stream) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "cblas_dgemm: sync failed");
    }
    report("cblas_dgemm", rc);
}

// vectors: stage n elements with their stride collapsed to 1
struct StagedVec {
    double *dev = nullptr;
    long long inc = 1;
    bool owned = false;
    double *host = nullptr;
    int n = 0, inch = 1;
    std::vector<double> pack;
    int in(const double *p, int incp, int count, bool copy_in) {
        n = count;
        if (is_device_ptr(p) || count <= 0) { dev = const_cast<double *>(p); inc = incp; return FM_OK; }
        host = const_cast<double *>(p); inch = incp;
        const int len = (incp == 0) ? 1 : count;
        if (!(dev = static_cast<double *>(fm_alloc_bytes((size_t)(len < 2 ? 2 : len) * 8)))) return set_error(FM_ERR_NOMEM, "staging allocation failed");
        owned = true;
        inc = (incp == 0) ? 0 : 1;
        if (copy_in) {
            const long long ainc = incp < 0 ? -(long long)incp : incp;
            const double *start = host;  // BLAS: negative inc walks backwards from (1-n)*inc
            if (ainc <= 1) {
                FM_CUDA(cudaMemcpyAsync(dev, start, (size_t)len * 8, cudaMemcpyHostToDevice, ctx().stream));
            } else {
                FM_CUDA(cudaMemcpy2DAsync(dev, 8, start, (size_t)ainc * 8, 8, len, cudaMemcpyHostToDevice, ctx().stream));
            }
        }
        return FM_OK;
    }
    int out() {
        if (!owned) return FM_OK;
        const long long ainc = inch < 0 ? -(long long)inch : inch;
        const int len = (inch == 0) ? 1 : n;  // as staged by in(): a stride-0 vector is one element
        if (ainc <= 1) FM_CUDA(cudaMemcpyAsync(host, dev, (size_t)len * 8, cudaMemcpyDeviceToHost, ctx().stream));
        else FM_CUDA(cudaMemcpy2DAsync(host, (size_t)ainc * 8, dev, 8, 8, len, cudaMemcpyDeviceToHost, ctx().stream));
        FM_CUDA(cudaStreamSynchronize(ctx().stream));
        return FM_OK;
    }
    ~StagedVec() { if (owned && dev) { cudaStreamSynchronize(ctx().stream); fm_free(dev); } }
};

// The reference's `plus` / `minus` arrive as ONE cblas_daxpy PER COLUMN with unit strides and host pointers (flomat.rkt:805-813): an
// n = 8192 matrix sum is 8192 calls of 64 KB each, 90 112 calls per expm.  For that shape the generic staging (two allocations, two
// pageable uploads, each with its own driver-side staging and synchronisation) costs ~55 us per call; here X and Y are packed into one
// pinned buffer, travel up in one copy, and Y comes back through the same buffer.  Per device: 2 x 512 KB pinned + 2 x 512 KB of HBM.
struct SmallVecStage {
    static constexpr size_t CAP = 65536;  // doubles per vector
    double *pin = nullptr, *dev = nullptr;
};
static PerDevice<SmallVecStage> g_small;
static int daxpy_small_host(int n, double alpha, const double *X, double *Y) {
    SmallVecStage &sv = g_small();
    if (!sv.pin) {
        FM_CUDA(cudaMallocHost(&sv.pin, 2 * SmallVecStage::CAP * sizeof(double)));
        FM_CUDA(cudaMalloc(&sv.dev, 2 * SmallVecStage::CAP * sizeof(double)));
    }
    const size_t bytes = (size_t)n * sizeof(double);
    memcpy(sv.pin, X, bytes);
    memcpy(sv.pin + n, Y, bytes);
    cudaStream_t st = ctx().stream;
    FM_CUDA(cudaMemcpyAsync(sv.dev, sv.pin, 2 * bytes, cudaMemcpyHostToDevice, st));
    FM_TRY(k_axpy_strided(n, alpha, sv.dev, 1, sv.dev + n, 1));
    FM_CUDA(cudaMemcpyAsync(sv.pin + n, sv.dev + n, bytes, cudaMemcpyDeviceToHost, st));
    FM_CUDA(cudaStreamSynchronize(st));
    memcpy(Y, sv.pin + n, bytes);
    return FM_OK;
}

void cblas_daxpy(int n, double alpha, const double *X, int incX, double *Y, int incY) {
    if (n <= 0) return;
    int rc = ensure_init();
    if (rc == FM_OK && (incX < 0 || incY <= 0)) rc = set_error(FM_ERR_ARG, "cblas_daxpy: negative increments (and incY = 0) are not used by flomat and not supported");
    if (rc == FM_OK && incX == 1 && incY == 1 && (size_t)n <= SmallVecStage::CAP && !is_device_ptr(X) && !is_device_ptr(Y)) {
        rc = daxpy_small_host(n, alpha, X, Y);
        report("cblas_daxpy", rc);
        return;
    }
    if (rc == FM_OK) {
        StagedVec x, y;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = y.in(Y, incY, n, true);
        if (rc == FM_OK) rc = k_axpy_strided(n, alpha, x.dev, x.inc, y.dev, y.inc);
        if (rc == FM_OK) rc = y.out();
    }
    report("cblas_daxpy", rc);
}
void cblas_dcopy(int n, const double *X, int incX, double *Y, int incY) {
    if (n <= 0) return;
    int rc = ensure_init();
    if (rc == FM_OK && (incX < 0 || incY <= 0)) rc = set_error(FM_ERR_ARG, "cblas_dcopy: negative increments (and incY = 0) are not used by flomat and not supported");
    if (rc == FM_OK) {
        StagedVec x, y;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = y.in(Y, incY, n, incY > 1);  // strided host Y: keep the gaps' contents
        if (rc == FM_OK) rc = k_copy_strided(n, x.dev, x.inc, y.dev, y.inc);
        if (rc == FM_OK) rc = y.out();
    }
    report("cblas_dcopy", rc);
}
void cblas_dscal(int n, double alpha, double *X, int incX) {
    if (n <= 0 || incX <= 0) return;
    int rc = ensure_init();
    if (rc == FM_OK) {
        StagedVec x;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = k_scal_strided(n, alpha, x.dev, x.inc);
        if (rc == FM_OK) rc = x.out();
    }
    report("cblas_dscal", rc);
}

void cblas_dgemv(int order, int transA, int M, int N, double alpha, const double *A, int lda, const double *X, int incX, double beta,
                 double *Y, int incY) {
    if (M <= 0 || N <= 0) return;
    int rc = ensure_init();
    if (rc == FM_OK && order != 102) rc = set_error(FM_ERR_ARG, "cblas_dgemv: only CblasColMajor (102) is used by flomat (flomat.rkt:949)");
    if (rc == FM_OK && (incX <= 0 || incY <= 0)) rc = set_error(FM_ERR_ARG, "cblas_dgemv: increments must be positive");
    if (rc == FM_OK) {
        const bool t = transA != FM_NO_TRANS;
        Staged a;
        StagedVec x, y;
        rc = a.in(A, lda, M, N, true);
        if (rc == FM_OK) rc = x.in(X, incX, t ? M : N, true);
        if (rc == FM_OK) rc = y.in(Y, incY, t ? N : M, true);
        if (rc == FM_OK) rc = k_gemv(t, M, N, alpha, a.dev, a.ld, x.dev, x.inc, beta, y.dev, y.inc);
        if (rc == FM_OK) rc = y.out();
    }
    report("cblas_dgemv", rc);
}
double cblas_ddot(int n, const double *X, int incX, const double *Y, int incY) {
    if (n <= 0) return 0.0;
    double out = 0.0;
    int rc = ensure_init();
    // incX = 0 / incY = 0 are legal here (`unsafe-sum`, flomat.rkt:2074: ddot against one 1.0); negative strides are not used by flomat
    if (rc == FM_OK && (incX < 0 || incY < 0)) rc = set_error(FM_ERR_ARG, "cblas_ddot: negative increments are not used by flomat and not supported");
    if (rc == FM_OK) {
        StagedVec x, y;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = y.in(Y, incY, n, true);
        if (rc == FM_OK) rc = fm_dot(n, x.dev, (int)x.inc, y.dev, (int)y.inc, &out);
    }
    report("cblas_ddot", rc);
    return rc == FM_OK ? out : nan("");
}
double cblas_dnrm2(int n, const double *X, int incX) {
    if (n <= 0 || incX <= 0) return 0.0;
    double out = 0.0;
    int rc = ensure_init();
    if (rc == FM_OK) {
        StagedVec x;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = fm_nrm2(n, x.dev, (int)x.inc, &out);
    }
    report("cblas_dnrm2", rc);
    return rc == FM_OK ? out : nan("");
}
int cblas_idamax(int n, const double *X, int incX) {
    if (n <= 0 || incX <= 0) return 0;
    int out = 0;
    int rc = ensure_init();
    if (rc == FM_OK) {
        StagedVec x;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = fm_iamax(n, x.dev, (int)x.inc, &out);
    }
    report("cblas_idamax", rc);
    return out;
}
void cblas_dswap(int n, double *X, int incX, double *Y, int incY) {
    if (n <= 0) return;
    int rc = ensure_init();
    if (rc == FM_OK && (incX <= 0 || incY <= 0)) rc = set_error(FM_ERR_ARG, "cblas_dswap: increments must be positive");
    if (rc == FM_OK) {
        StagedVec x, y;
        rc = x.in(X, incX, n, true);
        if (rc == FM_OK) rc = y.in(Y, incY, n, true);
        if (rc == FM_OK) rc = k_swap(n, x.dev, x.inc, y.dev, y.inc);
        if (rc == FM_OK) rc = x.out();
        if (rc == FM_OK) rc = y.out();
    }
    report("cblas_dswap", rc);
}
void dpotrf_(const char *uplo, const int *n, double *a, const int *lda, int *info) {
    const char u = uplo ? uplo[0] : 'L';
    *info = 0;
    if (u != 'U' && u != 'u' && u != 'L' && u != 'l') { *info = -1; return; }
    if (*n < 0) { *info = -2; return; }
    if (*lda < (*n > 1 ? *n : 1)) { *info = -4; return; }
    if (*n == 0) return;
    int rc = ensure_init();
    if (rc == FM_OK) {
        Staged A;
        rc = A.in(a, *lda, *n, *n, true);
        int32_t *d_info = rc == FM_OK ? static_cast<int32_t *>(fm_alloc_bytes(16)) : nullptr;
        if (rc == FM_OK && !d_info) rc = FM_ERR_NOMEM;
        if (rc == FM_OK) rc = k_potrf(u, *n, A.dev, A.ld, d_info);
        if (rc == FM_OK) rc = A.out();
        int32_t hinfo = 0;
        if (rc == FM_OK) rc = fm_download_i32(&hinfo, d_info, 1);
        if (rc == FM_OK) *info = hinfo;
        if (d_info) fm_free(d_info);
    }
    report("dpotrf_", rc);
}
void dlarnv_(const int *idist, int32_t *iseed, const int *n, double *x) {
    if (*n <= 0) return;
    int rc = ensure_init();
    if (rc == FM_OK) {
        StagedVec v;
        rc = v.in(x, 1, *n, false);
        if (rc == FM_OK) rc = k_larnv(*idist, iseed, *n, v.dev);
        if (rc == FM_OK) rc = v.out();
    }
    report("dlarnv_", rc);
}

double dlange_(const char *norm, const int *m, const int *n, const double *a, const int *lda, double *work) {
    (void)work;  // the row-sum workspace lives in HBM
    int rc = ensure_init();
    double out = 0.0;
    if (rc == FM_OK) {
        Staged A;
        rc = A.in(a, *lda, *m, *n, true);
        if (rc == FM_OK) rc = fm_norm(norm[0], *m, *n, A.dev, A.ld, &out);
    }
    report("dlange_", rc);
    return rc == FM_OK ? out : nan("");
}

static int lapack_arg_check_getrf(int m, int n, int lda) {
    if (m < 0) return -1;
    if (n < 0) return -2;
    if (lda < (m > 1 ? m : 1)) return -4;
    return 0;
}

void dgetrf_(const int *m, const int *n, double *a, const int *lda, int32_t *ipiv, int *info) {
    *info = lapack_arg_check_getrf(*m, *n, *lda);
    if (*info != 0 || *m == 0 || *n == 0) return;
    int rc = ensure_init();
    if (rc == FM_OK && *m == *n && *n >= lu_pipeline_min_n() && !is_device_ptr(a)) {
        int32_t hinfo = 0;
        rc = getrf_host_pipeline(*n, a, *lda, ipiv, 0, nullptr, 1, &hinfo);
        if (rc == FM_OK) *info = hinfo;
    } else if (rc == FM_OK) {
        const int mn = *m < *n ? *m : *n;
        Staged A;
        rc = A.in(a, *lda, *m, *n, true);
        int32_t *d_ipiv = nullptr, *d_info = nullptr;
        const bool ipiv_dev = is_device_ptr(ipiv);
        if (rc == FM_OK) {
            d_ipiv = ipiv_dev ? ipiv : static_cast<int32_t *>(fm_alloc_bytes((size_t)mn * 4 + 16));
            if (!d_ipiv) rc = FM_ERR_NOMEM;
            else d_info = ipiv_dev ? static_cast<int32_t *>(fm_alloc_bytes(16)) : d_ipiv + mn;
            if (rc == FM_OK && !d_info) rc = FM_ERR_NOMEM;
        }
        if (rc == FM_OK) rc = k_getrf(*m, *n, A.dev, A.ld, 0, d_ipiv, 0, d_info, 1);
        if (rc == FM_OK) rc = A.out();
        int32_t hinfo = 0;
        if (rc == FM_OK) rc = fm_download_i32(&hinfo, d_info, 1);
        if (rc == FM_OK && !ipiv_dev) rc = fm_download_i32(ipiv, d_ipiv, (size_t)mn);
        if (rc == FM_OK) *info = hinfo;
        if (ipiv_dev) { if (d_info) fm_free(d_info); } else if (d_ipiv) fm_free(d_ipiv);
    }
    report("dgetrf_", rc);
    if (rc != FM_OK) *info = rc;
}

void dgesv_(const int *n, const int *nrhs, double *a, const int *lda, int32_t *ipiv, double *b, const int *ldb, int *info) {
    *info = 0;
    if (*n < 0) *info = -1;
    else if (*nrhs < 0) *info = -2;
    else if (*lda < (*n > 1 ? *n : 1)) *info = -4;
    else if (*ldb < (*n > 1 ? *n : 1)) *info = -7;
    if (*info != 0 || *n == 0) return;
    int rc = ensure_init();
    if (rc == FM_OK && *n >= lu_pipeline_min_n() && !is_device_ptr(a)) {
        int32_t hinfo = 0;
        rc = getrf_host_pipeline(*n, a, *lda, ipiv, *nrhs, b, *ldb, &hinfo);
        if (rc == FM_OK) *info = hinfo;
    } else if (rc == FM_OK) {
        Staged A, B;
        rc = A.in(a, *lda, *n, *n, true);
        if (rc == FM_OK) rc = B.in(b, *ldb, *n, *nrhs, true);
        const bool ipiv_dev = is_device_ptr(ipiv);
        int32_t *d_ipiv = nullptr, *d_info = nullptr;
        if (rc == FM_OK) {
            d_ipiv = ipiv_dev ? ipiv : static_cast<int32_t *>(fm_alloc_bytes((size_t)*n * 4 + 16));
            if (!d_ipiv) rc = FM_ERR_NOMEM;
            else d_info = ipiv_dev ? static_cast<int32_t *>(fm_alloc_bytes(16)) : d_ipiv + *n;
            if (rc == FM_OK && !d_info) rc = FM_ERR_NOMEM;
        }
        if (rc == FM_OK) rc = k_getrf(*n, *n, A.dev, A.ld, 0, d_ipiv, 0, d_info, 1);
        int32_t hinfo = 0;
        if (rc == FM_OK) rc = fm_download_i32(&hinfo, d_info, 1);
        // LAPACK dgesv: solve only when the factorisation succeeded (info == 0); B is left untouched otherwise
        if (rc == FM_OK && hinfo == 0 && *nrhs > 0) rc = k_getrs(*n, *nrhs, A.dev, A.ld, 0, d_ipiv, 0, B.dev, B.ld, 0, 1);
        if (rc == FM_OK) rc = A.out();
        if (rc == FM_OK && hinfo == 0) rc = B.out();
        if (rc == FM_OK && !ipiv_dev) rc = fm_download_i32(ipiv, d_ipiv, (size_t)*n);
        if (rc == FM_OK) { cudaStreamSynchronize(ctx().stream); *info = hinfo; }
        if (ipiv_dev) { if (d_info) fm_free(d_info); } else if (d_ipiv) fm_free(d_ipiv);
    }
    report("dgesv_", rc);
    if (rc != FM_OK) *info = rc;
}

void dgetri_(const int *n, double *a, const int *lda, const int32_t *ipiv, double *work, const int *lwork, int *info) {
    (void)work;
    *info = 0;
    if (*n < 0) *info = -1;
    else if (*lda < (*n > 1 ? *n : 1)) *info = -3;
    if (*info != 0 || *n == 0) return;
    if (lwork && *lwork == -1) { if (work && !is_device_ptr(work)) work[0] = (double)*n; return; }  // workspace query
    int rc = ensure_init();
    if (rc == FM_OK) {
        Staged A;
        rc = A.in(a, *lda, *n, *n, true);
        const bool ipiv_dev = is_device_ptr(ipiv);
        int32_t *d_ipiv = const_cast<int32_t *>(ipiv);
        if (rc == FM_OK && !ipiv_dev) {
            d_ipiv = static_cast<int32_t *>(fm_alloc_bytes((size_t)*n * 4));
            if (!d_ipiv) rc = FM_ERR_NOMEM;
            else rc = fm_upload_i32(d_ipiv, ipiv, (size_t)*n);
        }
        // LAPACK dgetri: info = i if U(i,i) is exactly zero (singular), checked before any work (fm_getri scans the diagonal)
        int32_t *d_info = rc == FM_OK ? static_cast<int32_t *>(fm_alloc_bytes(16)) : nullptr;
        if (rc == FM_OK && !d_info) rc = FM_ERR_NOMEM;
        if (rc == FM_OK) rc = fm_getri(*n, A.dev, A.ld, d_ipiv, d_info);
        int32_t hinfo = 0;
        if (rc == FM_OK) rc = read_i32(d_info, &hinfo);
        if (rc == FM_OK) *info = hinfo;
        if (d_info) fm_free(d_info);
        if (rc == FM_OK && *info == 0) rc = A.out();
        if (rc == FM_OK) cudaStreamSynchronize(ctx().stream);
        if (!ipiv_dev && d_ipiv) fm_free(d_ipiv);
    }
    report("dgetri_", rc);
    if (rc != FM_OK) *info = rc;
}

}  // extern "C"
