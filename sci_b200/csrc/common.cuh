// common.cuh -- shared context, error plumbing and small device helpers for libflomat_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include <vector>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>

#include "../../include/flomat_cuda.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libflomat_cuda is written for sm_100a (B200) only"
#endif

namespace fm {

constexpr int FM_MAX_DEVICES = 16;

// One Context per CUDA device.  A host thread is bound to one device at a time (fm_init, or the default device on its first call);
// ctx() is the bound device's context.  Calls that target the same device must be serialised by the caller (the reference
// calls from one OS thread); the single-process multi-GPU entry points (fm_mg_*, mg.cu) drive every device from its own worker thread.
struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;      // library stream (legacy default unless fm_set_stream)
    unsigned long long launches = 0;    // kernels launched by this library
    // scratch that lives for the life of the library (grown on demand, never shrunk)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    double *d_scalar = nullptr;         // 64 doubles of device scratch for reductions
    double *h_scalar = nullptr;         // pinned mirror
    int max_smem_optin = 0;
};

Context &ctx();
int set_error(int code, const char *fmt, ...);
int ensure_init();
int bind_device(int device);  // bind the calling thread to `device` (cudaSetDevice + context creation on first use)
int bound_device();           // the calling thread's device, -1 if it has none yet
unsigned long long launches_of_device(int d);
const char *error_text();     // the calling thread's last error message
bool is_device_pointer(const void *p, int *device_out);  // device (or managed) memory? and on which device
int xfer_streams(cudaStream_t *up, cudaStream_t *down);  // the bound device's transfer streams (created on first use)
int dgemm_host_pipeline(int m, int n, int k, double alpha, const double *a_host, int lda, const double *a_dev, int ldda, const double *b,
                        int ldb, double beta, double *c, int ldc, int async);

// state that exists once per device (kernel attributes, workspaces, event pools ...): `static PerDevice<T> x;  x()` is the slot of the
// calling thread's device
template <class T>
struct PerDevice {
    T v[FM_MAX_DEVICES]{};
    T &operator()() { const int d = ctx().device; return v[d < 0 ? 0 : d]; }
};
void *scratch(size_t bytes);  // returns nullptr + sets error on failure

inline void count_launch(int n = 1) { ctx().launches += (unsigned long long)n; }

#define FM_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fm::set_error(FM_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
                                 cudaGetErrorString(e__));                                              \
    } while (0)

#define FM_LAUNCH_CHECK()                                                                               \
    do {                                                                                                \
        cudaError_t e__ = cudaGetLastError();                                                           \
        if (e__ != cudaSuccess)                                                                         \
            return fm::set_error(FM_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__,  \
                                 cudaGetErrorString(e__));                                              \
        fm::count_launch();                                                                             \
    } while (0)

#define FM_TRY(expr)                  \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != FM_OK) return rc__; \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace fm

// ---- kernels' host entry points (device pointers, library stream) ------------------------------
namespace fm {
// ewise.cu
// host <-> HBM copies that stage pageable host memory through pinned bounce buffers with several host threads (xfer.cu)
int copy_h2d(double *dev, size_t ldd, const double *host, size_t ldh, size_t m, size_t n, cudaStream_t st);
int copy_d2h(double *host, size_t ldh, const double *dev, size_t ldd, size_t m, size_t n, cudaStream_t st);
int k_fill(double *a, int lda, int m, int n, double x);
int k_eye(double *a, int lda, int m, int n, int k);
int k_copy2d(double *dst, int ldd, const double *src, int lds, int m, int n);
int k_ewise(int op, int m, int n, const double *a, int lda, double sa, const double *b, int ldb, double sb, double *c, int ldc);
int k_axpy2d(int m, int n, double alpha, const double *a, int lda, double *b, int ldb);
int k_scal2d(int m, int n, double s, double *a, int lda);
int k_addsub(int m, int n, const double *e, int lde, const double *o, int ldo, double *num, int ldn, double *den, int ldd);
int k_axpy_strided(int n, double alpha, const double *x, long long incx, double *y, long long incy);
int k_copy_strided(int n, const double *x, long long incx, double *y, long long incy);
int k_scal_strided(int n, double alpha, double *x, long long incx);
int k_lincomb(int m, int n, double ci, double c1, const double *a1, int ld1, double c2, const double *a2, int ld2, double *out, int ldo,
              long long s1, long long s2, long long so, int batch);
// out_k := ci_k*I + c1_k*a1 + c2_k*a2 over `batch` contiguous ld x n workspaces (ld even, item stride ld*n); a2, o1 may be null
int k_pade_terms(int n, int ld, int batch, const double *a1, const double *a2, double ci0, double c10, double c20, double *o0, double ci1,
                 double c11, double c21, double *o1);
int k_div_pow2_batched(int n, const double *a, int lda, long long sa, double *x, int ldx, long long sx, const double *scale_dev, int batch);
// norm.cu
int k_norm_device(char type, int m, int n, const double *a, int lda, double *d_out);  // result stays on device
int k_norm1_batched(int n, const double *a, int lda, long long stride, int batch, double *d_out);
// dgemm.cu
struct GemmEpilogue {
    double alpha = 1.0, beta = 0.0;
    double diag = 0.0;          // added to C[i,i] after alpha/beta (fused `+ c*I`)
    bool has_diag = false;
    int diag_col0 = 0;          // C is the column block [diag_col0, ...) of a larger matrix: diag goes to C[j + diag_col0, j]
    double *const *peers = nullptr;  // host array of peer C base pointers (column-shard P2P epilogue)
    int npeers = 0;
    const int32_t *active_until = nullptr;  // batched: item z is skipped when round >= active_until[z]
    int round = 0;
    int max_sms = 0;             // > 0: use at most this many SMs (look-ahead LU leaves room for the concurrent panel)
    bool dynamic_tiles = false;  // hand tiles out through an atomic counter (full grid even with max_sms: CTAs that become resident late join in)
    bool pdl_dependent = false;  // launch as a programmatic dependent of the previous kernel in the stream (TMA path only)
};
int k_dgemm(int transA, int transB, int m, int n, int k, const double *a, int lda, long long sa, const double *b, int ldb, long long sb,
            double *c, int ldc, long long sc, int batch, const GemmEpilogue &ep);
int gemm_tile_plan(int m, int n, int k, int batch, int sms, bool splitk_allowed, int *ksplit_out);
// lu.cu
int k_getrf(int m, int n, double *a, int lda, long long sa, int32_t *ipiv, long long sp, int32_t *info, int batch);
int k_getrf_streamed(int n, double *a, int lda, int32_t *ipiv, int32_t *info, const std::vector<int> &bounds,
                     const std::function<int(int, int)> &arrive, const std::function<int(int, int, int)> &upper_done,
                     const std::function<int(int, int, int)> &done);
std::vector<int> lu_chunk_bounds(int n, int first, int cap);
std::vector<int> lu_host_bounds(int n);
int k_getrf_nopiv(int n, double *a, int lda, long long sa, int32_t *ipiv, long long sp, int32_t *info, int32_t *viol, int batch);
int k_getrs(int n, int nrhs, const double *lu, int lda, long long sa, const int32_t *ipiv, long long sp, double *b, int ldb, long long sb, int batch);
int k_det_device(int n, const double *lu, int lda, const int32_t *ipiv, double *d_out);
int k_getri(int n, double *a, int lda, const int32_t *ipiv);
int k_first_zero_diag(int n, const double *lu, int lda, int32_t *d_out);  // d_out: device int32, 1-based index or 0
int k_potrf(char uplo, int n, double *a, int lda, int32_t *info);  // info: device int32
// level12.cu
int k_transpose(int m, int n, const double *a, int lda, double *b, int ldb);
int k_tri(int lower, int unit, int m, int n, const double *a, int lda, double *b, int ldb);
int k_pivots_to_matrix(int k, const int32_t *ipiv, double *p, int ldp);
int k_dot_device(long long n, const double *x, long long incx, const double *y, long long incy, double *d_out);
int k_nrm2_device(long long n, const double *x, long long incx, double *d_out);
int k_iamax_device(long long n, const double *x, long long incx, long long *d_out);
int k_swap(long long n, double *x, long long incx, double *y, long long incy);
int k_gemv(int trans, int m, int n, double alpha, const double *a, int lda, const double *x, long long incx, double beta, double *y, long long incy);
int k_larnv(int idist, int32_t *iseed_host, long long n, double *x);
// expm.cu
// one matrix exponential over the devices of a group (mg.cu: fm_mg_expm): every device keeps full copies of the operands, computes
// its column block of every product and stores the finished tiles into every peer's copy (the fused all-gather epilogue of dgemm.cu)
struct ExpmShard {
    int rank = 0, world = 1;
    int j0 = 0, nloc = 0;                         // this device's column block
    const int *device = nullptr;                  // [world] CUDA device of every rank
    double *const (*ws)[6] = nullptr;             // [world][6] the six n x n workspaces (leading dimension n + (n & 1)) of every rank
    std::function<int(int)> barrier;              // barrier(rc): everything queued so far on every device completes before anything queued
                                                  // afterwards on any device starts; returns the group's common verdict
};
int k_expm_workspace(int n, double *ws[6]);
int k_expm_sharded(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *flags_host, const ExpmShard &sh,
                   int *result_ws);
int k_expm(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *info_host);
int k_expm_host_out(int n, const double *a, int lda, double *host_out, int ldh, int mode, double *s_out, int32_t *info_host);
// host(m x n) := dev(m x n) on the download stream, behind the work queued on the compute stream so far and beside what is queued
// next; `last`: the compute stream then waits for the downloads issued so far (their source may be reused afterwards).  api.cu
int download_behind_compute(double *host, size_t ldh, const double *dev, size_t ldd, size_t m, size_t n, bool last);
int k_expm_batched(int n, const double *a, int lda, long long sa, double *out, int ldo, long long so, int batch, int mode, double *s_out);
}  // namespace fm
