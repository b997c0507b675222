/* flomat_cuda.h -- C ABI of libflomat_cuda, the B200 (sm_100a) backend for flomat's dense FP64 hot path.
 *
 * This is the drop-in boundary: the symbols below are what the reference's FFI binding layer
 * (`define-cblas` / `define-lapack`, /root/reference/flomat/flomat.rkt:139-140) binds for this path, plus the
 * device-resident tier that replaces the Racket-side per-element loops and GC `malloc` buffers.
 * Plain C: pointers, 32-bit ints, doubles.  No torch / C++ types cross this boundary.
 *
 * Conventions
 *   - matrices are column-major, element (i,j) at a[i + j*lda]  (flomat.rkt:288-300), all dims 32-bit `int`
 *   - LAPACK-style entry points take scalars by reference and report through `info`
 *     (0 ok, >0 first exactly-zero pivot (1-based), <0 = -(index of the bad argument))
 *   - ipiv is int32, 1-based: row i was swapped with row ipiv[i]-1  (flomat.rkt:1457-1463)
 *   - fm_* functions return 0 on success or a negative FM_ERR_* code; fm_last_error() gives the message.
 *     CUDA failures never abort the process and never fall back to a CPU path.
 *   - device-tier calls are asynchronous on the library stream (fm_set_stream / default: legacy stream 0)
 *     unless they return a scalar to the host.
 *   - threading: the reference calls from ONE OS thread.  Calls that target the same device must be serialised by the
 *     caller; fm_last_error() is thread-local.
 */
#ifndef FLOMAT_CUDA_H
#define FLOMAT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM_OK 0
#define FM_ERR_CUDA (-1001)    /* a CUDA runtime/driver call failed (see fm_last_error) */
#define FM_ERR_ARG (-1002)     /* invalid argument */
#define FM_ERR_NOGPU (-1003)   /* no usable sm_100 device */
#define FM_ERR_NOMEM (-1004)   /* device allocation failed */
#define FM_ERR_NONFINITE (-1005) /* expm: norm is inf/NaN (the reference would not terminate, expm.rkt:209) */

/* CBLAS enums as flomat.rkt:846-853 passes them */
#define FM_COL_MAJOR 102
#define FM_NO_TRANS 111
#define FM_TRANS 112
/* fm_ewise ops / fm_map functions */
#define FM_OP_ADD 0
#define FM_OP_SUB 1
#define FM_OP_MUL 2
#define FM_OP_DIV 3
#define FM_OP_EXPT 4
#define FM_MAP_SIN 5
#define FM_MAP_COS 6
#define FM_MAP_TAN 7
#define FM_MAP_SQR 8
#define FM_MAP_SQRT 9
#define FM_MAP_LOG 10
#define FM_MAP_EXP 11

/* ------------------------------------------------------------------------------------------------
 * Tier 1: compat symbols -- exact names and C signatures implied by the reference `_fun` types.
 * Operands may be HOST or DEVICE pointers (decided per pointer with cudaPointerGetAttributes):
 * host operands are staged through HBM inside the call, device operands run in place.
 * ---------------------------------------------------------------------------------------------- */

/* flomat.rkt:855-875, call site :888 (always order=102) */
void cblas_dgemm(int order, int transA, int transB, int M, int N, int K, double alpha, const double *A, int lda,
                 const double *B, int ldb, double beta, double *C, int ldc);
/* flomat.rkt:751-757, call sites :782 :797 :810 */
void cblas_daxpy(int n, double alpha, const double *X, int incX, double *Y, int incY);
/* flomat.rkt:483-488, call sites :500 and :544 (incX = 0 broadcast fill is relied upon) */
void cblas_dcopy(int n, const double *X, int incX, double *Y, int incY);
/* flomat.rkt:1014-1018, call sites :1026 :1029 */
void cblas_dscal(int n, double alpha, double *X, int incX);
/* flomat.rkt:1317-1324, call site :1340; norm in {'M','1','O','I','F','E'}; work may be NULL */
double dlange_(const char *norm, const int *m, const int *n, const double *a, const int *lda, double *work);
/* flomat.rkt:1505-1523, call site :1527; ipiv has min(m,n) entries (the reference allocates m) */
void dgetrf_(const int *m, const int *n, double *a, const int *lda, int32_t *ipiv, int *info);
/* flomat.rkt:2112-2124, call sites :2131 :2140 */
void dgesv_(const int *n, const int *nrhs, double *a, const int *lda, int32_t *ipiv, double *b, const int *ldb,
            int *info);
/* flomat.rkt:1754-1768, call site :1778; work/lwork are accepted and ignored (workspace is device-side) */
void dgetri_(const int *n, double *a, const int *lda, const int32_t *ipiv, double *work, const int *lwork, int *info);
/* --- neighbours of the hot path: same host-or-device convention --- */
/* flomat.rkt:934-947, call site :949 (flomat*vector :957-978); order is always 102 */
void cblas_dgemv(int order, int transA, int M, int N, double alpha, const double *A, int lda, const double *X, int incX,
                 double beta, double *Y, int incY);
/* flomat.rkt:1916-1921, call sites :1924 and `unsafe-sum` :2074 (incY = 0 against a single 1.0) */
double cblas_ddot(int n, const double *X, int incX, const double *Y, int incY);
/* flomat.rkt:1299-1302, call site :1943 */
double cblas_dnrm2(int n, const double *X, int incX);
/* flomat.rkt:1156-1160, call sites :1167-1176; 0-based index of the first largest |x| (CBLAS convention) */
int cblas_idamax(int n, const double *X, int incX);
/* flomat.rkt:1074-1079, call sites :1088 (rows, inc = lda) and :1103 (columns) */
void cblas_dswap(int n, double *X, int incX, double *Y, int incY);
/* flomat.rkt:1794-1809, call site :1818 (flomat-cholesky!); the triangle that is not selected by uplo is left untouched */
void dpotrf_(const char *uplo, const int *n, double *a, const int *lda, int *info);
/* flomat.rkt:2798-2803, call sites :2818-2850 (`rand` idist 1, `randn` idist 3); iseed: 4 ints in 0..4095, updated */
void dlarnv_(const int *idist, int32_t *iseed, const int *n, double *x);

/* ------------------------------------------------------------------------------------------------
 * Tier 2: device-resident flomat buffers (replaces alloc-flomat :437, copy-flomat :516, make-flomat :538,
 * unsafe-ref/unsafe-set! :986-1002 and the pointwise macros :2991-3146 of flomat.rkt)
 * ---------------------------------------------------------------------------------------------- */

/* library / device */
int fm_init(int device);                 /* idempotent; binds the CALLING THREAD to the device and creates its context on first use;
                                            a process may use several devices (one context each), a thread one at a time */
int fm_device_count(void);
int fm_set_stream(void *cuda_stream);    /* cudaStream_t; NULL = legacy default stream */
int fm_sync(void);                       /* wait for everything queued on the library stream */
const char *fm_last_error(void);         /* thread-local, never NULL */
const char *fm_version(void);
uint64_t fm_kernel_launches(void);       /* number of kernels this library has launched so far (all devices) */

/* memory: alloc-flomat / free; upload & download are the explicit host sync points */
double *fm_alloc(int m, int n);                          /* m*n doubles of HBM, lda = m; NULL on failure */
void *fm_alloc_bytes(size_t bytes);
int fm_free(void *dev);                   /* returns the block to the library's cache (no device sync) */
int fm_trim_cache(void);                  /* cudaFree every cached block */
int fm_upload(double *dev, int ldd, const double *host, int ldh, int m, int n);   /* H2D, strided */
int fm_download(double *host, int ldh, const double *dev, int ldd, int m, int n); /* D2H, strided */
/* asynchronous variants on the library's transfer stream (no reference counterpart: the reference has no device).  An upload may
 * overlap kernels queued before it and is visible to everything queued after it; a download waits for the work queued before it and
 * overlaps what is queued after it -- the host buffer is valid after fm_transfer_sync() / fm_sync().  Use pinned host memory. */
int fm_upload_async(double *dev, int ldd, const double *host, int ldh, int m, int n);
int fm_download_async(double *host, int ldh, const double *dev, int ldd, int m, int n);
int fm_transfer_sync(void);
int fm_get(const double *dev, size_t index, double *out);  /* unsafe-ref on a device buffer (syncs) */
int fm_set(double *dev, size_t index, double value);       /* unsafe-set! */
int fm_upload_i32(int32_t *dev, const int32_t *host, size_t n);
int fm_download_i32(int32_t *host, const int32_t *dev, size_t n);

/* fills and copies: make-flomat :538 (memset / dcopy incX=0), flomat-eye :2642, unsafe-matrix-copy! :502 */
int fm_fill(double *a, int lda, int m, int n, double x);
int fm_eye(double *a, int lda, int m, int n, int k);      /* zeros with ones on the k-th diagonal */
int fm_copy2d(double *dst, int ldd, const double *src, int lds, int m, int n);

/* pointwise macros .+ .- .* ./ (flomat.rkt:3021-3146): C := op(A|sa, B|sb); a NULL matrix pointer means
 * "use the scalar".  op: 0 add, 1 sub, 2 mul, 3 div.  C may alias A or B (the `!` forms). */
int fm_ewise(int op, int m, int n, const double *a, int lda, double sa, const double *b, int ldb, double sb, double *c,
             int ldc);
/* constant*flomat+flomat! (flomat.rkt:805): B := alpha*A + B for a whole matrix (replaces n daxpy calls) */
int fm_axpy2d(int m, int n, double alpha, const double *a, int lda, double *b, int ldb);
/* constant*matrix! (flomat.rkt:1020): A := s*A */
int fm_scal2d(int m, int n, double s, double *a, int lda);
/* fused expm helper: num := E + O, den := E - O in one pass (expm.rkt:197-198) */
int fm_addsub(int m, int n, const double *e, int lde, const double *o, int ldo, double *num, int ldn, double *den,
              int ldd);

/* flomat-norm (flomat.rkt:1328): type '1' 'I' 'F' 'M'; result to host (syncs) */
int fm_norm(char type, int m, int n, const double *a, int lda, double *out);

/* C := alpha*op(A)*op(B) + beta*C, all pointers device (flomat.rkt:877-896).  beta == 0 never reads C. */
int fm_dgemm(int transA, int transB, int m, int n, int k, double alpha, const double *a, int lda, const double *b,
             int ldb, double beta, double *c, int ldc);
/* C := alpha*A*B + beta*C (NN) with all three operands in HOST memory -- what the unmodified reference hands to cblas_dgemm
 * (flomat.rkt:888).  A goes up once; B and C travel in column blocks so that the upload of block j+1, the product of block j and
 * the download of block j-1 overlap.  cblas_dgemm takes this path for host operands with N >= 1024 and M*N*K >= 2^30.
 * async != 0: returns without waiting, C is valid after fm_transfer_sync() / fm_sync(); use pinned host memory for a real overlap. */
int fm_dgemm_host(int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta,
                  double *c, int ldc, int async);
/* the column blocks fm_dgemm_host cuts an m x n result into on a device with `sms` SMs (<= 0: 148): block width (a multiple of 128)
 * and block count, the last block taking the remainder.  Host arithmetic only; callable without a GPU. */
int fm_dgemm_host_plan(int m, int n, int sms, int *width, int *nblocks);
/* which configuration of the GEMM kernel an aligned NN product takes on `sms` SMs: *tile_rows = 128 (128x128 tiles, one CTA per SM) or 64
 * (64x128 tiles, two CTAs per SM), *ksplit = k slices (1 = none); and the column chunks of the host-operand LU pipeline of dgesv_ /
 * dgetrf_ (bounds[0] = 0 < ... < bounds[*nchunks] = n).  Pure host arithmetic: callable without a GPU. */
int fm_dgemm_tile_plan(int m, int n, int k, int batch, int sms, int *tile_rows, int *ksplit);
int fm_lu_host_plan(int n, int *bounds, int max_bounds, int *nchunks);
/* as fm_dgemm (NN only) and then adds diag to C[i,i]: the `(plus (.* c I) (times A P))` step of mpoly, expm.rkt:176 */
int fm_dgemm_diag(int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta,
                  double *c, int ldc, double diag);
/* batched NN gemm over `batch` equally strided items (strides in doubles) */
int fm_dgemm_batched(int m, int n, int k, double alpha, const double *a, int lda, long long stride_a, const double *b,
                     int ldb, long long stride_b, double beta, double *c, int ldc, long long stride_c, double diag,
                     int batch);

/* LU family, device pointers (a, b, ipiv, info all in HBM; info is one int32 per matrix) */
int fm_getrf(int m, int n, double *a, int lda, int32_t *ipiv, int32_t *info);
/* LU WITHOUT interchanges, verified (the solve inside expm uses it: its Pade denominator is column diagonally dominant, so dgetrf
 * performs no interchange on it): a := L\U, ipiv := 1..n, info := 0, and *violated (device int32) := 1 if some multiplier exceeds
 * 0.999 in magnitude or a pivot is zero / not finite -- i.e. if partial pivoting WOULD have interchanged -- in which case the result
 * must be discarded and fm_getrf used on the original matrix.  When *violated == 0 the factors are dgetrf's (same pivots: none). */
int fm_getrf_nopiv(int n, double *a, int lda, int32_t *ipiv, int32_t *info, int32_t *violated);
/* The factorisation dgesv_ / dgetrf_ run on a HOST matrix of n >= 6144 (columns arrive in chunks of `chunk` columns, a multiple of
 * 128, while earlier chunks are factored; left-looking over the chunks), here on a matrix already in HBM: for tests and timing.
 * Same result as fm_getrf up to the summation order of the Schur updates. */
int fm_getrf_chunked(int n, double *a, int lda, int32_t *ipiv, int32_t *info, int chunk);
int fm_getrs(int n, int nrhs, const double *lu, int lda, const int32_t *ipiv, double *b, int ldb);
/* dgesv semantics (flomat.rkt:2112): with info > 0 (exactly singular U) the solve is skipped and B is left untouched, which is
 * what flomat-solve-many! (:2135-2142, info ignored) then returns.  Reads info back once (4 bytes, syncs). */
int fm_gesv(int n, int nrhs, double *a, int lda, int32_t *ipiv, double *b, int ldb, int32_t *info);
/* a := inv from its LU; info (device, may be NULL) = i if U(i,i) is exactly zero, in which case a keeps the LU factors
 * (dgetri/dtrtri semantics, flomat.rkt:1754-1779).  Reads the scan result back once (syncs). */
int fm_getri(int n, double *a, int lda, const int32_t *ipiv, int32_t *info);
/* flomat-determinant-from-plu (flomat.rkt:1837): sign(ipiv) * left-to-right product of diag(LU); syncs */
int fm_det(int n, const double *lu, int lda, const int32_t *ipiv, double *out);

/* ---- neighbours of the hot path kept on the device ---- */
/* B (n x m) := A' ; flomat-transpose flomat.rkt:1360-1369 (a Racket loop in the reference) */
int fm_transpose(int m, int n, const double *a, int lda, double *b, int ldb);
/* B (m x n) := triangle of A, zeros outside, optionally ones on the diagonal:
 * flomat-extract-upper :1540 (lower=0), flomat-extract-lower :1566 (lower=1), flomat-extract-lower/ones :1553 (unit=1) */
int fm_tri(int lower, int unit_diag, int m, int n, const double *a, int lda, double *b, int ldb);
/* P (k x k) := the permutation matrix of an LU's pivots; pivots->flomat flomat.rkt:1471-1484 (ipiv 1-based, device) */
int fm_pivots_to_matrix(int k, const int32_t *ipiv, double *p, int ldp);
/* C := f(A) elementwise, fn = FM_MAP_*: the `.sin .cos .tan .sqr .sqrt .log .exp` family, flomat.rkt:2990-3019;
 * `.expt` is fm_ewise with FM_OP_EXPT (flomat.rkt:3076) */
int fm_map(int fn, int m, int n, const double *a, int lda, double *c, int ldc);
/* y := alpha op(A) x + beta y (device operands); constant*matrix*vector+constant*vector! flomat.rkt:949-959 */
int fm_gemv(int trans, int m, int n, double alpha, const double *a, int lda, const double *x, int incx, double beta,
            double *y, int incy);
int fm_dot(int n, const double *x, int incx, const double *y, int incy, double *out);   /* cblas_ddot, syncs */
int fm_nrm2(int n, const double *x, int incx, double *out);                             /* cblas_dnrm2, syncs */
int fm_iamax(int n, const double *x, int incx, int *out);                               /* cblas_idamax (0-based), syncs */
int fm_swap(int n, double *x, int incx, double *y, int incy);                           /* cblas_dswap */
/* out[j] = sum_i A(i,j) / out[i] = sum_j A(i,j): flomat-column-sums :2094, flomat-row-sums :2101 (n or m ddots there) */
int fm_colsums(int m, int n, const double *a, int lda, double *out);
int fm_rowsums(int m, int n, const double *a, int lda, double *out);
/* Cholesky in place (device operands, info = device int32): uplo 'L' A = L L^T / 'U' A = U^T U; flomat-cholesky! :1815 */
int fm_potrf(char uplo, int n, double *a, int lda, int32_t *info);
/* dlarnv on the device: x[0..n) from LAPACK's 48-bit multiplicative generator (bit-identical uniforms; normals by the same
 * Box-Muller pairing, within a few ulp of libm); iseed (host, 4 ints in 0..4095) is advanced exactly like DLARNV does */
int fm_larnv(int idist, int32_t *iseed, long long n, double *x);

/* expm (expm.rkt:203): out := exp(A) by Pade [8/8] scaling-and-squaring, everything on device.
 * *s_out receives the scaling exponent 1+ceil(log2 ||A||_1) the reference computes (may be <= 0).
 * mode 0: reference Horner order (7+s products after dropping the two products with 0 and c*I);
 * mode 1: minimum-product regrouping (5+s products). */
/* `out` may also be HOST memory: the last squaring then runs in column blocks whose device->host copies overlap the blocks
 * still to be computed; a pinned buffer is valid after fm_transfer_sync() / fm_sync(), a pageable one on return. */
/* *info_host (may be NULL) receives the info of the inner dgesv (0, or i for an exactly singular Pade denominator, in which case the
 * result is what the reference computes: dgesv leaves num untouched and the squarings run on it, flomat.rkt:2135-2142). */
int fm_expm(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *info_host);
/* batch of equally strided n x n matrices; s_out (host, `batch` doubles) may be NULL */
int fm_expm_batched(int n, const double *a, int lda, long long stride_a, double *out, int ldo, long long stride_o,
                    int batch, int mode, double *s_out);

/* ------------------------------------------------------------------------------------------------
 * Tier 3: multi-GPU (one process per GPU).  Column-sharded times: every rank holds all of A and the
 * column block [j0, j0+nloc) of B, and C is reassembled on every rank.  With peer pointers the GEMM
 * epilogue itself writes each finished C tile into every peer's copy of C over NVLink (no separate
 * all-gather); without them the caller all-gathers with NCCL.
 * ---------------------------------------------------------------------------------------------- */
#define FM_IPC_HANDLE_BYTES 64
int fm_ipc_export(void *dev, void *handle_out /* FM_IPC_HANDLE_BYTES */, size_t *offset_out);
int fm_ipc_open(const void *handle, size_t offset, void **dev_out);
int fm_ipc_close(void *dev);
/* C[:, j0:j0+nloc] := A * Bloc, stored to the local C and to every peer_c[r] (r < npeers, may be 0) */
int fm_dgemm_colshard(int m, int nloc, int k, const double *a, int lda, const double *bloc, int ldb, double *c, int ldc,
                      int j0, double *const *peer_c, int npeers);

/* ------------------------------------------------------------------------------------------------
 * Tier 3b: multi-GPU from ONE process and ONE calling thread -- the form the reference can use: `flomat*` is one call from one
 * Racket OS thread (flomat.rkt:905-914 -> :888).  fm_mg_init starts one worker thread per device (rank 0 = the caller's device)
 * and enables peer access between all pairs of the group (<= 8 devices of one NVSwitch domain).
 *
 * fm_mg_dgemm: C := alpha A B + beta C sharded by column blocks of B and C, A replicated.  Operands are either
 *   all HOST memory  -- each device uploads its own slice of A over its own PCIe link, the slices are all-gathered over NVLink,
 *                       and block j of B / C runs through device j's upload | product | download pipeline; returns when C is complete;
 *   all DEVICE memory on the caller's device -- the other devices pull A and their block of B over NVLink and their GEMM epilogues
 *                       store the finished tiles of C directly into the caller's C (fused peer-store epilogue); asynchronous: the
 *                       caller's library stream waits for one event per device (fm_sync / any later fm_* call sees the result).
 * Only the untransposed product is sharded; trans flags other than 111/111 are rejected (use fm_dgemm).
 * fm_mg_expm_batched: as fm_expm_batched with the items split into one contiguous block per device (no exchange); host operands
 *   or device operands on the caller's device.  s_out (host, `batch` doubles, may be NULL).
 * fm_mg_expm: ONE matrix exponential (`expm`, expm.rkt:203) over the whole group (SURVEY 8e "single large expm ... via its dgemms"):
 *   every device holds full copies of the operands; each of the 5+s / 11+s products is sharded by column blocks and its tiles are
 *   stored into every peer's copy by the GEMM epilogue (fused all-gather); the Pade denominator's LU is replicated, the triangular
 *   solves are sharded by right-hand-side columns; elementwise steps are replicated.  a and out: both host memory (each device moves
 *   its column block over its own PCIe link) or both on the calling thread's device.  Blocks until `out` is complete.  Same results
 *   as fm_expm up to the tile shapes of the products (<= 64 n eps, tested); n < 128 * devices runs on one device.
 * LU / solve / inv / det are single-GPU (north_star).
 * ---------------------------------------------------------------------------------------------- */
int fm_mg_init(int ndev);                 /* <= 0: all visible devices (at most 8); idempotent for the same count */
int fm_mg_device_count(void);             /* devices of the group, 0 before fm_mg_init */
int fm_mg_dgemm(int transA, int transB, int m, int n, int k, double alpha, const double *a, int lda, const double *b,
                int ldb, double beta, double *c, int ldc);
int fm_mg_expm_batched(int n, const double *a, int lda, long long stride_a, double *out, int ldo, long long stride_o,
                       int batch, int mode, double *s_out);
int fm_mg_expm(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *info);
int fm_mg_sync(void);                     /* wait for every device of the group */
/* the partitions the two calls above use: columns in whole 128-column tile columns, items in contiguous blocks (host arithmetic only) */
int fm_mg_column_block(int n, int world, int r, int *first, int *count);
int fm_mg_item_block(int batch, int world, int r, int *first, int *count);

#ifdef __cplusplus
}
#endif
#endif /* FLOMAT_CUDA_H */
