// expm.cu -- matrix exponential by diagonal Pade [8/8] with scaling and squaring, the algorithm of
// /root/reference/flomat/expm.rkt:134-215 (NOT Pade-13: the reference uses the n8/d8 coefficient lists, :146-147),
// kept entirely on the device: norm -> scaling exponent -> X = A/2^s -> even/odd polynomials in X^2 -> solve
// (even-odd) R = (even+odd) by LU -> s squarings.  One 8-byte D2H (the 1-norm) is the only host round trip.
//
// mode bit 0:  0 = the reference's Horner order (expm.rkt:167-177) with the two trivial products dropped
//                  (AA*0 and AA*(c*I) are formed exactly by an elementwise pass): 7+s DGEMMs
//              1 = minimum-product regrouping: X^2, X^4, (c6 X^2 + c8 X^4) X^4, (c5 I + c7 X^2) X^4, X*(...): 5+s DGEMMs
// mode bit 1:  0 = reference behaviour for s <= 0 (expm.rkt:205-210 scales A *up* by 2^-s and never squares:
//                  bug-compatible);  1 = clamp s at 0 (mathematically correct result).
// Non-finite norm: the reference's `repeated-squaring` never terminates (F5-iii); we return FM_ERR_NONFINITE.
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace fm {
namespace {

// c_i = (16-i)! 8! / (16! i! (8-i)!)  (expm.rkt:136-146), rounded once from the exact rational
const double kC[9] = {1.0,
                      0.5,
                      0x1.ddddddddddddep-4,
                      0x1.1111111111111p-6,
                      0x1.a41a41a41a41ap-10,
                      0x1.c01c01c01c01cp-14,
                      0x1.45e5d2ba42ea0p-18,
                      0x1.29f6b20961c00p-23,
                      0x1.08db48ebe51c7p-29};

struct Pool {
    void *buf[8] = {};
    size_t bytes[8] = {};
    void *get(int slot, size_t want) {
        if (bytes[slot] >= want) return buf[slot];
        if (buf[slot]) cudaFree(buf[slot]);
        buf[slot] = nullptr;
        bytes[slot] = 0;
        if (cudaMalloc(&buf[slot], want) != cudaSuccess) {
            cudaGetLastError();
            set_error(FM_ERR_NOMEM, "expm workspace allocation of %zu bytes failed", want);
            return nullptr;
        }
        bytes[slot] = want;
        return buf[slot];
    }
};
PerDevice<Pool> g_pool_pd;

// out_z := (s_z is odd ? b1_z : b0_z)   -- collects each item's result from the ping-pong buffer it ended in
__global__ void __launch_bounds__(256) select_copy_kernel(int n, const double *b0, const double *b1, long long ldw, long long sw,
                                                           const int32_t *rounds, double *out, long long ldo, long long so) {
    const int z = blockIdx.z;
    const double *src = ((rounds[z] > 0 ? rounds[z] : 0) & 1) ? b1 : b0;
    src += z * sw;
    out += z * so;
    for (int j = blockIdx.y; j < n; j += gridDim.y)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[j * ldo + i] = src[j * ldw + i];
}

// largest n whose Pade solve uses the factored Neumann product instead of an LU (min-products mode, one matrix; 0 = never)
int neumann_max_n() {
    static const int v = getenv("FLOMAT_EXPM_NEUMANN_MAX_N") ? atoi(getenv("FLOMAT_EXPM_NEUMANN_MAX_N")) : 1024;
    return v;
}
double neumann_bound() {  // developer switch (tests lower it to force the fallback)
    static const double v = getenv("FLOMAT_EXPM_NEUMANN_BOUND") ? atof(getenv("FLOMAT_EXPM_NEUMANN_BOUND")) : 0.3;
    return v;
}
__global__ void flag_unless_below_kernel(const double *x, double bound, int32_t *flag) {
    if (!(*x < bound)) *flag = 1;  // also for NaN
}

double scaling_exponent(double norm1) {
    // (+ 1 (ceiling (log N 2)))  with (log N 2) = log N / log 2   (expm.rkt:214)
    return 1.0 + std::ceil(std::log(norm1) / std::log(2.0));
}

// The whole pipeline for `batch` equally strided matrices.  Workspaces have leading dimension ldw and item stride sw.
// host_out: `out` is HOST memory (batch == 1): the last squaring runs in column blocks and each block is handed to the download stream
// as soon as it is computed, so the device->host copy of the result overlaps the rest of the squaring instead of following it.
// The LU's info words are copied to pinned host memory right behind the factorisation and the rest of the pipeline is queued without
// waiting for them (no bubble between the solve and the squarings); `info_ready` is recorded behind that copy.  The caller checks the
// words afterwards and, in the (practically unreachable) case of an exactly singular Pade denominator, runs the pipeline again with
// `honour_singular`: LAPACK's dgesv leaves B untouched when info > 0 and the reference ignores info (flomat-solve-many!,
// flomat.rkt:2135-2142), so such an item continues with R = num.
struct InfoBack {
    int32_t *h = nullptr;  // pinned, `cap` words
    int cap = 0;
    cudaEvent_t ready = nullptr;
};
PerDevice<InfoBack> g_info_pd;

// `pivot`: factor the denominator with the pivoting LU.  The default first pass uses the verified no-interchange LU (lu.cu,
// k_getrf_nopiv): the Pade denominator is column diagonally dominant (||den - I||_1 <= 0.282 for the reference's scaling), so dgesv
// performs no interchange on it and the pivot search -- the latency chain that dominated expm at small n -- is skipped; the violation
// flag of that factorization travels with the info words, and the caller repeats the pipeline with pivot = true if it is raised.
int expm_pipeline(int n, const double *a, int lda, long long sa, double *out, int ldo, long long so, int batch, int mode, double *s_host,
                  bool host_out = false, bool honour_singular = false, bool pivot = true) {
    const bool minprod = (mode & 1) != 0, clamp = (mode & 2) != 0;
    cudaStream_t st = ctx().stream;
    InfoBack &g_info = g_info_pd();
    const int ldw = n + (n & 1);
    const long long sw = (long long)ldw * n;
    const size_t mat_bytes = (size_t)sw * batch * sizeof(double);
    double *B[6];
    for (int i = 0; i < 6; ++i) {
        B[i] = static_cast<double *>(g_pool_pd().get(i, mat_bytes));
        if (!B[i]) return FM_ERR_NOMEM;
    }
    // small arrays: norms | scales (doubles), rounds | ipiv | info (int32)
    const size_t small_bytes = (size_t)batch * (2 * sizeof(double) + 2 * sizeof(int32_t)) + (size_t)batch * n * sizeof(int32_t) + 128;
    char *small = static_cast<char *>(g_pool_pd().get(6, small_bytes));
    if (!small) return FM_ERR_NOMEM;
    double *d_norm = reinterpret_cast<double *>(small);
    double *d_scale = d_norm + batch;
    int32_t *d_rounds = reinterpret_cast<int32_t *>(d_scale + batch);
    int32_t *d_info = d_rounds + batch;  // batch info words, then ONE violation word of the no-interchange LU
    int32_t *d_viol = d_info + batch;
    int32_t *d_ipiv = d_viol + 1;

    // ---- scaling exponents (expm.rkt:212-215) ----
    FM_TRY(k_norm1_batched(n, a, lda, sa, batch, d_norm));
    std::vector<double> h_norm(batch), h_scale(batch);
    std::vector<int32_t> h_rounds(batch);
    FM_CUDA(cudaMemcpyAsync(h_norm.data(), d_norm, batch * sizeof(double), cudaMemcpyDeviceToHost, st));
    FM_CUDA(cudaStreamSynchronize(st));
    int max_rounds = 0;
    for (int z = 0; z < batch; ++z) {
        const double nz = h_norm[z];
        if (std::isnan(nz) || std::isinf(nz))
            return set_error(FM_ERR_NONFINITE, "expm: ||A||_1 of item %d is %f; the reference does not terminate on this input (expm.rkt:209)", z, nz);
        double s = scaling_exponent(nz);  // -inf for the zero matrix
        if (clamp && !(s > 0.0)) s = 0.0;
        if (s_host) s_host[z] = s;
        h_scale[z] = std::pow(2.0, s);  // (expt 2 s); 0.0 when s = -inf, reproducing the reference's A/0 = NaN
        h_rounds[z] = s > 0.0 ? (int32_t)s : 0;
        if (h_rounds[z] > max_rounds) max_rounds = h_rounds[z];
    }
    FM_CUDA(cudaMemcpyAsync(d_scale, h_scale.data(), batch * sizeof(double), cudaMemcpyHostToDevice, st));
    FM_CUDA(cudaMemcpyAsync(d_rounds, h_rounds.data(), batch * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    // no second sync: a copy FROM pageable host memory returns once the source has been staged, so h_* may go out of scope

    auto gemm = [&](const double *x, const double *y, double *z, double beta, bool has_diag, double diag) {
        GemmEpilogue ep;
        ep.alpha = 1.0;
        ep.beta = beta;
        ep.has_diag = has_diag;
        ep.diag = diag;
        return k_dgemm(FM_NO_TRANS, FM_NO_TRANS, n, n, n, x, ldw, sw, y, ldw, sw, z, ldw, sw, batch, ep);
    };
    // polynomial terms over the contiguous workspaces: o := ci I + c1 a1 + c2 a2, and optionally a second result of the same pass
    auto lin = [&](double ci, double c1, const double *a1, double c2, const double *a2, double *o) {
        return k_pade_terms(n, ldw, batch, a1, a2, ci, c1, c2, o, 0.0, 0.0, 0.0, nullptr);
    };
    auto lin2 = [&](const double *a1, const double *a2, double ci0, double c10, double c20, double *o0, double ci1, double c11, double c21, double *o1) {
        return k_pade_terms(n, ldw, batch, a1, a2, ci0, c10, c20, o0, ci1, c11, c21, o1);
    };

    double *X = B[0], *A2 = B[1];
    FM_TRY(k_div_pow2_batched(n, a, lda, sa, X, ldw, sw, d_scale, batch));   // A/2^s   (expm.rkt:206)
    FM_TRY(gemm(X, X, A2, 0.0, false, 0.0));                                // AA       (expm.rkt:194)
    double *E, *O, *num, *den, *spare;
    if (!minprod) {
        // even = c0 I + AA(c2 I + AA(c4 I + AA(c6 I + AA(c8 I + AA*0))))   (expm.rkt:195 via mpoly :167-177)
        FM_TRY(lin(kC[6], kC[8], A2, 0.0, nullptr, B[2]));         // c6 I + AA*(c8 I), exact elementwise
        FM_TRY(gemm(A2, B[2], B[3], 0.0, true, kC[4]));
        FM_TRY(gemm(A2, B[3], B[2], 0.0, true, kC[2]));
        FM_TRY(gemm(A2, B[2], B[3], 0.0, true, kC[0]));
        E = B[3];
        // odd = X * (c1 I + AA(c3 I + AA(c5 I + AA(c7 I + AA*0))))          (expm.rkt:196)
        FM_TRY(lin(kC[5], kC[7], A2, 0.0, nullptr, B[2]));
        FM_TRY(gemm(A2, B[2], B[4], 0.0, true, kC[3]));
        FM_TRY(gemm(A2, B[4], B[2], 0.0, true, kC[1]));
        FM_TRY(gemm(X, B[2], B[4], 0.0, false, 0.0));
        O = B[4];
        num = B[2];
        den = B[1];
        spare = B[5];
    } else {
        double *A4 = B[2];
        FM_TRY(gemm(A2, A2, A4, 0.0, false, 0.0));
        // one pass over X^2, X^4:  c6 X^2 + c8 X^4  and  c0 I + c2 X^2 + c4 X^4
        FM_TRY(lin2(A2, A4, 0.0, kC[6], kC[8], B[3], kC[0], kC[2], kC[4], B[4]));
        FM_TRY(gemm(B[3], A4, B[4], 1.0, false, 0.0));              // even
        E = B[4];
        // one pass over X^2:  c5 I + c7 X^2  and  c1 I + c3 X^2
        FM_TRY(lin2(A2, nullptr, kC[5], kC[7], 0.0, B[3], kC[1], kC[3], 0.0, B[5]));
        FM_TRY(gemm(B[3], A4, B[5], 1.0, false, 0.0));
        FM_TRY(gemm(X, B[5], B[3], 0.0, false, 0.0));               // odd
        O = B[3];
        num = B[5];
        den = B[1];
        spare = B[2];
    }
    // num = even + odd, den = even - odd, R = den \ num   (expm.rkt:197-199 -> flomat.rkt:2145-2152)
    // the workspaces of a batch are contiguous (item stride = ldw*n), i.e. one ldw x (n*batch) matrix: one launch
    FM_TRY(k_addsub(n, n * batch, E, ldw, O, ldw, num, ldw, den, ldw));
    FM_CUDA(cudaMemsetAsync(d_viol, 0, sizeof(int32_t), st));
    // Small single matrices in the minimal-work mode: R = den^-1 num WITHOUT a factorisation.  den = I - E0 with ||E0||_1 <= 0.282 (see
    // above), so den^-1 = (I + E0)(I + E0^2)(I + E0^4)(I + E0^8)(I + E0^16) up to a relative E0^32 < 3e-18: four squarings of E and five
    // updates R += E_k R, nine products in all -- tensor work only, where the LU's diagonal blocks and triangular sweeps are latency
    // chains (n = 512: 0.18 ms against 0.54 ms).  ||E0||_1 is CHECKED on the device (> 0.3, or NaN, raises the violation word and
    // the caller repeats the pass with the pivoting LU); info stays 0, as dgesv's is for such a matrix.
    const bool neumann = minprod && batch == 1 && !pivot && !honour_singular && n <= neumann_max_n();
    if (neumann) {
        double *Ea = B[3], *Eb = B[4], *T = B[0];
        FM_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
        FM_TRY(lin(1.0, -1.0, den, 0.0, nullptr, Ea));  // E0 = I - den
        FM_TRY(k_norm1_batched(n, Ea, ldw, sw, 1, d_norm));
        flag_unless_below_kernel<<<1, 1, 0, st>>>(d_norm, neumann_bound(), d_viol);
        FM_LAUNCH_CHECK();
        for (int k = 0; k < 5; ++k) {
            FM_TRY(gemm(Ea, num, T, 0.0, false, 0.0));        // T = E_k R
            FM_TRY(k_axpy2d(n, n, 1.0, T, ldw, num, ldw));    // R += T
            if (k < 4) {
                FM_TRY(gemm(Ea, Ea, Eb, 0.0, false, 0.0));    // E_{k+1} = E_k^2
                double *t = Ea; Ea = Eb; Eb = t;
            }
        }
    } else if (pivot) FM_TRY(k_getrf(n, n, den, ldw, sw, d_ipiv, n, d_info, batch));
    else FM_TRY(k_getrf_nopiv(n, den, ldw, sw, d_ipiv, n, d_info, d_viol, batch));
    if (g_info.cap < batch + 1) {
        if (g_info.h) cudaFreeHost(g_info.h);
        g_info.h = nullptr;
        g_info.cap = 0;
        FM_CUDA(cudaMallocHost(&g_info.h, (size_t)(batch < 63 ? 64 : batch + 1) * sizeof(int32_t)));
        g_info.cap = batch < 63 ? 64 : batch + 1;
    }
    if (!g_info.ready) FM_CUDA(cudaEventCreateWithFlags(&g_info.ready, cudaEventDisableTiming));
    FM_CUDA(cudaMemcpyAsync(g_info.h, d_info, (size_t)(batch + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));  // info words + violation word
    FM_CUDA(cudaEventRecord(g_info.ready, st));
    if (neumann) {
        // (solved above)
    } else if (!honour_singular) {
        FM_TRY(k_getrs(n, n, den, ldw, sw, pivot ? d_ipiv : nullptr, n, num, ldw, sw, batch));  // (no interchanges: no row permutation of num)
    } else {
        // second pass after a singular denominator was seen: items with info > 0 keep num (dgesv leaves B untouched)
        FM_CUDA(cudaEventSynchronize(g_info.ready));
        std::vector<int> bad;
        for (int z = 0; z < batch; ++z)
            if (g_info.h[z] != 0) bad.push_back(z);
        if (batch == 1) {
            if (bad.empty()) FM_TRY(k_getrs(n, n, den, ldw, sw, d_ipiv, n, num, ldw, sw, 1));
        } else {
            double *keep = bad.empty() ? nullptr : static_cast<double *>(fm_alloc_bytes((size_t)sw * bad.size() * sizeof(double)));
            if (!bad.empty() && !keep) return FM_ERR_NOMEM;
            for (size_t i = 0; i < bad.size(); ++i) FM_TRY(k_copy2d(keep + i * sw, ldw, num + bad[i] * sw, ldw, n, n));
            FM_TRY(k_getrs(n, n, den, ldw, sw, d_ipiv, n, num, ldw, sw, batch));
            for (size_t i = 0; i < bad.size(); ++i) FM_TRY(k_copy2d(num + bad[i] * sw, ldw, keep + i * sw, ldw, n, n));
            if (keep) fm_free(keep);
        }
    }
    // ---- repeated squaring (expm.rkt:209-210), per-item round counts ----
    double *cur = num, *nxt = spare;
    bool uniform = true;  // every item squares the same number of times (always true for one matrix, the common case for a batch)
    for (int z = 1; z < batch; ++z) uniform = uniform && (h_rounds[z] == h_rounds[0]);
    if (host_out) {  // (batch == 1, hence uniform)
        if (max_rounds == 0) return download_behind_compute(out, ldo, cur, ldw, n, n, true);
        for (int r = 0; r + 1 < max_rounds; ++r) {
            GemmEpilogue ep;
            FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, n, n, n, cur, ldw, sw, cur, ldw, sw, nxt, ldw, sw, 1, ep));
            double *t = cur; cur = nxt; nxt = t;
        }
        int cw = 0, nblocks = 0;
        FM_TRY(fm_dgemm_host_plan(n, n, ctx().sm_count, &cw, &nblocks));
        for (int j = 0; j < nblocks; ++j) {
            const int c0 = j * cw, w = (j == nblocks - 1) ? n - c0 : cw;
            GemmEpilogue ep;
            FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, n, w, n, cur, ldw, 0, cur + (size_t)c0 * ldw, ldw, 0, nxt + (size_t)c0 * ldw, ldw, 0, 1, ep));
            FM_TRY(download_behind_compute(out + (size_t)c0 * ldo, ldo, nxt + (size_t)c0 * ldw, ldw, n, w, j == nblocks - 1));
        }
        return FM_OK;
    }
    if (uniform) {
        if (max_rounds == 0) {
            if (batch == 1) return k_copy2d(out, ldo, cur, ldw, n, n);
            for (int z = 0; z < batch; ++z) FM_TRY(k_copy2d(out + z * so, ldo, cur + z * sw, ldw, n, n));
            return FM_OK;
        }
        for (int r = 0; r < max_rounds; ++r) {  // the last squaring writes straight into the output
            GemmEpilogue ep;
            const bool last = (r == max_rounds - 1);
            FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, n, n, n, cur, ldw, sw, cur, ldw, sw, last ? out : nxt, last ? ldo : ldw, last ? so : sw, batch, ep));
            double *t = cur; cur = nxt; nxt = t;
        }
        return FM_OK;
    }
    for (int r = 0; r < max_rounds; ++r) {
        GemmEpilogue ep;
        ep.active_until = d_rounds;
        ep.round = r;
        FM_TRY(k_dgemm(FM_NO_TRANS, FM_NO_TRANS, n, n, n, cur, ldw, sw, cur, ldw, sw, nxt, ldw, sw, batch, ep));
        double *t = cur; cur = nxt; nxt = t;
    }
    // item z ended in `num` if its round count is even, in `spare` if odd
    dim3 grid(ceil_div(n, 256), n > 256 ? 256 : n, batch);
    select_copy_kernel<<<grid, 256, 0, st>>>(n, num, spare, ldw, sw, d_rounds, out, ldo, so);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

}  // namespace

namespace {
// run the pipeline, then look at the LU's info words (they arrived long before the squarings finish); a singular denominator
// sends the call through the pipeline a second time with dgesv's "B untouched" rule applied.  Returns the first nonzero info.
int expm_checked(int n, const double *a, int lda, long long sa, double *out, int ldo, long long so, int batch, int mode, double *s_out,
                 bool host_out, int32_t *info_host) {
    static const bool always_pivot = getenv("FLOMAT_EXPM_PIVOT") != nullptr;  // developer switch: the pivoting LU from the start (A/B timing)
    // (no-interchange cluster panels with look-ahead exist up to 16384 rows; beyond that the pivoting LU)
    static const int nopiv_max_n = getenv("FLOMAT_EXPM_NOPIV_MAX_N") ? atoi(getenv("FLOMAT_EXPM_NOPIV_MAX_N")) : 16384;
    bool pivot = always_pivot || n > nopiv_max_n;
    FM_TRY(expm_pipeline(n, a, lda, sa, out, ldo, so, batch, mode, s_out, host_out, false, pivot));
    InfoBack &g_info = g_info_pd();
    FM_CUDA(cudaEventSynchronize(g_info.ready));
    if (!pivot && g_info.h[batch] != 0) {
        // the no-interchange factorization was not what partial pivoting computes (never for a finite, nonzero input: see lu.cu): the
        // whole pipeline again with the pivoting LU -- the queued rest of the first pass only wrote workspaces and `out`
        pivot = true;
        FM_TRY(expm_pipeline(n, a, lda, sa, out, ldo, so, batch, mode, s_out, host_out, false, true));
        FM_CUDA(cudaEventSynchronize(g_info.ready));
    }
    int32_t first = 0;
    for (int z = 0; z < batch && first == 0; ++z) first = g_info.h[z];
    if (info_host) *info_host = first;
    if (first == 0) return FM_OK;
    return expm_pipeline(n, a, lda, sa, out, ldo, so, batch, mode, s_out, host_out, true, true);
}
}  // namespace

int k_expm(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *info_host) {
    if (n < 0 || lda < (n > 1 ? n : 1) || ldo < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "expm: bad dimensions");
    if (info_host) *info_host = 0;
    if (n == 0) return FM_OK;
    return expm_checked(n, a, lda, 0, out, ldo, 0, 1, mode, s_out, false, info_host);
}

int k_expm_host_out(int n, const double *a, int lda, double *host_out, int ldh, int mode, double *s_out, int32_t *info_host) {
    if (n < 0 || lda < (n > 1 ? n : 1) || ldh < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "expm: bad dimensions");
    if (info_host) *info_host = 0;
    if (n == 0) return FM_OK;
    return expm_checked(n, a, lda, 0, host_out, ldh, 0, 1, mode, s_out, true, info_host);
}

// ---------------------------------------------------------------------------------------------
// One matrix exponential on `world` devices ("single large expm -- via its dgemms", each product column-sharded and
// all-gathered).  Every device runs this function with the full matrix A in its own memory:
//   * elementwise steps and the norm are REPLICATED (n^2 work, cheaper than moving the result);
//   * every product is SHARDED: rank r computes columns [j0, j0 + nloc) of Z = X Y from its full X and its block of Y, and the GEMM
//     epilogue stores each finished tile into the same place of every peer's Z (the fused all-gather of dgemm.cu; the `c I` term of
//     the reference's Horner steps goes to row j0 + j of local column j: GemmEpilogue::diag_col0);
//   * the LU of the Pade denominator is replicated (no exchange at all), the two triangular solves are sharded by right-hand-side
//     columns, and each device copies its block of the solution to its peers;
//   * a group barrier (ExpmShard::barrier: events across the devices' streams) separates a product from what precedes and follows it.
// out (device memory, may be null): receives the result; *result_ws: the workspace that holds it on every device.
// flags_host[0] = LU info of this device's factorisation, flags_host[1] = the no-interchange LU's violation word; the caller falls
// back to the one-device pipeline when either is set (a singular or non-dominant denominator: not reachable for finite input).
// ---------------------------------------------------------------------------------------------
int k_expm_workspace(int n, double *ws[6]) {
    const int ldw = n + (n & 1);
    for (int i = 0; i < 6; ++i) {
        ws[i] = static_cast<double *>(g_pool_pd().get(i, (size_t)ldw * n * sizeof(double)));
        if (!ws[i]) return FM_ERR_NOMEM;
    }
    return FM_OK;
}

int k_expm_sharded(int n, const double *a, int lda, double *out, int ldo, int mode, double *s_out, int32_t *flags_host, const ExpmShard &sh,
                   int *result_ws) {
    const bool minprod = (mode & 1) != 0, clamp = (mode & 2) != 0;
    cudaStream_t st = ctx().stream;
    const int ldw = n + (n & 1), r = sh.rank;
    double *const *B = sh.ws[r];
    int rc = FM_OK;
    char *small = static_cast<char *>(g_pool_pd().get(6, 2 * sizeof(double) + 4 * sizeof(int32_t) + (size_t)n * sizeof(int32_t) + 128));
    if (!small) rc = FM_ERR_NOMEM;
    double *d_norm = reinterpret_cast<double *>(small), *d_scale = d_norm + 1;
    int32_t *d_info = reinterpret_cast<int32_t *>(d_scale + 1), *d_viol = d_info + 1, *d_ipiv = d_viol + 1;
    double h_norm = 0.0, h_scale = 1.0;
    int rounds = 0;
    if (rc == FM_OK) rc = k_norm1_batched(n, a, lda, 0, 1, d_norm);
    if (rc == FM_OK && cudaMemcpyAsync(&h_norm, d_norm, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "expm (sharded): norm readback failed");
    if (rc == FM_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "expm (sharded): synchronisation failed");
    if (rc == FM_OK) {
        if (std::isnan(h_norm) || std::isinf(h_norm))
            rc = set_error(FM_ERR_NONFINITE, "expm: ||A||_1 is %f; the reference does not terminate on this input (expm.rkt:209)", h_norm);
        else {
            double s = scaling_exponent(h_norm);
            if (clamp && !(s > 0.0)) s = 0.0;
            if (s_out) *s_out = s;
            h_scale = std::pow(2.0, s);
            rounds = s > 0.0 ? (int)s : 0;
            if (cudaMemcpyAsync(d_scale, &h_scale, sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "expm (sharded): scale upload failed");
        }
    }
    // every device has read the same A and therefore found the same s; a failure anywhere ends the call everywhere
    rc = sh.barrier(rc);
    if (rc != FM_OK) return rc;

    const size_t boff = (size_t)sh.j0 * ldw;
    auto gemm = [&](int xi, int yi, int zi, double beta, bool has_diag, double diag) -> int {
        int rc = sh.barrier(FM_OK);  // nobody still reads or writes buffer zi
        if (rc != FM_OK) return rc;
        double *peers[7];
        int np = 0;
        for (int p = 0; p < sh.world; ++p)
            if (p != r) peers[np++] = sh.ws[p][zi] + boff;
        GemmEpilogue ep;
        ep.alpha = 1.0;
        ep.beta = beta;
        ep.has_diag = has_diag;
        ep.diag = diag;
        ep.diag_col0 = sh.j0;
        ep.peers = peers;
        ep.npeers = np;
        if (sh.nloc > 0) rc = k_dgemm(FM_NO_TRANS, FM_NO_TRANS, n, sh.nloc, n, B[xi], ldw, 0, B[yi] + boff, ldw, 0, B[zi] + boff, ldw, 0, 1, ep);
        return sh.barrier(rc);  // every block of zi has arrived on every device
    };
#define FM_STEP(call) do { rc = (call); if (rc != FM_OK) return fail(rc); } while (0)
    // an error between two barriers must not leave the peers waiting at the next one: take part in it with the error, which ends them all
    auto fail = [&](int rc) { sh.barrier(rc); return rc; };
    auto lin = [&](double ci, double c1, const double *a1, double c2, const double *a2, double *o) {
        return k_pade_terms(n, ldw, 1, a1, a2, ci, c1, c2, o, 0.0, 0.0, 0.0, nullptr);
    };
    auto lin2 = [&](const double *a1, const double *a2, double ci0, double c10, double c20, double *o0, double ci1, double c11, double c21, double *o1) {
        return k_pade_terms(n, ldw, 1, a1, a2, ci0, c10, c20, o0, ci1, c11, c21, o1);
    };
    int E, O, num, den, spare;
    FM_STEP(k_div_pow2_batched(n, a, lda, 0, B[0], ldw, 0, d_scale, 1));  // X = A/2^s, full copy on every device
    rc = gemm(0, 0, 1, 0.0, false, 0.0);                                   // AA
    if (rc != FM_OK) return rc;
    if (!minprod) {  // the reference's Horner order (expm.rkt:167-177, 195-196), see expm_pipeline
        FM_STEP(lin(kC[6], kC[8], B[1], 0.0, nullptr, B[2]));
        if ((rc = gemm(1, 2, 3, 0.0, true, kC[4])) != FM_OK) return rc;
        if ((rc = gemm(1, 3, 2, 0.0, true, kC[2])) != FM_OK) return rc;
        if ((rc = gemm(1, 2, 3, 0.0, true, kC[0])) != FM_OK) return rc;
        E = 3;
        FM_STEP(lin(kC[5], kC[7], B[1], 0.0, nullptr, B[2]));
        if ((rc = gemm(1, 2, 4, 0.0, true, kC[3])) != FM_OK) return rc;
        if ((rc = gemm(1, 4, 2, 0.0, true, kC[1])) != FM_OK) return rc;
        if ((rc = gemm(0, 2, 4, 0.0, false, 0.0)) != FM_OK) return rc;
        O = 4; num = 2; den = 1; spare = 5;
    } else {
        if ((rc = gemm(1, 1, 2, 0.0, false, 0.0)) != FM_OK) return rc;  // X^4
        FM_STEP(lin2(B[1], B[2], 0.0, kC[6], kC[8], B[3], kC[0], kC[2], kC[4], B[4]));
        if ((rc = gemm(3, 2, 4, 1.0, false, 0.0)) != FM_OK) return rc;  // even
        E = 4;
        FM_STEP(lin2(B[1], nullptr, kC[5], kC[7], 0.0, B[3], kC[1], kC[3], 0.0, B[5]));
        if ((rc = gemm(3, 2, 5, 1.0, false, 0.0)) != FM_OK) return rc;
        if ((rc = gemm(0, 5, 3, 0.0, false, 0.0)) != FM_OK) return rc;  // odd
        O = 3; num = 5; den = 1; spare = 2;
    }
    FM_STEP(k_addsub(n, n, B[E], ldw, B[O], ldw, B[num], ldw, B[den], ldw));
    if (cudaMemsetAsync(d_viol, 0, sizeof(int32_t), st) != cudaSuccess) return fail(set_error(FM_ERR_CUDA, "expm (sharded): memset failed"));
    static const bool always_pivot = getenv("FLOMAT_EXPM_PIVOT") != nullptr;
    const bool pivot = always_pivot || n > 16384;
    if (pivot) FM_STEP(k_getrf(n, n, B[den], ldw, 0, d_ipiv, n, d_info, 1));  // replicated: every device factors its own copy
    else FM_STEP(k_getrf_nopiv(n, B[den], ldw, 0, d_ipiv, n, d_info, d_viol, 1));
    // info + violation word to this device's pinned words (a copy to pageable memory would hold this thread until the LU has run)
    InfoBack &ib = g_info_pd();
    if (ib.cap < 2) {
        if (ib.h) cudaFreeHost(ib.h);
        ib.h = nullptr;
        ib.cap = 0;
        if (cudaMallocHost(&ib.h, 64 * sizeof(int32_t)) != cudaSuccess) return fail(set_error(FM_ERR_CUDA, "expm (sharded): pinned allocation failed"));
        ib.cap = 64;
    }
    if (cudaMemcpyAsync(ib.h, d_info, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(set_error(FM_ERR_CUDA, "expm (sharded): info readback failed"));
    if (sh.nloc > 0) FM_STEP(k_getrs(n, sh.nloc, B[den], ldw, 0, pivot ? d_ipiv : nullptr, n, B[num] + boff, ldw, 0, 1));  // own right-hand sides
    rc = sh.barrier(FM_OK);  // every device is past its addsub (which wrote all of its `num`) and has its block of the solution
    if (rc != FM_OK) return rc;
    for (int i = 1; i < sh.world && sh.nloc > 0; ++i) {
        const int p = (r + i) % sh.world;
        if (cudaMemcpyPeerAsync(sh.ws[p][num] + boff, sh.device[p], B[num] + boff, sh.device[r], (size_t)ldw * sh.nloc * sizeof(double), st) != cudaSuccess)
            return fail(set_error(FM_ERR_CUDA, "expm (sharded): peer copy of the solution block failed"));
    }
    int cur = num, nxt = spare;
    for (int q = 0; q < rounds; ++q) {  // repeated squaring (expm.rkt:209-210); gemm() opens with the barrier the copies above need
        if ((rc = gemm(cur, cur, nxt, 0.0, false, 0.0)) != FM_OK) return rc;
        const int t = cur; cur = nxt; nxt = t;
    }
    rc = sh.barrier(FM_OK);
    if (rc != FM_OK) return rc;
    if (result_ws) *result_ws = cur;  // every device holds the full result in this workspace
    if (out) rc = k_copy2d(out, ldo, B[cur], ldw, n, n);
    if (rc == FM_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = set_error(FM_ERR_CUDA, "expm (sharded): synchronisation failed");
    flags_host[0] = ib.h[0];
    flags_host[1] = ib.h[1];
    return rc;
#undef FM_STEP
}

int k_expm_batched(int n, const double *a, int lda, long long sa, double *out, int ldo, long long so, int batch, int mode, double *s_out) {
    if (n < 0 || batch < 0 || lda < (n > 1 ? n : 1) || ldo < (n > 1 ? n : 1)) return set_error(FM_ERR_ARG, "expm_batched: bad dimensions");
    if (n == 0 || batch == 0) return FM_OK;
    if (batch > 65535) return set_error(FM_ERR_ARG, "expm_batched: at most 65535 items per call");
    return expm_checked(n, a, lda, sa, out, ldo, so, batch, mode, s_out, false, nullptr);
}

}  // namespace fm
