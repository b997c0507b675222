// norm.cu -- dlange_ replacement (flomat.rkt:1317-1340): '1' max column abs-sum (drives expm's scaling exponent,
// expm.rkt:213), 'I' max row abs-sum, 'F' Frobenius (scaled, overflow-safe like DLASSQ), 'M' max |a_ij|.
// HBM-bound: 8 algorithmic bytes per element, one pass ('F': two passes, the first finds the scale).
// All reductions are deterministic (fixed tree shapes, no floating-point atomics); NaNs propagate as in LAPACK.
#include "common.cuh"
#include <cmath>

namespace fm {
namespace {

constexpr int kT = 256;

__device__ __forceinline__ double nan_max(double a, double b) {
    // LAPACK: "if (value < temp || isnan(temp)) value = temp"
    if (isnan(a)) return a;
    if (isnan(b)) return b;
    return a < b ? b : a;
}

template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v) {
    __shared__ double sh[kT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = IS_MAX ? nan_max(v, w) : v + w;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < kT / 32 ? sh[lane] : (IS_MAX ? 0.0 : 0.0);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = IS_MAX ? nan_max(v, w) : v + w;
        }
    }
    return v;  // valid in thread 0
}

// one CTA per column (grid-stride): colsum[j] = sum_i |a_ij|; batched over blockIdx.y
__global__ void __launch_bounds__(kT) colsum_kernel(int m, int n, const double *a, long long lda, long long stride, double *colsum) {
    a += blockIdx.y * stride;
    colsum += (long long)blockIdx.y * n;
    for (int j = blockIdx.x; j < n; j += gridDim.x) {
        const double *col = a + j * lda;
        double s = 0.0;
        for (int i = threadIdx.x; i < m; i += kT) s += fabs(col[i]);
        s = block_reduce<false>(s);
        if (threadIdx.x == 0) colsum[j] = s;
    }
}

// One warp per column, 16-byte loads: for many columns (the batched 256x256 items of expm, or any n >= 2048) a CTA per column
// spends its time in the block reduction and keeps one load per thread in flight; a warp streams its column with four
// independent 16-byte loads per lane and reduces with shuffles only.
__global__ void __launch_bounds__(kT) colsum_warp_kernel(int m, int n, long long total, const double *a, long long lda, long long stride,
                                                         double *colsum) {
    const long long w = (long long)blockIdx.x * (kT / 32) + (threadIdx.x >> 5);
    if (w >= total) return;
    const int lane = threadIdx.x & 31;
    const long long z = w / n;
    const double *col = a + z * stride + (w - z * n) * lda;
    double s = 0.0;
    if ((reinterpret_cast<uintptr_t>(col) & 15u) == 0) {
        const int m2 = m >> 1;
        const double2 *c2 = reinterpret_cast<const double2 *>(col);
#pragma unroll 4
        for (int i = lane; i < m2; i += 32) {
            const double2 v = c2[i];
            s += fabs(v.x) + fabs(v.y);
        }
        if ((m & 1) && lane == 0) s += fabs(col[m - 1]);
    } else {
#pragma unroll 4
        for (int i = lane; i < m; i += 32) s += fabs(col[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) colsum[w] = s;
}

// out[z] = max_j v[z*n + j] with NaN propagation; one CTA per batch item
__global__ void __launch_bounds__(kT) max_kernel(int n, const double *v, double *out) {
    v += (long long)blockIdx.x * n;
    double r = 0.0;
    for (int j = threadIdx.x; j < n; j += kT) r = nan_max(r, v[j]);
    r = block_reduce<true>(r);
    if (threadIdx.x == 0) out[blockIdx.x] = r;
}

// partial row sums: part[c*m + i] = sum over the c-th chunk of columns of |a_ij|
__global__ void __launch_bounds__(kT) rowsum_partial_kernel(int m, int n, const double *a, long long lda, int cols_per_chunk, double *part) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= m) return;
    const int j0 = blockIdx.y * cols_per_chunk;
    const int j1 = min(n, j0 + cols_per_chunk);
    double s = 0.0;
    for (int j = j0; j < j1; ++j) s += fabs(a[j * lda + i]);
    part[(long long)blockIdx.y * m + i] = s;
}
__global__ void __launch_bounds__(kT) rowsum_final_kernel(int m, int chunks, const double *part, double *rows) {
    const int i = blockIdx.x * kT + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += part[(long long)c * m + i];
    rows[i] = s;
}

// per-CTA partial of max|a| (MODE 0) or sum (a/scale)^2 (MODE 1)
template <int MODE>
__global__ void __launch_bounds__(kT) elem_partial_kernel(int m, int n, const double *a, long long lda, const double *scale, double *part) {
    double r = 0.0;
    const double inv = (MODE == 1) ? *scale : 1.0;
    for (int j = blockIdx.y; j < n; j += gridDim.y)
        for (int i = blockIdx.x * kT + threadIdx.x; i < m; i += gridDim.x * kT) {
            const double v = fabs(a[j * lda + i]);
            if (MODE == 0) r = nan_max(r, v);
            else { const double q = v / inv; r += q * q; }
        }
    r = block_reduce<MODE == 0>(r);
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = r;
}
__global__ void __launch_bounds__(kT) sum_kernel(int n, const double *v, double *out) {
    double r = 0.0;
    for (int j = threadIdx.x; j < n; j += kT) r += v[j];
    r = block_reduce<false>(r);
    if (threadIdx.x == 0) *out = r;
}
__global__ void frob_finish_kernel(const double *scale, const double *sumsq, double *out) {
    const double s = *scale;
    *out = (s == 0.0) ? 0.0 : (isnan(s) || isinf(s) ? s : s * sqrt(*sumsq));
}

}  // namespace

int k_norm1_batched(int n, const double *a, int lda, long long stride, int batch, double *d_out) {
    if (n <= 0 || batch <= 0) return FM_OK;
    double *colsum = static_cast<double *>(scratch((size_t)n * batch * sizeof(double)));
    if (!colsum) return FM_ERR_NOMEM;
    const long long total = (long long)n * batch;
    if (total >= 2048) {
        colsum_warp_kernel<<<(unsigned)((total + kT / 32 - 1) / (kT / 32)), kT, 0, ctx().stream>>>(n, n, total, a, lda, stride, colsum);
    } else {
        dim3 grid(n, batch);
        colsum_kernel<<<grid, kT, 0, ctx().stream>>>(n, n, a, lda, stride, colsum);
    }
    FM_LAUNCH_CHECK();
    max_kernel<<<batch, kT, 0, ctx().stream>>>(n, colsum, d_out);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int k_norm_device(char type, int m, int n, const double *a, int lda, double *d_out) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1)) return set_error(FM_ERR_ARG, "fm_norm: bad dimensions");
    cudaStream_t st = ctx().stream;
    if (m == 0 || n == 0) {
        FM_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double), st));
        return FM_OK;
    }
    switch (type) {
        case '1': case 'O': case 'o': {
            double *colsum = static_cast<double *>(scratch((size_t)n * sizeof(double)));
            if (!colsum) return FM_ERR_NOMEM;
            if (n >= 2048) colsum_warp_kernel<<<(unsigned)((n + kT / 32 - 1) / (kT / 32)), kT, 0, st>>>(m, n, n, a, lda, 0, colsum);
            else colsum_kernel<<<dim3(n, 1), kT, 0, st>>>(m, n, a, lda, 0, colsum);
            FM_LAUNCH_CHECK();
            max_kernel<<<1, kT, 0, st>>>(n, colsum, d_out);
            FM_LAUNCH_CHECK();
            return FM_OK;
        }
        case 'I': case 'i': {
            int chunks = n >= 64 ? 32 : 1;
            const int cpc = ceil_div(n, chunks);
            chunks = ceil_div(n, cpc);
            double *part = static_cast<double *>(scratch(((size_t)chunks + 1) * m * sizeof(double)));
            if (!part) return FM_ERR_NOMEM;
            double *rows = part + (size_t)chunks * m;
            rowsum_partial_kernel<<<dim3(ceil_div(m, kT), chunks), kT, 0, st>>>(m, n, a, lda, cpc, part);
            FM_LAUNCH_CHECK();
            rowsum_final_kernel<<<ceil_div(m, kT), kT, 0, st>>>(m, chunks, part, rows);
            FM_LAUNCH_CHECK();
            max_kernel<<<1, kT, 0, st>>>(m, rows, d_out);
            FM_LAUNCH_CHECK();
            return FM_OK;
        }
        case 'M': case 'm': case 'F': case 'f': case 'E': case 'e': {
            int gx = ceil_div(m, kT); if (gx > 64) gx = 64;
            int gy = n; if (gy > 64) gy = 64;
            double *part = static_cast<double *>(scratch(((size_t)gx * gy + 2) * sizeof(double)));
            if (!part) return FM_ERR_NOMEM;
            double *scale = part + (size_t)gx * gy, *sumsq = scale + 1;
            const bool is_max = (type == 'M' || type == 'm');
            elem_partial_kernel<0><<<dim3(gx, gy), kT, 0, st>>>(m, n, a, lda, nullptr, part);
            FM_LAUNCH_CHECK();
            max_kernel<<<1, kT, 0, st>>>(gx * gy, part, is_max ? d_out : scale);
            FM_LAUNCH_CHECK();
            if (is_max) return FM_OK;
            elem_partial_kernel<1><<<dim3(gx, gy), kT, 0, st>>>(m, n, a, lda, scale, part);
            FM_LAUNCH_CHECK();
            sum_kernel<<<1, kT, 0, st>>>(gx * gy, part, sumsq);
            FM_LAUNCH_CHECK();
            frob_finish_kernel<<<1, 1, 0, st>>>(scale, sumsq, d_out);
            FM_LAUNCH_CHECK();
            return FM_OK;
        }
        default:
            return set_error(FM_ERR_ARG, "fm_norm: unknown norm type '%c'", type);
    }
}

}  // namespace fm
