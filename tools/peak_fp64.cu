// FP64 issue-rate microbenchmark for B200 (sm_100a): DMMA (mma.sync f64) vs DFMA.
// Measures the roofline denominator for every dgemm-shaped kernel in this repo.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peak_fp64 peak_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma16816(double *out, int iters, double a0, double b0) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i * 1e-12;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i * 1e-12;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x * 1e-9 + i;
    double a = a0, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s sms %d clock_khz %d\n", p.name, p.multiProcessorCount, clk_khz);
    double *out; CK(cudaMalloc(&out, 8));
    const int sms = p.multiProcessorCount;
    const int iters = 4096;
    printf("kind,warps_per_cta,ctas_per_sm,nacc,ms,tflops\n");
    int wlist[] = {4, 8, 16};
    for (int w : wlist) {
        for (int cps = 1; cps <= 2; ++cps) {
            int grid = sms * cps, block = 32 * w;
            double warps = (double)grid * w;
#define RUN884(N) { double ms = time_ms([&] { k_dmma884<N><<<grid, block>>>(out, iters, 1.0, 0.5); }, 5); \
                    printf("dmma884,%d,%d,%d,%.3f,%.2f\n", w, cps, N, ms, warps * iters * N * 512.0 / (ms * 1e-3) / 1e12); }
            RUN884(1) RUN884(2) RUN884(4) RUN884(8) RUN884(16) RUN884(32)
#define RUN16816(N) { double ms = time_ms([&] { k_dmma16816<N><<<grid, block>>>(out, iters / 4, 1.0, 0.5); }, 5); \
                    printf("dmma16816,%d,%d,%d,%.3f,%.2f\n", w, cps, N, ms, warps * (iters / 4) * N * 4096.0 / (ms * 1e-3) / 1e12); }
            RUN16816(1) RUN16816(2) RUN16816(4) RUN16816(8)
#define RUNFMA(N) { double ms = time_ms([&] { k_dfma<N><<<grid, block>>>(out, iters * 4, 1.0000001, 0.5); }, 5); \
                    printf("dfma,%d,%d,%d,%.3f,%.2f\n", w, cps, N, ms, warps * 32 * (iters * 4.0) * N * 2.0 / (ms * 1e-3) / 1e12); }
            RUNFMA(4) RUNFMA(8) RUNFMA(16)
        }
    }
    // sustained: run the best config for ~3 s to see clocks under power cap
    {
        int grid = sms * 2, block = 256; double warps = (double)grid * 8;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 0;
        for (int r = 0; r < 200; ++r) { k_dmma884<16><<<grid, block>>>(out, iters * 8, 1.0, 0.5); ++launches; }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("sustained_dmma884,8,2,16,%.3f,%.2f\n", ms, warps * (iters * 8.0) * 16 * 512.0 * launches / (ms * 1e-3) / 1e12);
    }
    return 0;
}
