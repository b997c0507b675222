// Dependent-chain latencies (cycles) of the warp-level primitives used by the LU column step, one warp, B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__global__ void k(long long *out, double seed) {
    __shared__ double sm[64];
    unsigned x = threadIdx.x * 2654435761u + 12345u;
    double d = seed + threadIdx.x;
    sm[threadIdx.x] = d;
    __syncthreads();
    long long t0, t1;
    // REDUX max (u32)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __reduce_max_sync(0xffffffffu, x ^ i) + threadIdx.x;
    t1 = clock64(); if (threadIdx.x == 0) out[0] = (t1 - t0) / N;
    // SHFL xor u32
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + i;
    t1 = clock64(); if (threadIdx.x == 0) out[1] = (t1 - t0) / N;
    // SHFL xor double + max (one butterfly stage)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double o = __shfl_xor_sync(0xffffffffu, d, 1 + (i & 15)); d = o > d ? o + 1e-9 : d + 1e-9; }
    t1 = clock64(); if (threadIdx.x == 0) out[2] = (t1 - t0) / N;
    // ballot
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __ballot_sync(0xffffffffu, (x >> (i & 7)) & 1) + threadIdx.x;
    t1 = clock64(); if (threadIdx.x == 0) out[3] = (t1 - t0) / N;
    // LDS dependent
    int idx = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = ((int)sm[idx & 63] + idx + 1) & 31;
    t1 = clock64(); if (threadIdx.x == 0) out[4] = (t1 - t0) / N;
    // DFMA dependent
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) d = fma(d, 1.0000001, 0.5);
    t1 = clock64(); if (threadIdx.x == 0) out[5] = (t1 - t0) / N;
    // double division dependent
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) d = 1.0 / (d + 1.5);
    t1 = clock64(); if (threadIdx.x == 0) out[6] = (t1 - t0) / N;
    // DSETP+select dependent (fabs compare)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) d = (fabs(d) > 0.3 + i) ? d * 0.5 : d + 1.0;
    t1 = clock64(); if (threadIdx.x == 0) out[7] = (t1 - t0) / N;
    // __syncthreads (block of blockDim threads)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) out[8] = (t1 - t0) / N;
    if (x == 0xdeadbeef || d == 123.0 || idx == 77) out[9] = x + idx;
}
int main() {
    long long *o; cudaMalloc(&o, 80); long long h[10];
    const char *names[] = {"REDUX.max u32", "SHFL u32", "SHFL f64 + max stage", "ballot", "LDS dependent", "DFMA dependent", "f64 division", "fabs cmp+select", "__syncthreads"};
    for (int threads : {32, 256, 1024}) {
        k<<<1, threads>>>(o, 1.0); cudaDeviceSynchronize();
        k<<<1, threads>>>(o, 1.0); cudaDeviceSynchronize();
        cudaMemcpy(h, o, 80, cudaMemcpyDeviceToHost);
        printf("threads=%d:", threads);
        for (int i = 0; i < 9; ++i) printf("  %s=%lld", names[i], h[i]);
        printf("\n");
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
