// Dependent-chain latencies (cycles/op, one warp) of the operations on the 8x8 factor's critical path. Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, double seed) {
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0, t1;
    double res[8];
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) x = fma(x, y, 1e-9);
    }
    t1 = clock64(); res[0] = (double)(t1 - t0) / 512.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) x = x * y;
    }
    t1 = clock64(); res[1] = (double)(t1 - t0) / 512.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) x = __drcp_rn(x + 1.5);
    }
    t1 = clock64(); res[2] = (double)(t1 - t0) / 256.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) x = rsqrt(x + 1.5);
    }
    t1 = clock64(); res[3] = (double)(t1 - t0) / 256.0;
    float f = (float)x, g = 1.0000001f;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) f = fmaf(f, g, 1e-9f);
    }
    t1 = clock64(); res[4] = (double)(t1 - t0) / 512.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    }
    t1 = clock64(); res[5] = (double)(t1 - t0) / 512.0;
    __shared__ double sm[64];
    sm[threadIdx.x] = x; sm[threadIdx.x + 32] = y;
    __syncwarp();
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { sm[idx] = x; __syncwarp(); x = sm[(idx + 1) & 31]; __syncwarp(); }
    }
    t1 = clock64(); res[6] = (double)(t1 - t0) / 512.0;
    // two independent DFMA chains (does a second chain come for free?)
    double z = y;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { x = fma(x, y, 1e-9); z = fma(z, y, 1e-9); }
    }
    t1 = clock64(); res[7] = (double)(t1 - t0) / 512.0;
    if (threadIdx.x == 0) for (int i = 0; i < 8; ++i) out[i] = res[i];
    out[8 + threadIdx.x] = x + f + z;
}
int main() {
    double* o; cudaMalloc(&o, 64 * 8);
    for (int rep = 0; rep < 2; ++rep) {
        probe<<<1, 32>>>(o, 1.0);
        cudaDeviceSynchronize();
    }
    double h[8]; cudaMemcpy(h, o, 64, cudaMemcpyDeviceToHost);
    printf("dependent-chain cycles/op, one warp: DFMA %.1f  DMUL %.1f  drcp_rn %.1f  rsqrt(double) %.1f  FFMA %.1f  SHFL64 %.1f  STS+LDS roundtrip %.1f  2xDFMA chains %.1f per pair\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    return 0;
}
