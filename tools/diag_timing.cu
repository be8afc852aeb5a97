// Phase timing of the diagonal-block kernels (clock64 stamps compiled in with -DGPX_DIAG_TIMING) + a DMMA latency
// probe.  Not part of the product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -lcuda
//                                          -o build/diag_timing tools/diag_timing.cu
#define GPX_DIAG_TIMING 1
#include "../gperks_b200/csrc/gpx_chol.cu"
#include <cstdio>
#include <vector>
#include <cmath>
int main() {
    using namespace gpx;
    const int Np = 128;
    std::vector<double> A(Np * Np);
    for (int i = 0; i < Np; ++i) for (int j = 0; j < Np; ++j) A[i * Np + j] = std::exp(-0.5 * (i - j) * (i - j) / 400.0) + (i == j ? 0.01 : 0.0);
    double *K, *S, *DT, *ld; int* info;
    cudaMalloc(&K, Np * Np * 8); cudaMalloc(&S, Np * Np * 8); cudaMalloc(&DT, Np * Np * 8); cudaMalloc(&ld, 64 * 8); cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4); cudaMemset(ld, 0, 64 * 8);
    cudaFuncSetAttribute(potrf_diag3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D3_SMEM);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(K, A.data(), Np * Np * 8, cudaMemcpyHostToDevice); cudaMemset(ld, 0, 64 * 8);
        potrf_diag3_kernel<<<1, D3_THREADS, D3_SMEM>>>(K, Np, 0, Np, S, DT, ld, info);
        cudaError_t e = cudaDeviceSynchronize();
        double h[32]; cudaMemcpy(h, ld, 32 * 8, cudaMemcpyDeviceToHost);
        printf("v3 rep %d err=%d cycles (warp 0 view): load=%.0f stage1+wait=%.0f stage2=%.0f stage3=%.0f writeback=%.0f | owner: factor8x8=%.0f Tinv=%.0f (sums over 16 panels)\n",
               rep, (int)e, h[8], h[9], h[10], h[11], h[12], h[20], h[21]);
    }
    cudaFuncSetAttribute(potrf_diag5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D5_SMEM);
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemcpy(K, A.data(), Np * Np * 8, cudaMemcpyHostToDevice); cudaMemset(ld, 0, 64 * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        potrf_diag5_kernel<true><<<1, D5_THREADS, D5_SMEM>>>(K, Np, 0, Np, S, DT, ld, info);
        { cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) printf("launch error: %s\n", cudaGetErrorString(le)); }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double h[64]; cudaMemcpy(h, ld, 64 * 8, cudaMemcpyDeviceToHost);
        printf("v5 rep %d err=%d %.1f us logdet=%.6f | factor warp: wait_block=%.0f load+ldl=%.0f T+store=%.0f wait#1=%.0f out+wait#2=%.0f\n", rep, (int)e, ms * 1e3, h[0],
               h[40], h[41], h[43], h[44], h[45]);
        printf("            compute warp 0: wait#1=%.0f stage2=%.0f wait#2=%.0f nextdiag=%.0f stage3=%.0f output=%.0f\n", h[48], h[49], h[50], h[51], h[52], h[53]);
    }
    cudaFuncSetAttribute(potrf_diag5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D5_SMEM);
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemcpy(K, A.data(), Np * Np * 8, cudaMemcpyHostToDevice); cudaMemset(ld, 0, 64 * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        potrf_diag5_kernel<false><<<1, D5_THREADS, D5_SMEM>>>(K, Np, 0, Np, S, DT, ld, info);
        { cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) printf("launch error: %s\n", cudaGetErrorString(le)); }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double h[64]; cudaMemcpy(h, ld, 64 * 8, cudaMemcpyDeviceToHost);
        printf("v5 Cholesky-only rep %d err=%d %.1f us logdet=%.6f | factor warp: wait_block=%.0f load+ldl=%.0f T+store=%.0f wait#1=%.0f out+wait#2=%.0f\n", rep, (int)e, ms * 1e3, h[0],
               h[40], h[41], h[43], h[44], h[45]);
        printf("            compute warp 0: wait#1=%.0f stage2=%.0f wait#2=%.0f nextdiag=%.0f stage3=%.0f output=%.0f\n", h[48], h[49], h[50], h[51], h[52], h[53]);
    }
    return 0;
}

// ---- DMMA latency probe: one warp, dependent chain vs 4 independent chains
__global__ void dmma_latency_kernel(double* out) {
    double c0 = 0, c1 = 0, d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
    double a = 1.0 + threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-3;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) gpx::dmma884(c0, c1, a, b);
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) gpx::dmma884(d0[u], d1[u], a, b);
    }
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / 256.0; out[1] = (double)(t2 - t1) / 256.0; }
    out[2 + threadIdx.x] = c0 + c1 + d0[0] + d0[1] + d0[2] + d0[3] + d1[0] + d1[1] + d1[2] + d1[3];
}
struct LatencyProbe {
    LatencyProbe() {
        double* o; cudaMalloc(&o, 64 * 8);
        dmma_latency_kernel<<<1, 32>>>(o);
        cudaDeviceSynchronize();
        double h[2]; cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
        printf("DMMA.8x8x4 single warp: dependent chain %.1f cycles/op, 4 independent chains %.1f cycles/op\n", h[0], h[1]);
    }
} g_probe;
