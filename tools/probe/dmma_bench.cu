// Throughput of mma.sync.m8n8k4.f64 (DMMA) against plain DFMA on one GPU: per-SM rates with 8 independent chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dmma(double* out, int iters) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters) {
    double c[16];
    for (int k = 0; k < 16; ++k) c[k] = threadIdx.x + k;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) c[k] = fma(c[k], a, b);
    }
    double s = 0;
    for (int k = 0; k < 16; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, 148 * 1024 * 8 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 20000;
        float ms;
        k_dmma<<<p.multiProcessorCount, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma<<<p.multiProcessorCount, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double dmma_per_sm = (double)warps * iters * 8;
        printf("warps/SM %2d  DMMA: %.3f ms  %.2f clk/DMMA/SM (at %d MHz)  %.1f TFLOP/s\n", warps, ms,
               ms * 1e-3 * p.clockRate * 1e3 / dmma_per_sm, p.clockRate / 1000,
               dmma_per_sm * p.multiProcessorCount * 512 / (ms * 1e-3) / 1e12);
        k_dfma<<<p.multiProcessorCount, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<<<p.multiProcessorCount, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double dfma_per_sm = (double)warps * iters * 16;
        printf("warps/SM %2d  DFMA: %.3f ms  %.2f clk/warp-DFMA/SM  %.1f TFLOP/s\n", warps, ms,
               ms * 1e-3 * p.clockRate * 1e3 / dfma_per_sm, dfma_per_sm * p.multiProcessorCount * 64 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
