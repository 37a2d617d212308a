// read-bandwidth ceiling probe: sum of a 2 GB float array with the same access pattern as k_extend
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int RB, int NC>
__global__ void __launch_bounds__(256) k_read(const float* __restrict__ X, long n, int d_pad, float* out) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const long n_units = (n + 31) >> 5;
    float acc = 0.f;
    for (long unit = warp; unit < n_units; unit += nwarps) {
        const long row0 = unit << 5;
#pragma unroll 1
        for (int r = 0; r < 32; r += RB) {
            float4 x[RB][NC];
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                const float* p = X + (row0 + r + rr) * (long)d_pad + lane * 4;
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(x[rr][c].x), "=f"(x[rr][c].y), "=f"(x[rr][c].z), "=f"(x[rr][c].w) : "l"(p + c * 128));
            }
#pragma unroll
            for (int rr = 0; rr < RB; ++rr)
#pragma unroll
                for (int c = 0; c < NC; ++c) acc += x[rr][c].x + x[rr][c].y + x[rr][c].z + x[rr][c].w;
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}
template <int RB>
void run(const float* X, long n, float* out, int blocks_per_sm, const char* tag) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = 148 * blocks_per_sm;
    for (int it = 0; it < 3; ++it) k_read<RB, 4><<<blocks, 256>>>(X, n, 512, out);
    cudaEventRecord(a);
    for (int it = 0; it < 10; ++it) k_read<RB, 4><<<blocks, 256>>>(X, n, 512, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%s RB=%d blocks/SM=%d: %.1f us  %.1f GB/s\n", tag, RB, blocks_per_sm, ms / 10 * 1e3, n * 2048.0 / (ms / 10 * 1e-3) / 1e9);
}
int main() {
    long n = 1000000; float *X, *out;
    cudaMalloc(&X, n * 2048); cudaMalloc(&out, 4); cudaMemset(X, 0, n * 2048);
    for (int bps : {1, 2, 3, 4, 6, 8}) { run<2>(X, n, out, bps, "read"); run<4>(X, n, out, bps, "read"); run<8>(X, n, out, bps, "read"); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
