// FP64 issue throughput of ONE CTA on one SM: W warps each issue `CH` independent DFMA chains of depth `DEPTH`
// (the shape of the rank-4 update in chol_diag_block).  Prints cycles per batch and DFMA lanes per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput_bench.bin fp64_tput_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, int DEPTH>
__global__ void k(double* out, long long* cyc, double x, double y) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 64; rep++) {
#pragma unroll
        for (int d = 0; d < DEPTH; d++)
#pragma unroll
            for (int i = 0; i < CH; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += a[i];
    __syncthreads();
    const long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
    for (int warps = 1; warps <= 16; warps *= 2) {
        long long h = 0;
        k<16, 4><<<1, warps * 32>>>(out, cyc, 0.999, 1e-3); cudaDeviceSynchronize();
        k<16, 4><<<1, warps * 32>>>(out, cyc, 0.999, 1e-3); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_batch = h / 64.0;     // one batch = 64 DFMA per thread (16 chains x depth 4)
        printf("warps %2d: %.0f cycles per 64-DFMA batch -> %.1f DFMA lanes/clk/SM (%s)\n", warps, per_batch, warps * 32 * 64.0 / per_batch, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
