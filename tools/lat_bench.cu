// Dependent-issue latencies on B200 (single warp): DFMA, rcp.approx.f64, SHFL (64-bit), LDS, bar.sync.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc) {
    __shared__ double sh[64];
    double x = 1.0 + threadIdx.x * 1e-9, y = 1.000000001;
    sh[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = fma(x, y, 1e-12);
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    long long t3 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) { x = sh[((int)x) & 63]; }
    long long t4 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) __syncthreads();
    long long t5 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) x = x * y;
    long long t6 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; }
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 8 * 256); cudaMalloc(&c, 8 * 8);
    for (int threads : {32, 256}) {
        k<<<1, threads>>>(o, c); cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
        printf("threads %3d: DFMA %.1f  RCP64H %.1f  SHFL64 %.1f  LDS(dep) %.1f  BAR %.1f  DMUL %.1f  cycles per dependent op\n", threads,
               h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 256.0);
    }
    return 0;
}
