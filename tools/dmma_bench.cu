// Micro-benchmark: FP64 tensor (DMMA.8x8x4) vs FP64 FMA issue rates on B200, to set the roofline
// denominator for the acquisition GEMM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench.bin dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters) {
    double c[NACC][2];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters) {
    double c[NACC];
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0; for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA + independent DFMA stream in the same warp
template <int NACC>
__global__ void k_mixed(double* out, int iters) {
    double c[NACC][2], f[NACC];
    for (int i = 0; i < NACC; i++) { c[i][0] = c[i][1] = 0.0; f[i] = i; }
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3, fa = 1.0 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) { dmma(c[i][0], c[i][1], a, b); f[i] = fma(f[i], fa, 1e-9); }
    }
    double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, int iters) {
    double x = -1e-3 * threadIdx.x, s = 0;
    for (int it = 0; it < iters; it++) { s += exp(x); x -= 1e-6; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d clock %d kHz\n", p.name, sms, clk);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        for (int cta : {1, 2}) {
            int threads = warps * 32;
            float ms = timeit([&] { k_dmma<16><<<sms * cta, threads>>>(out, iters); });
            double flops = 2.0 * 256 * 16 * (double)iters * warps * sms * cta;
            printf("DMMA   warps/CTA %2d CTAs/SM %d: %8.3f ms  %7.2f TFLOP/s\n", warps, cta, ms, flops / ms * 1e-9);
        }
    }
    for (int warps : {8, 16, 32}) {
        int threads = warps * 32;
        float ms = timeit([&] { k_dfma<16><<<sms * 2, threads>>>(out, iters); });
        double flops = 2.0 * 32 * 16 * (double)iters * warps * sms * 2;
        printf("DFMA   warps/CTA %2d CTAs/SM 2: %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flops / ms * 1e-9);
    }
    for (int warps : {8, 16}) {
        int threads = warps * 32;
        float ms = timeit([&] { k_mixed<16><<<sms * 2, threads>>>(out, iters); });
        double fl_mma = 2.0 * 256 * 16 * (double)iters * warps * sms * 2, fl_fma = 2.0 * 32 * 16 * (double)iters * warps * sms * 2;
        printf("MIXED  warps/CTA %2d CTAs/SM 2: %8.3f ms  DMMA %7.2f + DFMA %7.2f TFLOP/s\n", warps, ms, fl_mma / ms * 1e-9, fl_fma / ms * 1e-9);
    }
    {
        float ms = timeit([&] { k_exp<<<sms * 4, 256>>>(out, 4000); });
        double n = 4000.0 * 256 * sms * 4;
        printf("EXP    fp64: %8.3f ms  %7.2f Gexp/s\n", ms, n / ms * 1e-6);
    }
    return 0;
}
