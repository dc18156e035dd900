// Measurement aids: issue-rate peaks of the two pipes the hot path runs on, measured on the box the
// benchmark runs on (MEASURED_PEAKS.json has neither an int8 nor an fp64 entry).  bench.py divides the
// achieved rates of trigemm_i8_kernel / the fit by these, so every roofline fraction it prints is
// "of measured" and reproducible.  Nothing on the product path calls this file.
//
//   BOGP_PEAK_I8_UMMA : one CTA per SM issues back-to-back tcgen05.mma kind::i8 (M=128, N=256, K=32,
//                       operands resident in shared memory, two alternating TMEM accumulators) --
//                       no loads, no epilogue: the rate the tensor pipe accepts int8 MMAs at.
//   BOGP_PEAK_F64_DMMA: 16 independent DMMA.8x8x4 accumulators per warp, 8 warps x 2 CTAs per SM.
#include "common.cuh"

namespace bogp {

__device__ __forceinline__ uint64_t peak_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

constexpr int kPeakM = 128, kPeakN = 256, kPeakK = 32;

__global__ void __launch_bounds__(128, 1) peak_i8_kernel(int iters, int* sink) {
    __shared__ __align__(128) unsigned char sa[kPeakM * kPeakK];
    __shared__ __align__(128) unsigned char sb[kPeakN * kPeakK];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kPeakM * kPeakK; i += 128) sa[i] = (unsigned char)(i * 7 + 1);
    for (int i = tid; i < kPeakN * kPeakK; i += 128) sb[i] = (unsigned char)(i * 13 + 5);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        // D = s32, A = s8, B = s8, K-major, M = 128, N = 256
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPeakN >> 3) << 17) | ((uint32_t)(kPeakM >> 4) << 24);
        const uint64_t da = peak_desc_kmajor(smem_u32(sa), kPeakM * 16, 128);
        const uint64_t db = peak_desc_kmajor(smem_u32(sb), kPeakN * 16, 128);
        for (int it = 0; it < iters; it++) {
            const uint32_t acc = it >= 2 ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem + (uint32_t)((it & 1) * kPeakN)), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (sink && iters < 0) {           // never true: keeps the accumulators observable
        uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink[blockIdx.x * 128 + tid] = (int)r;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

__global__ void __launch_bounds__(256, 2) peak_dmma_kernel(int iters, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i][0] = c[i][1] = 0.0;
    const double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    if (s == 12345.678) sink[blockIdx.x * 256 + threadIdx.x] = s;
}

}  // namespace bogp

using namespace bogp;

// One timed launch of the chosen loop; returns its duration and the operations it executed.
static int peak_launch(bogp_ctx* ctx, int kind, float* ms, double* work) {
    cudaStream_t st = ctx->stream;
    BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], st));
    if (kind == BOGP_PEAK_I8_UMMA) {
        const int iters = 40000;
        peak_i8_kernel<<<ctx->sm_count, 128, 0, st>>>(iters, nullptr);
        *work = 2.0 * kPeakM * kPeakN * kPeakK * (double)iters * ctx->sm_count;
    } else {
        const int iters = 20000;
        peak_dmma_kernel<<<2 * ctx->sm_count, 256, 0, st>>>(iters, nullptr);
        *work = 2.0 * 256 * 16 * (double)iters * 8 * 2 * ctx->sm_count;
    }
    BOGP_LAUNCH_CHECK(ctx);
    BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], st));
    BOGP_CUDA_CHECK(cudaEventSynchronize(ctx->ev[1]));
    BOGP_CUDA_CHECK(cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
    return BOGP_OK;
}

extern "C" int bogp_measure_peak(bogp_ctx* ctx, int kind, double sustain_seconds, double* h_burst_tera, double* h_sustained_tera) {
    if (!ctx || !h_burst_tera || (kind != BOGP_PEAK_I8_UMMA && kind != BOGP_PEAK_F64_DMMA)) {
        set_error("bogp_measure_peak: bad argument"); return BOGP_ERR_BAD_ARG;
    }
    float ms = 0.f; double work = 0.0, best_ms = 1e30;
    for (int rep = 0; rep < 6; rep++) {          // rep 0 is the warm-up
        const int rc = peak_launch(ctx, kind, &ms, &work);
        if (rc) return rc;
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    *h_burst_tera = work / (best_ms * 1e-3) * 1e-12;
    if (h_sustained_tera) {                       // back-to-back launches for sustain_seconds: the rate under the power cap
        double total_ms = 0.0, total_work = 0.0;
        while (total_ms < sustain_seconds * 1e3) {
            const int rc = peak_launch(ctx, kind, &ms, &work);
            if (rc) return rc;
            total_ms += ms; total_work += work;
        }
        *h_sustained_tera = total_ms > 0.0 ? total_work / (total_ms * 1e-3) * 1e-12 : *h_burst_tera;
    }
    return BOGP_OK;
}
