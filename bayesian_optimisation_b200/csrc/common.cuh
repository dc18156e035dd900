// Shared device/host helpers for libbogp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/bogp.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbogp is written for sm_100a (B200) only"
#endif

namespace bogp {

constexpr int kDiagNB   = 64;    // Cholesky diagonal block / panel width
constexpr int kPad      = 256;   // fitted systems are padded to a multiple of this (= acquisition row block)
constexpr int kAcqBM    = 256;   // rows of W per acquisition CTA tile
constexpr int kAcqBN    = 64;    // candidates per acquisition CTA tile
constexpr int kAcqKB    = 16;    // k extent of one pipeline stage
constexpr int kAcqStages = 5;

void set_error(const char* fmt, ...);

// Function attributes (opt-in shared memory) are per device: `flags` is a per-kernel static array.
struct DeviceOnce {
    bool done[64] = {};
    bool need(int device) { if (device < 0 || device >= 64) return true; if (done[device]) return false; done[device] = true; return true; }
};

#define BOGP_CUDA_CHECK(expr)                                                             \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            bogp::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,          \
                            cudaGetErrorString(_e));                                      \
            return BOGP_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

// kernel ids for bogp_profile_read
#define BOGP_PROF_PANEL    0
#define BOGP_PROF_TRIGEMM  1
#define BOGP_PROF_FINALIZE 2
#define BOGP_PROF_MERGE    3

// Time one launch with CUDA events on the launching stream when profiling is on (serialises the stream).
#define BOGP_PROFILED(ctx, id, launch)                                                    \
    do {                                                                                  \
        if ((ctx)->profile) cudaEventRecord((ctx)->ev[0], (ctx)->stream);                 \
        launch;                                                                           \
        if ((ctx)->profile) {                                                             \
            cudaEventRecord((ctx)->ev[1], (ctx)->stream);                                 \
            cudaEventSynchronize((ctx)->ev[1]);                                           \
            float _ms = 0.f; cudaEventElapsedTime(&_ms, (ctx)->ev[0], (ctx)->ev[1]);      \
            (ctx)->prof_ms[id] += _ms; (ctx)->prof_n[id]++;                               \
        }                                                                                 \
    } while (0)

#define BOGP_LAUNCH_CHECK(ctx)                                                            \
    do {                                                                                  \
        (ctx)->launches++;                                                                \
        BOGP_CUDA_CHECK(cudaGetLastError());                                              \
    } while (0)

}  // namespace bogp

struct bogp_ctx {
    int          device;
    int          sm_count;
    cudaStream_t stream;
    int64_t      launches;
    // small device scratch
    double*      d_scalars;     // 64 doubles
    int*         d_flags;       // 64 ints
    double*      d_block_score; // kMaxBlocks
    long long*   d_block_index; // kMaxBlocks
    double*      h_pinned;      // 64 doubles pinned host staging
    // optional per-kernel timing of the acquisition sweep (bogp_profile): CUDA events on the launching stream
    // second stream + events: the k_* panel of chunk s+1 is built while chunk s is on the tensor cores
    cudaStream_t aux_stream;      // high priority: serial chains / panel builder
    cudaStream_t aux2_stream;     // default priority: work that may fill idle SMs (interleaved triangular inverse)
    cudaEvent_t  ev_fork, ev_panel[2], ev_done[2], ev_aux2;
    int64_t      inblock_launches;   // launches of the fused in-block kernel (its grid-barrier counter only grows)
    int          acquire_path;  // 0 = FP64 DMMA, 1 = INT8 digit slices on tcgen05 (bogp_set_acquire_path)
    int          profile;
    cudaEvent_t  ev[2];
    double       prof_ms[8];
    int64_t      prof_n[8];
};

namespace bogp {
constexpr int kMaxReduceBlocks = 4096;

// ---------------------------------------------------------------- device primitives
#ifdef __CUDACC__

// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor path (SASS: DMMA.8x8x4).
// lane l holds A[l>>2][l&3], B[l&3][l>>2], C[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte cp.async with zero-fill when !valid
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

// ---- mbarrier + bulk (TMA) copies
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (score, index) ordering of the reference: larger score wins, ties -> smaller flat index.
__device__ __forceinline__ bool better(double s, long long i, double bs, long long bi) {
    return (s > bs) || (s == bs && i < bi);
}

#endif  // __CUDACC__
}  // namespace bogp
