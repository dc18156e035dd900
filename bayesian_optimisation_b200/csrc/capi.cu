// Context management and error reporting of libbogp.
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace bogp {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace bogp

using namespace bogp;

extern "C" const char* bogp_version(void) { return "bogp 2 sm_100a"; }
extern "C" const char* bogp_last_error(void) { return g_error; }

extern "C" int bogp_device_count(int* out) {
    if (!out) { set_error("bogp_device_count: null output pointer"); return BOGP_ERR_BAD_ARG; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) { *out = 0; set_error("bogp_device_count: %s", cudaGetErrorString(e)); return BOGP_ERR_CUDA; }
    *out = count;
    return BOGP_OK;
}

extern "C" int bogp_create(int device, bogp_ctx** out) {
    if (!out) { set_error("bogp_create: null output pointer"); return BOGP_ERR_BAD_ARG; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("bogp_create: no CUDA device available (%s); libbogp has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return BOGP_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("bogp_create: device %d out of range [0,%d)", device, count); return BOGP_ERR_BAD_ARG; }
    BOGP_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    BOGP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("bogp_create: device %d is sm_%d%d; libbogp is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return BOGP_ERR_CUDA;
    }
    bogp_ctx* c = new bogp_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device; c->sm_count = prop.multiProcessorCount; c->stream = nullptr; c->launches = 0;
    {
        const char* e = getenv("BOGP_ACQUIRE_PATH");
        c->acquire_path = (e && (!strcmp(e, "fp64") || !strcmp(e, "dmma") || !strcmp(e, "0"))) ? BOGP_PATH_FP64_DMMA
                        : (e && (!strcmp(e, "i8") || !strcmp(e, "int8") || !strcmp(e, "1"))) ? BOGP_PATH_INT8_TCGEN05 : BOGP_PATH_DEFAULT;
    }
    c->screening = 1;
    c->fused = 0;      // measured: the two-stream pipeline of separate kernels is ~6 % faster at N = 4096
    BOGP_CUDA_CHECK(cudaMalloc(&c->d_scalars, 64 * sizeof(double)));
    BOGP_CUDA_CHECK(cudaMalloc(&c->d_flags, 64 * sizeof(int)));
    BOGP_CUDA_CHECK(cudaMalloc(&c->d_block_score, kMaxReduceBlocks * sizeof(double)));
    BOGP_CUDA_CHECK(cudaMalloc(&c->d_block_index, kMaxReduceBlocks * sizeof(long long)));
    BOGP_CUDA_CHECK(cudaMemset(c->d_scalars, 0, 64 * sizeof(double)));
    BOGP_CUDA_CHECK(cudaMemset(c->d_flags, 0, 64 * sizeof(int)));
    {   // the panel stream gets the highest priority: its CTAs fit next to a running tensor-core CTA, and the
        // work distributor only interleaves a second kernel ahead of pending CTAs if it has priority
        int lo = 0, hi = 0;
        BOGP_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        BOGP_CUDA_CHECK(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi));
    }
    BOGP_CUDA_CHECK(cudaStreamCreateWithFlags(&c->aux2_stream, cudaStreamNonBlocking));
    BOGP_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_aux2, cudaEventDisableTiming));
    BOGP_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
        BOGP_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_panel[i], cudaEventDisableTiming));
        BOGP_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    BOGP_CUDA_CHECK(cudaEventCreate(&c->ev[0]));
    BOGP_CUDA_CHECK(cudaEventCreate(&c->ev[1]));
    *out = c;
    return BOGP_OK;
}

extern "C" int bogp_set_acquire_path(bogp_ctx* ctx, int path) {
    if (!ctx || (path != BOGP_PATH_FP64_DMMA && path != BOGP_PATH_INT8_TCGEN05)) { set_error("bogp_set_acquire_path: bad argument"); return BOGP_ERR_BAD_ARG; }
    ctx->acquire_path = path;
    return BOGP_OK;
}
extern "C" int bogp_get_acquire_path(const bogp_ctx* ctx) { return ctx ? ctx->acquire_path : -1; }

extern "C" int bogp_set_screening(bogp_ctx* ctx, int enable) {
    if (!ctx) { set_error("bogp_set_screening: null context"); return BOGP_ERR_BAD_ARG; }
    ctx->screening = enable ? 1 : 0;
    return BOGP_OK;
}
extern "C" int bogp_get_screening(const bogp_ctx* ctx) { return ctx ? ctx->screening : -1; }

extern "C" int bogp_set_global_seed(bogp_ctx* ctx, int enable) {
    if (!ctx) { set_error("bogp_set_global_seed: null context"); return BOGP_ERR_BAD_ARG; }
    ctx->global_seed = enable ? 1 : 0;
    return BOGP_OK;
}

extern "C" int bogp_set_fused(bogp_ctx* ctx, int enable, int group) {
    if (!ctx || group < 0 || group > 64) { set_error("bogp_set_fused: bad argument"); return BOGP_ERR_BAD_ARG; }
    ctx->fused = enable ? 1 : 0;
    ctx->fused_group = group;
    return BOGP_OK;
}
extern "C" int bogp_get_fused(const bogp_ctx* ctx) { return ctx ? ctx->fused : -1; }

extern "C" int bogp_profile(bogp_ctx* ctx, int enable) {
    if (!ctx) { set_error("bogp_profile: null context"); return BOGP_ERR_BAD_ARG; }
    ctx->profile = enable ? 1 : 0;
    for (int i = 0; i < 8; i++) { ctx->prof_ms[i] = 0.0; ctx->prof_n[i] = 0; }
    return BOGP_OK;
}

extern "C" int bogp_profile_read(const bogp_ctx* ctx, int kernel_id, double* ms_total, int64_t* launches) {
    if (!ctx || kernel_id < 0 || kernel_id >= 8) { set_error("bogp_profile_read: bad argument"); return BOGP_ERR_BAD_ARG; }
    if (ms_total) *ms_total = ctx->prof_ms[kernel_id];
    if (launches) *launches = ctx->prof_n[kernel_id];
    return BOGP_OK;
}

extern "C" void bogp_destroy(bogp_ctx* ctx) {
    if (!ctx) return;
    cudaEventDestroy(ctx->ev[0]); cudaEventDestroy(ctx->ev[1]);
    cudaEventDestroy(ctx->ev_fork);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_panel[i]); cudaEventDestroy(ctx->ev_done[i]); }
    cudaEventDestroy(ctx->ev_aux2);
    cudaStreamDestroy(ctx->aux_stream);
    cudaStreamDestroy(ctx->aux2_stream);
    cudaFree(ctx->d_scalars); cudaFree(ctx->d_flags); cudaFree(ctx->d_block_score); cudaFree(ctx->d_block_index);
    delete ctx;
}

extern "C" int bogp_set_stream(bogp_ctx* ctx, void* cuda_stream) {
    if (!ctx) { set_error("bogp_set_stream: null context"); return BOGP_ERR_BAD_ARG; }
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    return BOGP_OK;
}
extern "C" int bogp_sm_count(const bogp_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t bogp_launch_count(const bogp_ctx* ctx) { return ctx ? ctx->launches : 0; }
