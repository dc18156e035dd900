// Shared device/host helpers for libbogp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include <nvtx3/nvToolsExt.h>     // header-only; ranges cost nothing unless a tool (nsys, ncu --nvtx) is attached

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

// NVTX range around the host side of a library call (what enqueues a fit, a sweep, a batched LML): visible in nsys / ncu
// timelines next to the kernels it launched.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

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
    int          acquire_path;  // 0 = FP64 DMMA, 1 = INT8 digit slices on tcgen05 (bogp_set_acquire_path)
    int          profile;
    int          screening;     // arg-max-only sweeps: screen by the posterior-mean bound, score survivors exactly (bogp_set_screening)
    int          global_seed;   // screened sweeps: the seed sample spans the WHOLE candidate set, not only [c_begin, c_end) (bogp_set_global_seed)
    int          fused;         // INT8 path: one persistent fused kernel per sweep (default) instead of per-chunk panel / product / finalize / merge kernels
    int          fused_group;   // candidate tiles per work group of the fused kernel (0 = automatic: ~32 MB of panel digits)
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


// ---- exp(t) for t <= 0: the kernel function exp(-0.5 * squared distance) of every Gram / k_* kernel.
// Table-driven: t = (64 n + j) ln2/64 + r, |r| <= ln2/128;  exp(t) = 2^n * 2^(j/64) * (1 + p(r)), p = degree-6
// Taylor polynomial of expm1 (truncation 3e-20).  12 fp64 operations, coefficients as constant-bank operands, no
// branches (the library exp costs about 18 plus two dozen constant moves and selects).  Maximum error 1.0 ulp
// measured against long-double expl over 2e7 arguments in [-708, 0]; results below 2^-1022 (t < -708) are flushed
// to zero.  NaN in -> NaN out (every step propagates the canonical NaN, whose low word selects n = 0).
// Precondition: t <= 0 or NaN.  `tab` is kExp2Tab or a shared-memory copy of it (same values, same result).
static __device__ const double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
static __constant__ double kExpC[8] = {
    0x1.71547652b82fep+6,        // 64 / ln2
    -0x1.62e42fe000000p-7,       // -(ln2 / 64), high 29 bits: kf * hi is exact
    -0x1.f473de6af278fp-36,      // -(ln2 / 64), rest
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.0};

__device__ __forceinline__ double exp_nonpos(double t, const double* __restrict__ tab) {
    const double kd = fma(t, kExpC[0], 6755399441055744.0);                  // + 1.5 * 2^52: round to nearest integer
    const int ki = __double2loint(kd);
    const double kf = kd - 6755399441055744.0;
    double r = fma(kf, kExpC[1], t);
    r = fma(kf, kExpC[2], r);
    const double T = tab[ki & 63];
    double q = fma(kExpC[3], r, kExpC[4]);
    q = fma(q, r, kExpC[5]);
    q = fma(q, r, kExpC[6]);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    const double m = fma(T, p, T);
    const double res = __hiloint2double(__double2hiint(m) + ((ki << 14) & 0xfff00000), __double2loint(m));
    return t < -708.0 ? 0.0 : res;
}
__device__ __forceinline__ double exp_nonpos(double t) { return exp_nonpos(t, kExp2Tab); }

// i-th candidate of the seed sample of a screened sweep over [begin, begin + total): the golden-ratio sequence
// frac(i * phi) * total.  (A plain stride is a trap on grids: total / 4096 is a power of the radix for the 8^10 grid, and
// every seed then sits on the face where the trailing coordinates are all 0.)
__device__ __forceinline__ long long seed_index(unsigned long long i, long long begin, long long total) {
    return begin + (long long)__umul64hi((i + 1ull) * 0x9E3779B97F4A7C15ull, (unsigned long long)total);
}

// (score, index) ordering of the reference: larger score wins, ties -> smaller flat index.
__device__ __forceinline__ bool better(double s, long long i, double bs, long long bi) {
    return (s > bs) || (s == bs && i < bi);
}

// acquisition value of one candidate: explore*sigma - mu (lower_confidence_bound, point_selector.py:204) or
// expected improvement over f_best (minimisation)
__device__ __forceinline__ double acquisition_value(int kind, double mu, double sigma, double explore, double f_best) {
    if (kind == BOGP_ACQ_LCB) return __dsub_rn(__dmul_rn(explore, sigma), mu);   // two roundings like numpy, no FMA
    const double imp = f_best - mu;
    if (!(sigma > 0.0)) return imp > 0.0 ? imp : 0.0;
    const double z = imp / sigma;
    return imp * normcdf(z) + sigma * (exp(-0.5 * z * z) * 0.3989422804014326779);
}

#endif  // __CUDACC__
}  // namespace bogp
