// C (+)= alpha * A * B^T on the FP64 tensor path with TMA-fed operands (the Cholesky trailing update).
//
// The generic cp.async GEMM (gemm_f64.cuh) keeps the DMMA pipe only ~57 % busy: every thread spends
// issue slots on address arithmetic and the CTA meets at a barrier per k step.  Here the two row-major,
// K-contiguous operands are fetched by the TMA unit instead: a tensor map with a box of {4 k, 128 rows}
// lands in shared memory as [row][4 doubles], which is exactly the DMMA.8x8x4 fragment order (a warp's
// fragment = one contiguous 256-byte line, no padding, no bank conflicts).  One producer lane issues
// 16 boxes per 32-wide k stage; 4 consumer warps (warp tile 32 x 64, CTA tile 128 x 64) do nothing but LDS + DMMA;
// stages are handed over with mbarriers (no CTA-wide barrier in the main loop); two CTAs share an SM.
#include "common.cuh"
#include "gemm_f64.cuh"
#include <cuda.h>

namespace bogp {

// CTA tile 128 x 64 with FOUR consumer warps (32 x 64 each) and two stages, TWO CTAs per SM: while one CTA is in the
// read-modify-write epilogue of its tile the other one keeps the DMMA pipe busy (with one 128 x 128 CTA per SM the
// pipe idled through every epilogue: 65 % active).
constexpr int TBM = 128, TBN = 64, TKB = 32, TSTAGES = 2, TCONSUMERS = 4;
constexpr int kTABytes = TBM * TKB * 8;                       // 32 KB
constexpr int kTBBytes = TBN * TKB * 8;                       // 16 KB
constexpr int kTStageBytes = kTABytes + kTBBytes;             // 48 KB
constexpr size_t kTSmem = (size_t)TSTAGES * kTStageBytes + 2 * TSTAGES * 8 + TSTAGES * 4 + 64;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct TmaGemmArgs {
    double* C; int64_t ldc;
    int M, N, K;
    double alpha; int accumulate, lower_only;
    int nbm, nbn, ntiles;        // tile grid and the number of tiles that are actually computed
    int* counter;                // dynamic tile scheduler (zeroed before the launch)
};

// Persistent CTAs (one per SM) with a dynamic tile scheduler: the producer lane draws the next tile from a global
// counter, announces it to the consumer warps through a per-stage shared-memory slot that travels with the stage's
// `full` barrier, and keeps the TMA ring filled across tile boundaries -- so the first stages of the next tile are in
// flight while the consumers are still in the read-modify-write epilogue of the current one.  Dynamic scheduling
// matters because this kernel shares the GPU with the serial chain and the interleaved inverse: a CTA whose SM is
// busy elsewhere simply takes fewer tiles.
__global__ void __launch_bounds__((TCONSUMERS + 1) * 32, 2)
gemm_tma_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TmaGemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TSTAGES * kTStageBytes);
    uint64_t* empt = full + TSTAGES;
    volatile int* tile_slot = reinterpret_cast<volatile int*>(empt + TSTAGES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = (g.K + TKB - 1) / TKB;

    if (tid == 0) {
        for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empt[s], TCONSUMERS); }
        fence_mbar_init();
    }
    __syncthreads();

    // valid tile index -> (row block, column block); lower_only keeps column blocks <= row block
    auto tile_of = [&](int t, int& bm, int& bn) {
        if (!g.lower_only) { bm = t / g.nbn; bn = t - bm * g.nbn; return; }
        int r = 0;
        for (; r < g.nbm; r++) { const int c = ((r + 1) * (TBM / TBN) < g.nbn) ? (r + 1) * (TBM / TBN) : g.nbn; if (t < c) break; t -= c; }
        bm = r; bn = t;
    };

    if (warp == TCONSUMERS) {
        if (lane == 0) {
            int it = 0;
            for (;;) {
                const int t = atomicAdd(g.counter, 1);
                const int s0 = it % TSTAGES;
                if (it >= TSTAGES) mbar_wait(&empt[s0], ((it / TSTAGES) - 1) & 1);
                if (t >= g.ntiles) { tile_slot[s0] = -1; mbar_arrive(&full[s0]); break; }     // sentinel: no more tiles
                tile_slot[s0] = t;
                int bm, bn; tile_of(t, bm, bn);
                const int m0 = bm * TBM, n0 = bn * TBN;
                for (int kt = 0; kt < nk; kt++, it++) {
                    const int s = it % TSTAGES;
                    if (kt > 0 && it >= TSTAGES) mbar_wait(&empt[s], ((it / TSTAGES) - 1) & 1);
                    unsigned char* dst = smem_raw + (size_t)s * kTStageBytes;
                    mbar_expect_tx(&full[s], kTStageBytes);
#pragma unroll
                    for (int kk = 0; kk < TKB / 4; kk++) {
                        tma_load_2d(dst + kk * (TBM * 32), &mapA, kt * TKB + kk * 4, m0, &full[s]);
                        tma_load_2d(dst + kTABytes + kk * (TBN * 32), &mapB, kt * TKB + kk * 4, n0, &full[s]);
                    }
                }
            }
        }
        return;
    }

    const int wm = warp, wn = 0;            // 4 consumer warps stacked along m: warp tile 32 x 64
    const int lr = lane >> 2, lk = lane & 3;
    int it = 0;
    for (;;) {
        const int s0 = it % TSTAGES;
        mbar_wait(&full[s0], (it / TSTAGES) & 1);
        const int t = tile_slot[s0];
        if (t < 0) break;
        int bm, bn; tile_of(t, bm, bn);
        const int m0 = bm * TBM, n0 = bn * TBN;
        double acc[4][8][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kt = 0; kt < nk; kt++, it++) {
            const int s = it % TSTAGES;
            if (kt > 0) mbar_wait(&full[s], (it / TSTAGES) & 1);
            const double* a = reinterpret_cast<const double*>(smem_raw + (size_t)s * kTStageBytes);
            const double* b = a + TBM * TKB;
#pragma unroll
            for (int kk = 0; kk < TKB / 4; kk++) {
                double af[4], bf[8];
#pragma unroll
                for (int i = 0; i < 4; i++) af[i] = a[(kk * TBM + wm * 32 + i * 8) * 4 + lane];     // [(row)*4 + k], lane = (row%8)*4 + k
#pragma unroll
                for (int j = 0; j < 8; j++) bf[j] = b[(kk * TBN + wn * 64 + j * 8) * 4 + lane];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empt[s]);
        }

#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = m0 + wm * 32 + i * 8 + lr;
            if (row >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int col = n0 + wn * 64 + j * 8 + 2 * lk;
                if (col >= g.N) continue;
                double* p = g.C + (int64_t)row * g.ldc + col;
                const double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
                const bool ok0 = !g.lower_only || col <= row;
                const bool ok1 = (col + 1 < g.N) && (!g.lower_only || col + 1 <= row);
                if (ok0 && ok1) {
                    double2 o = make_double2(v0, v1);
                    if (g.accumulate) { const double2 c = *reinterpret_cast<double2*>(p); o.x += c.x; o.y += c.y; }
                    *reinterpret_cast<double2*>(p) = o;
                } else {
                    if (ok0) p[0] = g.accumulate ? p[0] + v0 : v0;
                    if (ok1) p[1] = g.accumulate ? p[1] + v1 : v1;
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x K row-major fp64 operand, leading dimension ld (doubles); box = {4 k, 128 rows}
static bool make_operand_map(CUtensorMap* map, const double* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld % 2) != 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {4, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// C (+)= alpha A B^T for one matrix (no batching).  Returns BOGP_OK, or 1 if the TMA path is not
// applicable (caller falls back to the cp.async kernel).
int launch_gemm_tma_nt(bogp_ctx* ctx, const GemmArgs& g) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return BOGP_OK;
    alignas(64) CUtensorMap mapA, mapB;
    if (!make_operand_map(&mapA, g.A, g.M, g.K, g.lda, TBM) || !make_operand_map(&mapB, g.B, g.N, g.K, g.ldb, TBN)) return 1;
    static DeviceOnce configured;
    if (configured.need(ctx->device)) {
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(gemm_tma_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem));
    }
    const int nbm = (g.M + TBM - 1) / TBM, nbn = (g.N + TBN - 1) / TBN;
    int ntiles = nbm * nbn;
    if (g.lower_only) { ntiles = 0; for (int r = 0; r < nbm; r++) ntiles += ((r + 1) * (TBM / TBN) < nbn) ? (r + 1) * (TBM / TBN) : nbn; }
    // tile counters: a ring of 16 words, one per launch (launches on one stream are ordered; the ring keeps
    // launches that overlap on different streams apart)
    static int ring = 0;
    int* counter = ctx->d_flags + 32 + (ring++ & 15);
    BOGP_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), ctx->stream));
    TmaGemmArgs a{g.C, g.ldc, g.M, g.N, g.K, g.alpha, g.accumulate, g.lower_only, nbm, nbn, ntiles, counter};
    const int ctas = ntiles < 2 * ctx->sm_count ? ntiles : 2 * ctx->sm_count;
    gemm_tma_nt_kernel<<<ctas, (TCONSUMERS + 1) * 32, kTSmem, ctx->stream>>>(mapA, mapB, a);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

}  // namespace bogp
