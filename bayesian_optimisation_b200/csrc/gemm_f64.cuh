// Generic fp64 GEMM on the FP64 tensor path (DMMA.8x8x4) for the fit:
//   C = alpha * op(A) * op(B) (+ C)
// Row-major operands, cp.async multi-stage staging into padded shared memory.
// Used for the Cholesky trailing update (SYRK), the panel solve (multiplication with the
// inverted diagonal block), the recursive triangular inverse and K^-1 = W^T W.
//
// Every accumulation runs in a fixed order (k ascending inside one thread), so results
// are bit-reproducible run to run and GPU to GPU -- the replicated Cholesky of the
// multi-GPU sweep depends on that.
#pragma once
#include "common.cuh"

namespace bogp {

enum GemmBLayout { B_NK = 0 /* B is N x K row-major: C = A * B^T */, B_KN = 1 /* B is K x N row-major */ };
enum GemmALayout { A_MK = 0 /* A is M x K row-major */, A_KM = 1 /* A is K x M row-major: C = A^T * B */,
                   A_GEN = 2 /* A[m, k] is not stored: it is the ordered product of per-axis table rows (GemmGenA) */ };

// Generated A operand (screen_gemm.cu): row m stands for setting p0 + m of the leading grid axes, A[m, k] =
// (((1 f_0[k]) f_1[k]) ... f_{kl-1}[k]) with f_a = table row tab[toff[a] + digit_a(m) * ld + k].  The digits of the tile's
// rows are expected as 16-bit values at smem + GemmSmem::bytes, [BM][16] (the calling kernel puts them there).
struct GemmGenA {
    const double* tab; int toff[16]; int kl; int64_t ld;
};
enum GemmKRange  { K_ALL = 0, K_GE_N = 1 /* B[k,n] == 0 for k < n */, K_LE_M = 2 /* A[m,k] == 0 for k > m */,
                   K_GE_MAXMN = 3 /* A^T*B with both lower triangular: k >= max(m, n) */ };

struct GemmArgs {
    const double* A; const double* B; double* C;
    int64_t lda, ldb, ldc;
    int64_t strideA, strideB, strideC;   // per inner batch index  (blockIdx.z % inner)
    int64_t strideA2, strideB2, strideC2; // per outer batch index (blockIdx.z / inner)
    int inner;                            // inner batch count (0 or 1 -> single-level batch over blockIdx.z)
    int M, N, K;
    double alpha;
    int accumulate;     // C += ...
    int lower_only;     // only tiles / elements with row >= col are produced (SYRK)
    const GemmGenA* gen; // A_GEN only (device pointer)
};

constexpr int GK = 32;          // k extent per stage
constexpr int GPADK = GK + 4;   // row stride (doubles) of an [rows][k] tile: stride = 4 (mod 16) -> conflict-free fragment loads

template <int BM, int BN>
struct GemmSmem {
    static constexpr int kAStage = (BM * GPADK > GK * (BM + 4)) ? BM * GPADK : GK * (BM + 4);
    static constexpr int kBStage = (BN * GPADK > GK * (BN + 4)) ? BN * GPADK : GK * (BN + 4);
    static constexpr size_t kStageBytes = size_t(kAStage + kBStage) * sizeof(double);
    // as many stages as fit into the 227 KB of one CTA, at most 4, at least 2
    static constexpr int kStages = (4 * kStageBytes <= 227 * 1024) ? 4 : ((3 * kStageBytes <= 227 * 1024) ? 3 : 2);
    static constexpr size_t bytes = size_t(kStages) * kStageBytes;
};

// One BM x BN output tile at (m0, n0) by the 256 threads of a CTA (8 warps arranged 4 (m) x 2 (n); warp
// tile (BM/4) x (BN/2)).  Callable from the plain GEMM kernel below and from fused multi-phase kernels.
template <int BM, int BN, int AL, int BL, int KR>
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, const double* A, const double* B, double* C, int m0, int n0, double* smem) {
    using S = GemmSmem<BM, BN>;
    double* sA = smem;
    constexpr int GSTAGES = S::kStages;
    double* sB = smem + GSTAGES * S::kAStage;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    constexpr int WM = BM / 4, WN = BN / 2, MI = WM / 8, NI = WN / 8;

    if (g.lower_only && n0 > m0 + BM - 1) return;

    int kbeg = 0, kend = g.K;
    if (KR == K_GE_N) kbeg = (n0 / GK) * GK;
    if (KR == K_LE_M) kend = min(g.K, m0 + BM);
    if (KR == K_GE_MAXMN) kbeg = (max(m0, n0) / GK) * GK;
    const int nk = (kend - kbeg + GK - 1) / GK;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kbeg + kt * GK;
        double* a = sA + stage * S::kAStage;
        double* b = sB + stage * S::kBStage;
        if (AL == A_GEN) {     // [BM][GPADK]: every thread forms 2 consecutive k of BM / 16 rows from the table rows (L1 / L2 resident)
            const unsigned short* dg = reinterpret_cast<const unsigned short*>(smem + S::bytes / sizeof(double));
            const GemmGenA& ga = *g.gen;
            for (int c = tid; c < BM * (GK / 2); c += 256) {
                const int r = c / (GK / 2), q = c % (GK / 2);
                double2 v = make_double2(1.0, 1.0);
                const int64_t kq = k0 + q * 2;
                for (int ax = 0; ax < ga.kl; ax++) {
                    const double2 f = __ldg(reinterpret_cast<const double2*>(ga.tab + ga.toff[ax] + (int64_t)dg[r * 16 + ax] * ga.ld + kq));
                    v.x *= f.x; v.y *= f.y;
                }
                if (m0 + r >= g.M || kq >= kend) v = make_double2(0.0, 0.0);
                *reinterpret_cast<double2*>(a + r * GPADK + q * 2) = v;
            }
        } else if (AL == A_MK) {      // [BM][GPADK], 8 chunks of 16 B per row
            for (int c = tid; c < BM * (GK / 2); c += 256) {
                int r = c / (GK / 2), q = c % (GK / 2);
                bool ok = (m0 + r < g.M) && (k0 + q * 2 < kend);
                const double* src = A + (int64_t)(ok ? m0 + r : 0) * g.lda + (ok ? k0 + q * 2 : 0);
                cp_async16(a + r * GPADK + q * 2, src, ok);
            }
        } else {               // A is K x M: tile [GK][BM+4]
            for (int c = tid; c < GK * (BM / 2); c += 256) {
                int r = c / (BM / 2), q = c % (BM / 2);
                bool ok = (k0 + r < kend) && (m0 + q * 2 < g.M);
                const double* src = A + (int64_t)(ok ? k0 + r : 0) * g.lda + (ok ? m0 + q * 2 : 0);
                cp_async16(a + r * (BM + 4) + q * 2, src, ok);
            }
        }
        if (BL == B_NK) {
            for (int c = tid; c < BN * (GK / 2); c += 256) {
                int r = c / (GK / 2), q = c % (GK / 2);
                bool ok = (n0 + r < g.N) && (k0 + q * 2 < kend);
                const double* src = B + (int64_t)(ok ? n0 + r : 0) * g.ldb + (ok ? k0 + q * 2 : 0);
                cp_async16(b + r * GPADK + q * 2, src, ok);
            }
        } else {
            for (int c = tid; c < GK * (BN / 2); c += 256) {
                int r = c / (BN / 2), q = c % (BN / 2);
                bool ok = (k0 + r < kend) && (n0 + q * 2 < g.N);
                const double* src = B + (int64_t)(ok ? k0 + r : 0) * g.ldb + (ok ? n0 + q * 2 : 0);
                cp_async16(b + r * (BN + 4) + q * 2, src, ok);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int lr = lane >> 2, lk = lane & 3;
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        {   // prefetch stage kt + GSTAGES - 1 (its buffer was consumed at iteration kt-1)
            int nxt = kt + GSTAGES - 1;
            if (nxt < nk) load_stage(nxt % GSTAGES, nxt);
            cp_async_commit();
        }
        const double* a = sA + (kt % GSTAGES) * S::kAStage;
        const double* b = sB + (kt % GSTAGES) * S::kBStage;
#pragma unroll
        for (int kk = 0; kk < GK / 4; kk++) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; i++) {
                int r = wm * WM + i * 8 + lr;
                af[i] = (AL != A_KM) ? a[r * GPADK + kk * 4 + lk] : a[(kk * 4 + lk) * (BM + 4) + r];
            }
#pragma unroll
            for (int j = 0; j < NI; j++) {
                int c = wn * WN + j * 8 + lr;
                bf[j] = (BL == B_NK) ? b[c * GPADK + kk * 4 + lk] : b[(kk * 4 + lk) * (BN + 4) + c];
            }
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: lane holds C[row][col..col+1]
#pragma unroll
    for (int i = 0; i < MI; i++) {
        const int row = m0 + wm * WM + i * 8 + lr;
        if (row >= g.M) continue;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            const int col = n0 + wn * WN + j * 8 + 2 * lk;
            if (col >= g.N) continue;
            double* p = C + (int64_t)row * g.ldc + col;
            double v0 = g.alpha * acc[i][j][0], v1 = g.alpha * acc[i][j][1];
            const bool ok0 = !g.lower_only || col <= row;
            const bool ok1 = (col + 1 < g.N) && (!g.lower_only || col + 1 <= row);
            if (ok0 && ok1) {
                double2 o = make_double2(v0, v1);
                if (g.accumulate) { double2 c = __ldcg(reinterpret_cast<const double2*>(p)); o.x += c.x; o.y += c.y; }   // L2-coherent read
                *reinterpret_cast<double2*>(p) = o;
            } else {
                if (ok0) p[0] = g.accumulate ? __ldcg(p) + v0 : v0;
                if (ok1) p[1] = g.accumulate ? __ldcg(p + 1) + v1 : v1;
            }
        }
    }
}

template <int BM, int BN, int AL, int BL, int KR>
__global__ void __launch_bounds__(256) gemm_f64_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    // Tiles with a restricted k range are scheduled heaviest-first: K_LE_M grows with m (row blocks in
    // reverse order), K_GE_N shrinks with n (the column block becomes the slow grid index, ascending).
    int bx = (int)blockIdx.x, by = (int)blockIdx.y;
    if (KR == K_LE_M) by = (int)(gridDim.y - 1 - blockIdx.y);
    if (KR == K_GE_N) {
        const int lin = (int)(blockIdx.y * gridDim.x + blockIdx.x);
        by = lin % (int)gridDim.y;
        bx = lin / (int)gridDim.y;
    }
    const int zi = g.inner > 1 ? (int)(blockIdx.z % g.inner) : (g.inner == 1 ? 0 : (int)blockIdx.z);
    const int zo = g.inner >= 1 ? (int)(blockIdx.z / g.inner) : 0;
    gemm_tile<BM, BN, AL, BL, KR>(g, g.A + zi * g.strideA + zo * g.strideA2, g.B + zi * g.strideB + zo * g.strideB2,
                                  g.C + zi * g.strideC + zo * g.strideC2, by * BM, bx * BN, smem);
}

template <int BM, int BN, int AL, int BL, int KR>
inline int launch_gemm(bogp_ctx* ctx, const GemmArgs& g, int batch) {
    auto kern = gemm_f64_kernel<BM, BN, AL, BL, KR>;
    static DeviceOnce configured;
    constexpr size_t smem = GemmSmem<BM, BN>::bytes;
    if (configured.need(ctx->device)) {
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return BOGP_OK;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batch);
    kern<<<grid, 256, smem, ctx->stream>>>(g);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

}  // namespace bogp
