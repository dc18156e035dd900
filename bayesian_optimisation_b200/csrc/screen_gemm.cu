// Arg-max-only sweeps over a GRID: the posterior means of all candidates from fp64 GEMMs on the DMMA path.
//
// The screen of acquire.cu needs mu(c) = sum_j alpha_j k_*(x_j, c) for EVERY candidate (N operations each, with an exp
// in the panel formulation).  On a grid the kernel function factorises over the axes (acquire_i8.cuh: per-axis factor
// tables), so with the flat index split as c = p * T + t  (p: setting of the leading d - tt axes, t: setting of the tt
// trailing ones, T = their number of grid points)
//
//      mu[p, t] = sum_j G[p, j] F[t, j],     G[p, j] = prod_{leading k} f_k[j, p_k],   F[t, j] = alpha_j prod_{trailing k} f_k[j, t_k]
//
// is a plain matrix product (P x N) (N x T): ONE multiply-add per (candidate, measurement) on the FP64 tensor path instead
// of a squared distance, an exp and a multiply-add.  G is never stored: the GEMM forms its A tiles on the fly from the
// table rows (gemm_f64.cuh, A_GEN), 2 (kl - 1) multiplications per 32 x BN multiply-adds.
//
// The GEMM's mu differs from the mu of the exact kernels by rounding only (different association and summation order):
// |mu_gemm - mu_exact| <= (2 n_pad + 16) u sum_j |alpha_j| k_j <= eps := (2 n_pad + 16) 2^-53 1.0002 |alpha|_1, so the screen
// uses mu_gemm - eps, a rigorous lower bound of the exact mean (both acquisitions decrease with mu and grow with sigma,
// sigma^2 <= prior): a candidate is dropped only if A(mu_gemm - eps, sqrt(prior)) < best exact score so far.  Survivors
// are scored by the exact kernels from their flat indices through the same factor tables (acquire_i8.cuh,
// panel_tile_scattered), in one launch of the fused persistent kernel whatever their number, so the returned
// (score, index) is exactly the one of the unscreened sweep.                 point_selector.py:90-91,204-207
#include "common.cuh"
#include "fit.cuh"
#include "gemm_f64.cuh"
#include <cuda_fp16.h>

namespace bogp {

constexpr int kGsSeed = 4096;            // candidates of the strided seed sample scored before the first chunk
constexpr int kGsBatch = 4;              // chunks whose survivors are scored together by one exact pass

struct GsGeom {
    const double* ft; int toffT[BOGP_MAX_DIM]; int len[BOGP_MAX_DIM];
    int dim, kl, n, n_pad; long long ttot;
};

// out[r, j] = (alpha_j) * (((1 f_ka) f_ka+1) ... f_kb-1) for setting r of the axes [ka, kb), rows over j.
// grid (settings, n_pad / 256): F (with alpha) over the trailing axes, G (without) over the leading ones.
__global__ void __launch_bounds__(256) gs_prod_kernel(GsGeom g, int ka, int kb, const double* __restrict__ alpha, double* __restrict__ out) {
    const int j = blockIdx.y * 256 + threadIdx.x;
    long long t = blockIdx.x;
    int dig[BOGP_MAX_DIM];
    for (int k = kb - 1; k >= ka; k--) { dig[k] = (int)(t % g.len[k]); t /= g.len[k]; }
    double v = 1.0;
    for (int k = ka; k < kb; k++) v *= g.ft[g.toffT[k] + (int64_t)dig[k] * g.n_pad + j];
    if (alpha) v = j < g.n ? alpha[j] * v : 0.0;
    out[(int64_t)blockIdx.x * g.n_pad + j] = v;
}

// The same product without alpha, rounded to fp32, padding rows zeroed: the B operand of the max-times product below.
__global__ void __launch_bounds__(256) gs_prod32_kernel(GsGeom g, int ka, int kb, float* __restrict__ out) {
    const int j = blockIdx.y * 256 + threadIdx.x;
    long long t = blockIdx.x;
    int dig[BOGP_MAX_DIM];
    for (int k = kb - 1; k >= ka; k--) { dig[k] = (int)(t % g.len[k]); t /= g.len[k]; }
    double v = 1.0;
    for (int k = ka; k < kb; k++) v *= g.ft[g.toffT[k] + (int64_t)dig[k] * g.n_pad + j];
    out[(int64_t)blockIdx.x * g.n_pad + j] = j < g.n ? (float)v : 0.0f;
}

// Nearest-measurement bound of the posterior variance.  For ANY measurement j, sigma^2(c) <= prior - k_j(c)^2 / K_jj (the
// variance given one measurement bounds the variance given all: k^T K^-1 k >= k_j^2 / K_jj for positive definite K), so with
// m(c) = max_j k_j(c) the screen may use sigma_ub(c) = sqrt(prior - m^2 / K_jj) instead of sqrt(prior) -- decisive for
// explore * sigma - mu with a large explore, whose bound is otherwise loose exactly where mu is most negative (near the
// measurements).  On a grid m[p, t] = max_j G[p, j] F1[t, j] is a MAX-TIMES matrix product; it runs in fp32 on the CUDA
// cores (64 products and 64 maxima per thread and k; the result only has to be a lower bound of the true maximum, so it
// is scaled by 1 - 2^-21 afterwards, eight times the three roundings involved).  A[r, k] is generated from one or two
// stored fp64 product tables (the stored G, or the two composite tables of the generated mode) and rounded to fp32.
struct GsMaxArgs {
    const double* tab0; const double* tab1; long long radix1;   // A[r, k] = tab0[d0 * ld + k] (* tab1[d1 * ld + k]); p = d0 * radix1 + d1 (tab1 null: d0 = p)
    long long ld;
    const float* F1;                                            // [T][ld]
    float* out; int ldo;                                        // out[(p - p0) * ldo + t]
    long long p0; int M, N, K;
};

constexpr int kMxKB = 16;
// Packed-half arithmetic: two (product, maximum) pairs per instruction pair (HMUL2 + HMNMX2).  The operands are rounded
// to fp16 (a duplicated into both halves, two consecutive columns of B per word), three roundings of 2^-11 each in the
// normal range and of at most 2^-25 absolute below it, so the screen uses max(0, m (1 - 2^-9) - 2^-22) as its lower bound of
// the true maximum (GsScreenArgs.kmax_scale / kmax_sub): 0.2 % looser than fp32, far below what the bound gains.
template <bool TWO>      // TWO: A is the product of two table rows (generated mode), else one stored row
__global__ void __launch_bounds__(256, 2) gs_kmax_kernel(GsMaxArgs a) {
    __shared__ __align__(16) __half2 sA[2][kMxKB][128];       // (a, a)
    __shared__ __align__(16) __half2 sB[2][kMxKB][64];        // (b_2c, b_2c+1)
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    // loader: thread -> (row / column tid / 2, 8 consecutive k)
    const int lr = tid >> 1, lk = (tid & 1) * 8;
    const long long p = a.p0 + m0 + lr;
    const bool rowok = m0 + lr < a.M, colok = n0 + lr < a.N;
    const double* r0 = a.tab0 + (TWO ? p / a.radix1 : p) * a.ld + lk;
    const double* r1 = TWO ? a.tab1 + (p % a.radix1) * a.ld + lk : nullptr;
    const float* rb = a.F1 + (long long)(n0 + lr) * a.ld + lk;
    double2 pa[4], pb[4]; float4 pf[2];
    auto fetch = [&](int k0) {
        if (rowok) {
#pragma unroll
            for (int i = 0; i < 4; i++) pa[i] = __ldg(reinterpret_cast<const double2*>(r0 + k0) + i);
            if (TWO) {
#pragma unroll
                for (int i = 0; i < 4; i++) pb[i] = __ldg(reinterpret_cast<const double2*>(r1 + k0) + i);
            }
        }
        if (colok) { pf[0] = __ldg(reinterpret_cast<const float4*>(rb + k0)); pf[1] = __ldg(reinterpret_cast<const float4*>(rb + k0) + 1); }
    };
    auto stash = [&](int buf) {
        float va[8], vb[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double x = rowok ? pa[i].x : 0.0, y = rowok ? pa[i].y : 0.0;
            if (TWO && rowok) { x *= pb[i].x; y *= pb[i].y; }
            va[2 * i] = (float)x; va[2 * i + 1] = (float)y;
        }
        vb[0] = pf[0].x; vb[1] = pf[0].y; vb[2] = pf[0].z; vb[3] = pf[0].w; vb[4] = pf[1].x; vb[5] = pf[1].y; vb[6] = pf[1].z; vb[7] = pf[1].w;
        __half* sBh = reinterpret_cast<__half*>(&sB[buf][0][0]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            sA[buf][lk + i][lr] = __float2half2_rn(va[i]);
            sBh[(lk + i) * 128 + lr] = __float2half_rn(colok ? vb[i] : 0.0f);
        }
    };
    __half2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __float2half2_rn(0.0f);
    fetch(0);
    stash(0);
    __syncthreads();
    const int nk = a.K / kMxKB;
    for (int kt = 0; kt < nk; kt++) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * kMxKB);
#pragma unroll
        for (int k = 0; k < kMxKB; k++) {
            const uint4 a0 = *reinterpret_cast<const uint4*>(&sA[buf][k][ty * 8]), a1 = *reinterpret_cast<const uint4*>(&sA[buf][k][ty * 8 + 4]);
            const uint4 b0 = *reinterpret_cast<const uint4*>(&sB[buf][k][tx * 4]);
            const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const uint32_t bw[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    acc[i][j] = __hmax2(acc[i][j], __hmul2(*reinterpret_cast<const __half2*>(&aw[i]), *reinterpret_cast<const __half2*>(&bw[j])));
        }
        if (kt + 1 < nk) stash(buf ^ 1);           // the other buffer was last read in iteration kt - 1, before the barrier below
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int row = m0 + ty * 8 + i;
        if (row >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = n0 + tx * 8 + 2 * j;
            const float2 v = __half22float2(acc[i][j]);
            if (col < a.N) a.out[(long long)row * a.ldo + col] = v.x;
            if (col + 1 < a.N) a.out[(long long)row * a.ldo + col + 1] = v.y;
        }
    }
}

// Generated mode with room for a chunk of G: G[p - p0, j] = C1[p / P2, j] * C2[p % P2, j] written out (rows over j), so that the
// mean GEMM can run in the TMA-fed kernel on stored operands.  grid (rows of the chunk, n_pad / 256)
__global__ void __launch_bounds__(256) gs_compose_kernel(const double* __restrict__ c1, const double* __restrict__ c2, long long P2, long long p0,
                                                         int64_t ld, double* __restrict__ G) {
    const int j = blockIdx.y * 256 + threadIdx.x;
    const long long p = p0 + blockIdx.x;
    G[(int64_t)blockIdx.x * ld + j] = c1[(p / P2) * ld + j] * c2[(p % P2) * ld + j];
}

// mu[p - p0, t] = sum_j G[p, j] F[t, j] with G generated on the fly (gemm_f64.cuh, A_GEN): G[p, j] = (((1 f_0) f_1) ... f_kl-1)
// from the transposed table rows, never stored.  grid (ceil(T / BN), ceil(prefixes / 128)), 256 threads.
template <int BN>
__global__ void __launch_bounds__(256) gs_mu_gemm_kernel(GemmArgs g, GemmGenA gen, GsGeom geo, long long p0) {
    extern __shared__ __align__(16) double smem[];
    using S = GemmSmem<128, BN>;
    unsigned short* dg = reinterpret_cast<unsigned short*>(smem + S::bytes / sizeof(double));
    const int m0 = blockIdx.y * 128;
    if (threadIdx.x < 128) {
        long long p = p0 + m0 + threadIdx.x;
        for (int k = geo.kl - 1; k >= 0; k--) { dg[threadIdx.x * 16 + k] = (unsigned short)(p % geo.len[k]); p /= geo.len[k]; }
    }
    __syncthreads();
    GemmArgs gg = g;
    gg.gen = &gen;
    gemm_tile<128, BN, A_GEN, B_NK, K_ALL>(gg, nullptr, g.B, g.C, m0, blockIdx.x * BN, smem);
}

template <int BN>
static int launch_mu_gemm(bogp_ctx* ctx, const GemmArgs& m, const GemmGenA& gen, const GsGeom& geo, long long p0) {
    static DeviceOnce configured;
    constexpr size_t smem = GemmSmem<128, BN>::bytes + 128 * 16 * 2;
    if (configured.need(ctx->device)) {
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(gs_mu_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(gs_mu_gemm_kernel<BN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    gs_mu_gemm_kernel<BN><<<dim3((m.N + BN - 1) / BN, (m.M + 127) / 128), 256, smem, ctx->stream>>>(m, gen, geo, p0);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

__global__ void __launch_bounds__(256) gs_alpha_l1_kernel(const double* __restrict__ alpha, int n, double* __restrict__ out) {
    __shared__ double ws[8];
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) s += fabs(alpha[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { for (int w = 1; w < 8; w++) s += ws[w]; out[0] = s; }
}

__global__ void __launch_bounds__(256) gs_seed_kernel(long long c_begin, long long total, int count, long long* __restrict__ idx, int* __restrict__ n_out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < count) idx[i] = seed_index((unsigned long long)i, c_begin, total);
    if (i == 0) n_out[0] = count;
}

struct GsScreenArgs {
    const double* mu; long long base; int ldc; long long ttot;    // mu[(c - base) / T * ldc + (c - base) % T]
    long long lo, hi;                                             // candidates of this chunk inside the requested range
    const double* alpha_l1; double eps_factor;
    const float* kmax; int ldk; double prior, kjj_inv;           // nearest-measurement variance bound (or null)
    int kind; double explore, f_best, sigma_max;
    const bogp_result* best;
    long long* surv_idx; int* count;
    unsigned long long* stats;
};

__global__ void __launch_bounds__(256) gs_screen_kernel(GsScreenArgs a) {
    const long long c = a.lo + (long long)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false;
    if (c < a.hi) {
        const long long r = c - a.base;
        const long long pl = r / a.ttot;
        const double mu = a.mu[pl * a.ldc + (r - pl * a.ttot)] - a.eps_factor * a.alpha_l1[0];      // rigorous lower bound of the exact mean
        double smax = a.sigma_max;
        if (a.kmax) {      // sigma^2 <= prior - m^2 / K_jj, m a lower bound of max_j k_j; 1e-8 covers the rounding of the computed sigma^2 (~1e-13)
            double m = (double)a.kmax[pl * a.ldk + (r - pl * a.ttot)] * (1.0 - 0.001953125) - 2.384185791015625e-07;     // fp16 max-times: (1 - 2^-9) m - 2^-22
            m = m > 0.0 ? m : 0.0;
            const double s = sqrt(a.prior - m * m * a.kjj_inv + 1e-8);
            smax = s < smax ? s : smax;
        }
        double bound = acquisition_value(a.kind, mu, smax, a.explore, a.f_best);
        if (a.kind == BOGP_ACQ_EI) bound += 1e-12 * (fabs(a.f_best - mu) + a.sigma_max);
        keep = !(bound < a.best->score);                          // NaN means and bounds are kept: the exact kernels decide
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const unsigned inrange = __ballot_sync(0xffffffffu, c < a.hi);
    if (lane == 0 && inrange) { atomicAdd(&a.stats[0], (unsigned long long)__popc(inrange)); if (m) atomicAdd(&a.stats[1], (unsigned long long)__popc(m)); }
    if (m == 0u) return;
    int basepos = 0;
    if (lane == __ffs(m) - 1) basepos = atomicAdd(a.count, __popc(m));
    basepos = __shfl_sync(0xffffffffu, basepos, __ffs(m) - 1);
    if (keep) a.surv_idx[basepos + __popc(m & ((1u << lane) - 1u))] = c;      // capacity = candidates of a chunk: cannot overflow
}

static bool gs_geometry(const AcqChunk& tab, GsGeom& g) {
    if (!tab.ft || tab.points || tab.tt < 1) return false;
    g.ft = tab.ft; g.dim = tab.dim; g.kl = tab.dim - tab.tt; g.n_pad = tab.n_pad; g.n = tab.n;
    long long t = 1;
    for (int k = 0; k < BOGP_MAX_DIM; k++) { g.toffT[k] = tab.toffT[k]; g.len[k] = tab.len[k]; }
    for (int k = g.kl; k < g.dim; k++) t *= tab.len[k];
    g.ttot = t;
    if (tab.dim < 2) return false;
    for (int k = 0; k < tab.dim; k++) if (tab.len[k] > 65535) return false;  // digits travel as 16-bit values
    return true;
}

int gemm_screen_sweep(bogp_ctx* ctx, const bogp_fit* fit, const AcqChunk& tab, int64_t c_begin, int64_t c_end, int kind, double explore,
                      double f_best, double prior_diag, void* d_workspace, size_t workspace_bytes, bogp_result* d_result) {
    GsGeom g{};
    if (!gs_geometry(tab, g)) return 1;
    NvtxRange nvtx("bogp screened grid sweep: mean GEMM + bound + exact survivors");
    const int64_t n_pad = tab.n_pad;
    // workspace: [ring of the fused exact pass][stored operands G, F (stored mode)][mu chunk][survivor indices of a batch of chunks][count]
    const size_t ring = (fused_workspace_bytes(n_pad) + 255) / 256 * 256;
    if (workspace_bytes < ring + ((size_t)1 << 20)) return 1;
    const size_t avail = workspace_bytes - ring - 4096;

    // Split of the axes for the GEMM, c = p * T + t.  Stored mode: the split that minimises the two operand matrices
    // G (settings of the leading axes x N) and F (settings of the trailing axes x N), both built once per sweep -- a plain
    // GEMM with stored operands, if they fit half of the workspace.  Otherwise generated mode: the trailing axes of the
    // factor tables (T <= 256, F in the table reserve), G formed tile by tile inside the GEMM.
    int ks = 0; long long Ps = 0, Ts = 0;
    {
        long long tot = 1;
        for (int k = 0; k < g.dim; k++) tot *= g.len[k];
        long long P = 1, bestsum = -1;
        for (int k = 1; k < g.dim; k++) {
            P *= g.len[k - 1];
            const long long Tk = tot / P;
            if (Tk > 0x7fffffffLL || P > 0x7fffffffLL) continue;
            if (bestsum < 0 || P + Tk < bestsum) { bestsum = P + Tk; ks = k; Ps = P; Ts = Tk; }
        }
        if (bestsum < 0 || (size_t)bestsum * n_pad * 8 > avail / 2) ks = 0;
    }
    const bool stored = ks > 0;
    // generated mode: as many trailing axes as keep T <= 1024 columns (full 128-column tiles, F of a few tens of MB)
    int kg = g.dim - 1; long long Tg = g.len[g.dim - 1];
    while (kg > 1 && Tg * g.len[kg - 1] <= 1024) { kg--; Tg *= g.len[kg]; }
    const long long T = stored ? Ts : Tg;
    const int kl = stored ? ks : kg;
    const int ldc = (int)(T + (T & 1));                            // even row stride: the GEMM stores pairs
    char* base = static_cast<char*>(d_workspace);
    double* Gs = reinterpret_cast<double*>(base + ring);
    double* Fs = stored ? Gs + (size_t)Ps * n_pad : Gs;
    if (((size_t)(stored ? Ps : 0) + (size_t)T) * n_pad * 8 + ((size_t)1 << 20) > avail) return 1;
    char* rest = reinterpret_cast<char*>(Fs + (size_t)T * n_pad);
    const size_t rest_bytes = avail - (size_t)(rest - (base + ring));
    // Generated mode: the leading axes are folded into TWO composite axes whose product rows are stored (C1 over the
    // axes [0, ka), C2 over [ka, kl)), so a generated A element costs two loads and one multiplication whatever kl is.
    // If the composite tables do not fit, the per-axis table rows are used directly (kl loads per element).
    GemmGenA gen{};
    GsGeom gg = g; gg.kl = kl;
    gen.tab = g.ft; gen.kl = kl; gen.ld = n_pad;
    for (int k = 0; k < 16; k++) gen.toff[k] = g.toffT[k];
    int ka = 0; long long P1 = 0, P2 = 0;
    if (!stored && kl >= 2) {
        long long tot = 1;
        for (int k = 0; k < kl; k++) tot *= g.len[k];
        long long P = 1, bestsum = -1;
        for (int k = 1; k < kl; k++) {
            P *= g.len[k - 1];
            if (P > 65535 || tot / P > 65535) continue;
            if (bestsum < 0 || P + tot / P < bestsum) { bestsum = P + tot / P; ka = k; P1 = P; P2 = tot / P; }
        }
        if (bestsum < 0 || (size_t)bestsum * n_pad * 8 > rest_bytes / 2 || (size_t)P1 * n_pad > 0x7fffffffULL) ka = 0;
    }
    double* comp = reinterpret_cast<double*>(rest);
    size_t comp_bytes = 0;
    if (ka > 0) {
        comp_bytes = ((size_t)(P1 + P2) * n_pad * 8 + 255) / 256 * 256;
        gen.tab = comp; gen.kl = 2; gen.toff[0] = 0; gen.toff[1] = (int)((size_t)P1 * n_pad);
        gg.kl = 2; gg.len[0] = (int)P1; gg.len[1] = (int)P2;
    }
    // nearest-measurement variance bound (gs_kmax_kernel): for explore * sigma - mu, whose sigma <= sigma_max bound is loose
    // near the measurements; EI screens well without it (6e-5 of the 10^8 grid survive) and skips the extra pass
    GsMaxArgs mx{};
    const bool stored_ok = stored, comp_ok = ka > 0, axis_ok = !stored && ka == 0 && kl == 1;
    bool want_kmax = kind == BOGP_ACQ_LCB && explore > 0.0 && (stored_ok || comp_ok || axis_ok);
    size_t f1_bytes = want_kmax ? ((size_t)T * n_pad * 4 + 255) / 256 * 256 : 0;
    if (want_kmax && f1_bytes + comp_bytes + ((size_t)8 << 20) > rest_bytes) { want_kmax = false; f1_bytes = 0; }
    const int ldo = (int)((T + 3) / 4 * 4);
    // rows (settings of the leading axes) per chunk: at least ~2.5 M candidates, and a number of 128-row tiles that fills
    // whole waves of the SMs (the smallest row count with >= 88 % of the last wave used, else the best one that fits)
    // generated mode with composite tables: a chunk of G is written out for the TMA-fed GEMM if rows of 8 n_pad bytes fit
    const bool compose = !stored && ka > 0 && rest_bytes - comp_bytes - f1_bytes > ((size_t)1024 * n_pad * 8 + ((size_t)256 << 20));
    const size_t per_row = (size_t)ldc * 8 + (size_t)kGsBatch * T * 8 + (want_kmax ? (size_t)ldo * 4 : 0) + (compose ? (size_t)n_pad * 8 : 0);
    comp_bytes += f1_bytes;                                        // (the fp32 operand sits right behind the composite tables)
    const long long col_tiles = (T + (T <= 64 ? 63 : 127)) / (T <= 64 ? 64 : 128);
    long long Pc = 128; double beste = 0.0;
    for (long long r = 1; r <= 1024; r++) {
        if ((size_t)(r * 128) * per_row > rest_bytes - comp_bytes) break;
        const long long tiles = r * col_tiles, waves = (tiles + ctx->sm_count - 1) / ctx->sm_count;
        const double e = (double)tiles / (double)(waves * ctx->sm_count);
        if (e > beste + 1e-9) { beste = e; Pc = r * 128; }
        if (e >= 0.88 && r * 128 * T >= 2500000) { Pc = r * 128; break; }
        if (compose && (r + 1) * 128 * (size_t)n_pad * 8 > ((size_t)512 << 20)) { if (r * 128 * T >= 1000000) { Pc = r * 128; break; } }   // keep the G chunk under 512 MB (the TMA kernel schedules tiles dynamically)
    }
    if ((size_t)Pc * per_row > rest_bytes - comp_bytes) return 1;
    float* F1 = reinterpret_cast<float*>(rest + comp_bytes - f1_bytes);
    double* mu = reinterpret_cast<double*>(rest + comp_bytes);
    long long* surv = reinterpret_cast<long long*>(mu + (size_t)Pc * ldc);
    float* kmx = reinterpret_cast<float*>(surv + (size_t)kGsBatch * Pc * T);
    double* Gc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(kmx + (want_kmax ? (size_t)Pc * ldo : 0)) + 255) / 256 * 256);
    int* count = reinterpret_cast<int*>(Gc + (compose ? (size_t)Pc * n_pad : 0));
    const long long cap = kGsBatch * Pc * T;                       // survivors of a batch of chunks: at most all of their candidates
    if (cap < kGsSeed) return 1;

    cudaStream_t st = ctx->stream;
    double* l1 = ctx->d_scalars + 20;
    gs_alpha_l1_kernel<<<1, 256, 0, st>>>(fit_alpha(fit), (int)fit_n(fit), l1); BOGP_LAUNCH_CHECK(ctx);
    gs_prod_kernel<<<dim3((unsigned)T, (unsigned)(n_pad / 256)), 256, 0, st>>>(g, kl, g.dim, fit_alpha(fit), Fs); BOGP_LAUNCH_CHECK(ctx);
    if (stored) { gs_prod_kernel<<<dim3((unsigned)Ps, (unsigned)(n_pad / 256)), 256, 0, st>>>(g, 0, kl, nullptr, Gs); BOGP_LAUNCH_CHECK(ctx); }
    if (ka > 0) {
        gs_prod_kernel<<<dim3((unsigned)P1, (unsigned)(n_pad / 256)), 256, 0, st>>>(g, 0, ka, nullptr, comp); BOGP_LAUNCH_CHECK(ctx);
        gs_prod_kernel<<<dim3((unsigned)P2, (unsigned)(n_pad / 256)), 256, 0, st>>>(g, ka, kl, nullptr, comp + (size_t)P1 * n_pad); BOGP_LAUNCH_CHECK(ctx);
    }
    if (want_kmax) {
        gs_prod32_kernel<<<dim3((unsigned)T, (unsigned)(n_pad / 256)), 256, 0, st>>>(g, kl, g.dim, F1); BOGP_LAUNCH_CHECK(ctx);
        mx.tab0 = stored_ok ? Gs : (comp_ok ? comp : g.ft + g.toffT[0]);
        mx.tab1 = comp_ok && !stored_ok ? comp + (size_t)P1 * n_pad : nullptr; mx.radix1 = P2 > 0 ? P2 : 1;
        mx.ld = n_pad; mx.F1 = F1; mx.out = kmx; mx.ldo = ldo; mx.N = (int)T; mx.K = (int)n_pad;
    }

    // exact scoring of the candidates listed in surv[0 .. *count): one launch of the fused persistent kernel
    auto exact_pass = [&](int fold) -> int {
        AcqChunk a = tab;
        a.points = nullptr; a.cross_jitter = 0.0;
        a.alpha = fit_alpha(fit); a.wp = fit_wp(fit); a.wq = fit_wq(fit); a.wscale = fit_wscale(fit);
        a.c0 = 0; a.c_end = cap; a.cur = cap; a.S = 0;
        a.d_count = count; a.idx_list = surv;
        FusedFinal f{nullptr, nullptr, nullptr, surv, kind, explore, f_best, prior_diag, d_result, fold};
        const int rc = launch_acquire_fused(ctx, a, f, base, ring, st);
        if (rc == 1) { set_error("bogp_acquire: workspace too small for the exact pass of a screened sweep"); return BOGP_ERR_WORKSPACE; }
        return rc;
    };
    // seed: a low-discrepancy sample over the whole range is scored first, so that the running best is already high when the first
    // chunk is screened
    // (ctx->global_seed: the sample spans the whole grid -- this sweep is one shard of a sharded arg-max and all shards
    // then screen against the same floor)
    long long gtot = 1;
    for (int k = 0; k < g.dim; k++) gtot *= g.len[k];
    const long long s_begin = ctx->global_seed ? 0 : c_begin;
    const long long total = ctx->global_seed ? gtot : c_end - c_begin;
    const int nseed = (int)(total < kGsSeed ? total : kGsSeed);
    gs_seed_kernel<<<(nseed + 255) / 256, 256, 0, st>>>(s_begin, total, nseed, surv, count); BOGP_LAUNCH_CHECK(ctx);
    int rc = exact_pass(0); if (rc) return rc;

    GsScreenArgs sa{};
    sa.mu = mu; sa.ldc = ldc; sa.ttot = T; sa.alpha_l1 = l1;
    sa.eps_factor = (2.0 * (double)n_pad + 16.0) * 1.1102230246251565e-16 * 1.0002;
    sa.kind = kind; sa.explore = explore; sa.f_best = f_best; sa.sigma_max = sqrt(prior_diag);
    sa.best = d_result; sa.surv_idx = surv; sa.count = count;
    sa.kmax = want_kmax ? kmx : nullptr; sa.ldk = ldo; sa.prior = prior_diag; sa.kjj_inv = 1.0 / (1.0 + fit_jitter(fit));
    sa.stats = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 16);
    const long long p_begin = c_begin / T, p_end = (c_end + T - 1) / T;
    int pending = 0;                                               // whole chunks screened since the last exact pass
    long long sub_rows = (250000 + T - 1) / T;                     // rows per screened range, doubling
    BOGP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
    for (long long p0 = p_begin; p0 < p_end; p0 += Pc) {
        const long long pc = (p_end - p0 < Pc) ? (p_end - p0) : Pc;
        GemmArgs m{};
        m.A = stored ? Gs + (size_t)p0 * n_pad : nullptr; m.lda = n_pad; m.B = Fs; m.ldb = n_pad; m.C = mu; m.ldc = ldc;
        m.M = (int)pc; m.N = (int)T; m.K = (int)n_pad; m.alpha = 1.0; m.accumulate = 0; m.lower_only = 0;
        bool done = false;
        if (compose) {    // write the chunk of G out, then the TMA-fed kernel on stored operands
            gs_compose_kernel<<<dim3((unsigned)pc, (unsigned)(n_pad / 256)), 256, 0, st>>>(comp, comp + (size_t)P1 * n_pad, P2, p0, n_pad, Gc); BOGP_LAUNCH_CHECK(ctx);
            m.A = Gc;
            rc = launch_gemm_tma_nt(ctx, m);
            if (rc == 1) m.A = nullptr; else done = true;
        }
        if (done) {
        } else if (stored) {     // TMA-fed persistent NT kernel with dynamic tile scheduling (gemm_tma.cu), else the cp.async kernel
            rc = launch_gemm_tma_nt(ctx, m);
            if (rc == 1) rc = (T <= 64) ? launch_gemm<128, 64, A_MK, B_NK, K_ALL>(ctx, m, 1) : launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, m, 1);
        }
        else if (T <= 64) rc = launch_mu_gemm<64>(ctx, m, gen, gg, p0);
        else              rc = launch_mu_gemm<128>(ctx, m, gen, gg, p0);
        if (rc) return rc;
        if (want_kmax) {
            GsMaxArgs mq = mx;
            mq.p0 = p0; mq.M = (int)pc;
            if (compose && m.A == Gc) { mq.tab0 = Gc; mq.tab1 = nullptr; mq.p0 = 0; }      // the chunk of G that was just written out: one stored row per A row
            const dim3 mgrid((unsigned)((T + 127) / 128), (unsigned)((pc + 127) / 128));
            if (mq.tab1) gs_kmax_kernel<true><<<mgrid, 256, 0, st>>>(mq); else gs_kmax_kernel<false><<<mgrid, 256, 0, st>>>(mq);
            BOGP_LAUNCH_CHECK(ctx);
        }
        sa.base = p0 * T;
        // The chunk's candidates are screened in row ranges, each followed by the exact scoring of its survivors.  The running
        // best rises fastest at the beginning, so the ranges start small (~250 k candidates) and double after every exact pass
        // (which has a fixed cost of ~0.25 ms, the chain of the heaviest row block) until they cover whole chunks; then
        // kGsBatch chunks share one pass.
        for (long long r0 = 0; r0 < pc;) {
            const long long rows = (sub_rows < pc - r0) ? sub_rows : (pc - r0);
            sa.lo = sa.base + r0 * T > c_begin ? sa.base + r0 * T : c_begin;
            sa.hi = sa.base + (r0 + rows) * T < c_end ? sa.base + (r0 + rows) * T : c_end;
            if (sa.hi > sa.lo) { gs_screen_kernel<<<(unsigned)((sa.hi - sa.lo + 255) / 256), 256, 0, st>>>(sa); BOGP_LAUNCH_CHECK(ctx); }
            r0 += rows;
            const bool last = (r0 >= pc) && (p0 + Pc >= p_end);
            bool pass = last;
            if (sub_rows < Pc) { pass = true; sub_rows *= 2; }                     // still ramping up
            else if (r0 >= pc && ++pending == kGsBatch) pass = true;
            if (pass) {
                rc = exact_pass(1); if (rc) return rc;
                BOGP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
                pending = 0;
            }
        }
    }
    return BOGP_OK;
}

}  // namespace bogp
