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
// of a squared distance, an exp and a multiply-add -- 27 ms instead of 1.1 s for the 10^8-point grid at N = 4096.
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

namespace bogp {

constexpr int kGsSeed = 4096;            // candidates of the strided seed sample scored before the first chunk

struct GsGeom {
    const double* ft; int toffT[BOGP_MAX_DIM]; int len[BOGP_MAX_DIM];
    int dim, kl, n, n_pad; long long ttot;
};

// F[t, j] = alpha_j * (((1 f_kl) f_kl+1) ...), rows over j.  grid (n_pad / 256, T)
__global__ void __launch_bounds__(256) gs_fmat_kernel(GsGeom g, const double* __restrict__ alpha, double* __restrict__ F) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    long long t = blockIdx.y;
    int dig[BOGP_MAX_DIM];
    for (int k = g.dim - 1; k >= g.kl; k--) { dig[k] = (int)(t % g.len[k]); t /= g.len[k]; }
    double v = 1.0;
    for (int k = g.kl; k < g.dim; k++) v *= g.ft[g.toffT[k] + (int64_t)dig[k] * g.n_pad + j];
    F[(int64_t)blockIdx.y * g.n_pad + j] = j < g.n ? alpha[j] * v : 0.0;
}

// G[p - p0, j] = (((1 f_0) f_1) ... f_kl-1), rows over j.  grid (n_pad / 256, prefixes of the chunk)
__global__ void __launch_bounds__(256) gs_gmat_kernel(GsGeom g, long long p0, double* __restrict__ G) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    long long p = p0 + blockIdx.y;
    int dig[BOGP_MAX_DIM];
    for (int k = g.kl - 1; k >= 0; k--) { dig[k] = (int)(p % g.len[k]); p /= g.len[k]; }
    double v = 1.0;
    for (int k = 0; k < g.kl; k++) v *= g.ft[g.toffT[k] + (int64_t)dig[k] * g.n_pad + j];
    G[(int64_t)blockIdx.y * g.n_pad + j] = v;
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

__global__ void __launch_bounds__(256) gs_seed_kernel(long long c_begin, long long stride, int count, long long* __restrict__ idx, int* __restrict__ n_out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < count) idx[i] = c_begin + (long long)i * stride;
    if (i == 0) n_out[0] = count;
}

struct GsScreenArgs {
    const double* mu; long long base; int ldc; long long ttot;    // mu[(c - base) / T * ldc + (c - base) % T]
    long long lo, hi;                                             // candidates of this chunk inside the requested range
    const double* alpha_l1; double eps_factor;
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
        double bound = acquisition_value(a.kind, mu, a.sigma_max, a.explore, a.f_best);
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
    return t >= 1 && t <= 256;
}

size_t gemm_screen_f_doubles(const AcqChunk& tab) {
    GsGeom g{};
    return gs_geometry(tab, g) ? (size_t)g.ttot * tab.n_pad : 0;
}

int gemm_screen_sweep(bogp_ctx* ctx, const bogp_fit* fit, const AcqChunk& tab, int64_t c_begin, int64_t c_end, int kind, double explore,
                      double f_best, double prior_diag, void* d_workspace, size_t workspace_bytes, double* d_f, bogp_result* d_result) {
    GsGeom g{};
    if (!d_f || !gs_geometry(tab, g)) return 1;
    const int64_t n_pad = tab.n_pad;
    const long long T = g.ttot;
    const int ldc = (int)(T + (T & 1));                            // even row stride: the GEMM stores pairs
    // workspace: [ring of the fused exact pass][G chunk][mu chunk][survivor indices][count]
    const size_t ring = (fused_workspace_bytes(n_pad) + 255) / 256 * 256;
    if (workspace_bytes < ring + ((size_t)1 << 20)) return 1;
    const size_t avail = workspace_bytes - ring - 4096;
    const size_t per_prefix = (size_t)n_pad * 8 + (size_t)ldc * 8 + (size_t)T * 8;
    long long Pc = (long long)128 * ctx->sm_count;
    if ((size_t)Pc * per_prefix > avail) Pc = (long long)(avail / per_prefix);
    if (Pc > 65535) Pc = 65535;
    Pc = Pc / 128 * 128;
    if (Pc < 128) return 1;
    char* base = static_cast<char*>(d_workspace);
    double* G = reinterpret_cast<double*>(base + ring);
    double* mu = G + (size_t)Pc * n_pad;
    long long* surv = reinterpret_cast<long long*>(mu + (size_t)Pc * ldc);
    int* count = reinterpret_cast<int*>(surv + (size_t)Pc * T);
    const long long cap = Pc * T;                                  // survivors of a chunk: at most all of its candidates
    if ((size_t)(reinterpret_cast<char*>(count) - base) + 256 > workspace_bytes || cap < kGsSeed) return 1;

    cudaStream_t st = ctx->stream;
    double* l1 = ctx->d_scalars + 20;
    gs_alpha_l1_kernel<<<1, 256, 0, st>>>(fit_alpha(fit), (int)fit_n(fit), l1); BOGP_LAUNCH_CHECK(ctx);
    gs_fmat_kernel<<<dim3((unsigned)(n_pad / 256), (unsigned)T), 256, 0, st>>>(g, fit_alpha(fit), d_f); BOGP_LAUNCH_CHECK(ctx);

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
    // seed: a strided sample over the whole range is scored first, so that the running best is already high when the first
    // chunk is screened
    const long long total = c_end - c_begin;
    const int nseed = (int)(total < kGsSeed ? total : kGsSeed);
    gs_seed_kernel<<<(nseed + 255) / 256, 256, 0, st>>>(c_begin, total / nseed, nseed, surv, count); BOGP_LAUNCH_CHECK(ctx);
    int rc = exact_pass(0); if (rc) return rc;

    GsScreenArgs sa{};
    sa.mu = mu; sa.ldc = ldc; sa.ttot = T; sa.alpha_l1 = l1;
    sa.eps_factor = (2.0 * (double)n_pad + 16.0) * 1.1102230246251565e-16 * 1.0002;
    sa.kind = kind; sa.explore = explore; sa.f_best = f_best; sa.sigma_max = sqrt(prior_diag);
    sa.best = d_result; sa.surv_idx = surv; sa.count = count;
    sa.stats = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 16);
    const long long p_begin = c_begin / T, p_end = (c_end + T - 1) / T;
    for (long long p0 = p_begin; p0 < p_end; p0 += Pc) {
        const long long pc = (p_end - p0 < Pc) ? (p_end - p0) : Pc;
        gs_gmat_kernel<<<dim3((unsigned)(n_pad / 256), (unsigned)pc), 256, 0, st>>>(g, p0, G); BOGP_LAUNCH_CHECK(ctx);
        GemmArgs m{};
        m.A = G; m.lda = n_pad; m.B = d_f; m.ldb = n_pad; m.C = mu; m.ldc = ldc;
        m.M = (int)pc; m.N = (int)T; m.K = (int)n_pad; m.alpha = 1.0; m.accumulate = 0; m.lower_only = 0;
        rc = (T <= 64) ? launch_gemm<128, 64, A_MK, B_NK, K_ALL>(ctx, m, 1) : launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, m, 1);
        if (rc) return rc;
        BOGP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
        sa.base = p0 * T;
        sa.lo = sa.base > c_begin ? sa.base : c_begin;
        sa.hi = (p0 + pc) * T < c_end ? (p0 + pc) * T : c_end;
        if (sa.hi > sa.lo) {
            gs_screen_kernel<<<(unsigned)((sa.hi - sa.lo + 255) / 256), 256, 0, st>>>(sa); BOGP_LAUNCH_CHECK(ctx);
            rc = exact_pass(1); if (rc) return rc;
        }
    }
    return BOGP_OK;
}

}  // namespace bogp
