// K4: fused acquisition sweep                        point_selector.py:81,90-98,204-207
//
// For a chunk of S candidates (a "super-tile"):
//   1. panel_kernel   k_*[j, c] = exp(-0.5 sum_k (x_jk - p_ck)^2 / ell_k^2), written once, in
//                     DMMA-fragment order, plus the partial posterior means alpha . k_*;
//                     explicit candidate blocks are staged with a bulk-TMA copy.
//   2. trigemm_kernel V = W k_* on the FP64 tensor path (W = L^-1 lower triangular, packed),
//                     operands streamed by bulk-TMA copies through a 5-stage mbarrier ring,
//                     fused epilogue: column sums of V^2 (never stores V).
//   3. finalize_kernel sigma^2 = prior - |V_c|^2, mu, LCB / EI, warp-level arg-max with the
//                     reference's tie rule (largest score, then smallest flat index).
//   4. merge_kernel   folds the block winners into the running (score, index) of the sweep.
//
// Every reduction has a fixed order that depends only on n_pad, so a candidate's score is
// bit-identical no matter which chunk, CTA or GPU scored it.
#include "common.cuh"
#include "fit.cuh"

namespace bogp {

struct CandDesc {
    const double* points;     // explicit: c_total x dim
    const double* axes;       // grid: concatenated axes
    int           len[BOGP_MAX_DIM];
    int           off[BOGP_MAX_DIM];
    int64_t       c_total;
    double        cross_jitter;
};

struct PanelArgs {
    CandDesc cand;
    const double* x_pad; const double* inv_ell2; const double* alpha;
    double* panel; double* mupart;
    int64_t c0, c_end;    // chunk start (global flat index), end of the requested range
    int64_t S;            // chunk capacity (stride of mupart)
    int n, n_pad, dim;
};

// grid (ceil(cur/64), n_pad/256), 256 threads: warp w <-> candidates 8w..8w+7 of the tile,
// lane = (cand % 8) * 4 + (j % 4): exactly the B-fragment order of DMMA.8x8x4, so each warp
// store is one contiguous 256-byte line of the packed panel.
template <int DIMP>
__global__ void __launch_bounds__(256) panel_kernel(PanelArgs p) {
    __shared__ __align__(128) double ps_raw[kAcqBN * BOGP_MAX_DIM];   // candidate block, [cand][dim] as in HBM
    __shared__ double xs[DIMP][kAcqBM + 1];
    __shared__ double al[kAcqBM];
    __shared__ double sl[BOGP_MAX_DIM];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ct = blockIdx.x, jb = blockIdx.y;
    const int64_t cbase = p.c0 + (int64_t)ct * kAcqBN;
    const int dim = p.dim;

    const bool explicit_mode = p.cand.points != nullptr;
    // number of valid candidates in this tile (>= 1 by construction of the grid)
    const int64_t remain = p.c_end - cbase;
    const int nvalid = remain >= kAcqBN ? kAcqBN : (int)remain;
    bool used_tma = false;
    if (explicit_mode) {
        const double* src = p.cand.points + cbase * dim;
        const uint32_t bytes = (uint32_t)nvalid * dim * 8;
        // bulk-TMA staging of the candidate block (16-byte granularity and alignment required)
        if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            used_tma = true;
            if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
            __syncthreads();
            if (tid == 0) { mbar_expect_tx(&bar, bytes); bulk_g2s(ps_raw, src, bytes, &bar); }
        }
    }
    // measured points of this row block, transposed; alpha; 1/ell^2
    for (int i = tid; i < kAcqBM * DIMP; i += 256) {
        int r = i / DIMP, k = i % DIMP;
        xs[k][r] = k < dim ? p.x_pad[((int64_t)jb * kAcqBM + r) * dim + k] : 0.0;
    }
    al[tid] = p.alpha[jb * kAcqBM + tid];
    if (tid < BOGP_MAX_DIM) sl[tid] = tid < dim ? p.inv_ell2[tid] : 0.0;

    if (explicit_mode) {
        if (used_tma) {
            mbar_wait(&bar, 0);
        } else {
            for (int i = tid; i < nvalid * dim; i += 256) ps_raw[i] = p.cand.points[cbase * dim + i];
        }
    } else if (tid < nvalid) {
        int64_t f = cbase + tid;     // mixed-radix digits of the flat index, axis 0 slowest (select_parameters.py:273-279)
        for (int k = dim - 1; k >= 0; k--) {
            const int64_t q = f / p.cand.len[k];
            const int dgt = (int)(f - q * p.cand.len[k]);
            ps_raw[tid * dim + k] = p.cand.axes[p.cand.off[k] + dgt];
            f = q;
        }
    }
    __syncthreads();

    const int nl = warp * 8 + (lane >> 2), jj = lane & 3;
    const int64_t cglob = cbase + nl;
    const int ncl = nl < nvalid ? nl : nvalid - 1;     // tail tiles recompute the last valid candidate; masked later
    double pc[DIMP], il[DIMP];      // padded dimensions carry zeros and add exactly +0
#pragma unroll
    for (int k = 0; k < DIMP; k++) { pc[k] = k < dim ? ps_raw[ncl * dim + k] : 0.0; il[k] = sl[k]; }

    double* tile0 = p.panel + ((int64_t)ct * (p.n_pad / kAcqKB) + (int64_t)jb * (kAcqBM / kAcqKB)) * (kAcqKB * kAcqBN);
    double mu = 0.0;
    for (int t = 0; t < kAcqBM / 4; t++) {
        const int jl = t * 4 + jj;
        const int j = jb * kAcqBM + jl;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DIMP; k++) { const double df = pc[k] - xs[k][jl]; s += (df * df) * il[k]; }
        double v = (j < p.n) ? exp_nonpos(-0.5 * s) : 0.0;
        if (p.cand.cross_jitter != 0.0 && (int64_t)j == cglob) v += p.cand.cross_jitter;
        mu += al[jl] * v;
        // tile kt = jb*16 + t/4, kk = t%4, n8 = warp
        tile0[(int64_t)(t >> 2) * (kAcqKB * kAcqBN) + ((t & 3) * 8 + warp) * 32 + lane] = v;
    }
    mu += __shfl_xor_sync(0xffffffffu, mu, 1);
    mu += __shfl_xor_sync(0xffffffffu, mu, 2);
    if (jj == 0) p.mupart[(int64_t)jb * p.S + (int64_t)ct * kAcqBN + nl] = mu;
}

// ------------------------------------------------------------------------------------------------
struct TriArgs {
    const double* wp; const double* panel; double* qpart;
    int nI, nct, n_pad; int64_t S;
};

constexpr int kATileBytes = kAcqBM * kAcqKB * 8;   // 32 KB
constexpr int kBTileBytes = kAcqKB * kAcqBN * 8;   //  8 KB
constexpr int kStageBytes = kATileBytes + kBTileBytes;
constexpr size_t kTriSmem = (size_t)kAcqStages * kStageBytes + 2 * kAcqStages * 8 + 4 * kAcqBN * 8 + 64;

// CTA = (row block ib of W, candidate tile ct); 8 consumer warps (4 x 2, warp tile 64 x 32)
// + 1 producer warp issuing bulk-TMA copies.  Heaviest row blocks are scheduled first.
__global__ void __launch_bounds__(288, 1) trigemm_kernel(TriArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double*   sA   = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kAcqStages * kStageBytes);
    uint64_t* empt = full + kAcqStages;
    double*   red  = reinterpret_cast<double*>(empt + kAcqStages);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ib = g.nI - 1 - (int)(blockIdx.x / g.nct);
    const int ct = (int)(blockIdx.x % g.nct);
    const int nk = (ib + 1) * (kAcqBM / kAcqKB);

    if (tid == 0) {
        for (int s = 0; s < kAcqStages; s++) { mbar_init(&full[s], 1); mbar_init(&empt[s], 8); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 8) {
        if (lane == 0) {
            const double* wsrc = g.wp + (int64_t)ib * (ib + 1) / 2 * (kAcqBM / kAcqKB) * (kAcqBM * kAcqKB);
            const double* psrc = g.panel + (int64_t)ct * (g.n_pad / kAcqKB) * (kAcqKB * kAcqBN);
            for (int kt = 0; kt < nk; kt++) {
                const int s = kt % kAcqStages;
                if (kt >= kAcqStages) mbar_wait(&empt[s], ((kt / kAcqStages) - 1) & 1);
                unsigned char* dst = smem_raw + (size_t)s * kStageBytes;
                mbar_expect_tx(&full[s], kStageBytes);
                bulk_g2s(dst, wsrc + (int64_t)kt * (kAcqBM * kAcqKB), kATileBytes, &full[s]);
                bulk_g2s(dst + kATileBytes, psrc + (int64_t)kt * (kAcqKB * kAcqBN), kBTileBytes, &full[s]);
            }
        }
        return;
    }

    const int wm = warp >> 1, wn = warp & 1;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nk; kt++) {
        const int s = kt % kAcqStages;
        mbar_wait(&full[s], (kt / kAcqStages) & 1);
        const double* a = reinterpret_cast<const double*>(smem_raw + (size_t)s * kStageBytes);
        const double* b = a + kAcqBM * kAcqKB;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = a[((kk * 32 + wm * 8 + i) << 5) + lane];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = b[((kk * 8 + wn * 4 + j) << 5) + lane];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empt[s]);
    }

    // fused epilogue: column sums of squares over the 256 rows of this block
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 8; i++) s += acc[i][j][e] * acc[i][j][e];
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            if (lane < 4) red[wm * kAcqBN + wn * 32 + j * 8 + 2 * lane + e] = s;
        }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid < kAcqBN) {
        const double q = ((red[tid] + red[kAcqBN + tid]) + red[2 * kAcqBN + tid]) + red[3 * kAcqBN + tid];
        g.qpart[(int64_t)ib * g.S + (int64_t)ct * kAcqBN + tid] = q;
    }
}

// ------------------------------------------------------------------------------------------------
struct FinalArgs {
    const double* qpart; const double* mupart;
    double* mu_out; double* sigma_out; double* acq_out;   // already offset to this chunk (or null)
    double* block_score; long long* block_index; int* nan_flag;
    int64_t c0, cur, S; int nIq, nImu; int kind; double explore, f_best, prior;
    const int* d_count; const long long* idx_map;   // screened sweeps: compacted survivors (count on the device, global flat index per slot)
};

__device__ __forceinline__ void block_argmax(double s, long long i, double* block_score, long long* block_index, int slot) {
    __shared__ double ws[8]; __shared__ long long wi[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double os = __shfl_xor_sync(0xffffffffu, s, o);
        long long oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (better(os, oi, s, i)) { s = os; i = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ws[warp] = s; wi[warp] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) if (better(ws[w], wi[w], s, i)) { s = ws[w]; i = wi[w]; }
        block_score[slot] = s; block_index[slot] = i;
    }
}

constexpr long long kNoIndex = 0x7fffffffffffffffLL;
constexpr int kScreenSeed = 4096;          // candidates of the strided seed sample of a screened sweep

__global__ void __launch_bounds__(256) finalize_kernel(FinalArgs f) {
    const int64_t cl = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double score = -INFINITY; long long idx = kNoIndex;
    int64_t cur = f.cur;
    if (f.d_count) { const int64_t dc = (int64_t)*f.d_count - f.c0; cur = dc < cur ? dc : cur; }
    if (cl < cur) {
        double q = 0.0, mu = 0.0;
        for (int b = 0; b < f.nIq; b++) q += f.qpart[(int64_t)b * f.S + cl];
        for (int b = 0; b < f.nImu; b++) mu += f.mupart[(int64_t)b * f.S + cl];
        const double var = f.prior - q;
        const double sigma = sqrt(fabs(var));                   // np.sqrt(np.abs(.)), point_selector.py:98
        score = acquisition_value(f.kind, mu, sigma, f.explore, f.f_best);
        idx = f.idx_map ? f.idx_map[f.c0 + cl] : f.c0 + cl;
        if (f.mu_out) f.mu_out[cl] = mu;
        if (f.sigma_out) f.sigma_out[cl] = sigma;
        if (f.acq_out) f.acq_out[cl] = score;
        if (score != score) { atomicExch(f.nan_flag, 1); score = -INFINITY; }
    }
    block_argmax(score, idx, f.block_score, f.block_index, blockIdx.x);
}

// acquisition + arg-max on mu/sigma already on the device (lower_confidence_bound, :197-207)
__global__ void __launch_bounds__(256) score_kernel(const double* __restrict__ mu, const double* __restrict__ sigma, int64_t c,
                                                    int64_t index_offset, int kind, double explore, double f_best, double* acq_out,
                                                    double* block_score, long long* block_index, int* nan_flag) {
    double best = -INFINITY; long long bi = kNoIndex;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < c; i += (int64_t)gridDim.x * 256) {
        double s = acquisition_value(kind, mu[i], sigma[i], explore, f_best);
        if (acq_out) acq_out[i] = s;
        if (s != s) { atomicExch(nan_flag, 1); continue; }
        if (better(s, i + index_offset, best, bi)) { best = s; bi = i + index_offset; }
    }
    block_argmax(best, bi, block_score, block_index, blockIdx.x);
}

// best[0] (score), besti[0] (index): running winner, folded with nblocks block winners.
__global__ void __launch_bounds__(256) merge_kernel(const double* block_score, const long long* block_index, int nblocks,
                                                    double* best, long long* besti) {
    __shared__ double ws[8]; __shared__ long long wi[8];
    double s = -INFINITY; long long i = kNoIndex;
    for (int b = threadIdx.x; b < nblocks; b += 256) if (better(block_score[b], block_index[b], s, i)) { s = block_score[b]; i = block_index[b]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double os = __shfl_xor_sync(0xffffffffu, s, o);
        long long oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (better(os, oi, s, i)) { s = os; i = oi; }
    }
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wi[threadIdx.x >> 5] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) if (better(ws[w], wi[w], s, i)) { s = ws[w]; i = wi[w]; }
        if (better(s, i, best[0], besti[0])) { best[0] = s; besti[0] = i; }
    }
}

// ------------------------------------------------------------------------------------------------
// Arg-max-only sweeps: screen, then score exactly.
//
// Both acquisitions are non-decreasing in sigma (LCB = explore*sigma - mu with explore >= 0; dEI/dsigma = phi(z) > 0),
// and sigma^2 = prior - |L^-1 k_*|^2 <= prior.  So  U(c) = A(mu_c, sqrt(prior))  is an upper bound of candidate c's
// score that needs only the posterior MEAN (N flops, the mu-only pass of the panel kernel) and not the N^2-flop
// product.  A candidate with U(c) < best -- `best` being the EXACT score of some candidate already scored -- can
// neither be the maximum nor tie with it and is dropped; everything else (NaNs included: the comparison is false) is
// compacted and goes through the exact kernels.  The winner (score, index) is therefore exactly the one of the full
// sweep.  LCB: the bound is computed with the very operations of the exact score (fl is monotone), no slack.  EI: the
// computed value of the smooth formula can deviate from monotonicity by rounding, so 1e-12 (|f_best - mu| + sigma_max)
// -- four orders of magnitude above the rounding error of the two products -- is added to the bound.
// ------------------------------------------------------------------------------------------------
struct ScreenArgs {
    const double* mupart; int nImu; int64_t S;
    CandDesc cand; int dim;
    int64_t c0, cur, stride;      // flat indices c0 + i * stride, i < cur
    int seed;                     // 1: keep every candidate without testing (the strided seed sample)
    int kind; double explore, f_best, sigma_max;
    const bogp_result* best;
    long long* surv_idx; double* surv_pts; int* count; int capacity;
    unsigned long long* stats;    // [0] candidates screened, [1] survivors (bogp_screen_stats)
};

__global__ void __launch_bounds__(256) screen_kernel(ScreenArgs a) {
    const int64_t cl = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false;
    long long idx = 0;
    if (cl < a.cur) {
        idx = a.seed ? seed_index((unsigned long long)cl, a.c0, a.stride) : a.c0 + cl * a.stride;      // seed: stride holds the length of the range
        keep = true;
        if (!a.seed) {
            double mu = 0.0;
            for (int b = 0; b < a.nImu; b++) mu += a.mupart[(int64_t)b * a.S + cl];      // same order as finalize_kernel
            double bound = acquisition_value(a.kind, mu, a.sigma_max, a.explore, a.f_best);
            if (a.kind == BOGP_ACQ_EI) bound += 1e-12 * (fabs(a.f_best - mu) + a.sigma_max);
            keep = !(bound < a.best->score);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!a.seed) {
        const unsigned inrange = __ballot_sync(0xffffffffu, cl < a.cur);
        if (lane == 0 && inrange) { atomicAdd(&a.stats[0], (unsigned long long)__popc(inrange)); if (m) atomicAdd(&a.stats[1], (unsigned long long)__popc(m)); }
    }
    if (m == 0u) return;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(a.count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (!keep) return;
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (pos >= a.capacity) return;                      // cannot happen (capacity = batch size); guards the buffers
    a.surv_idx[pos] = idx;
    double* dst = a.surv_pts + (int64_t)pos * a.dim;
    if (a.cand.points) {
        for (int k = 0; k < a.dim; k++) dst[k] = a.cand.points[idx * a.dim + k];
    } else {
        long long f = idx;
        for (int k = a.dim - 1; k >= 0; k--) {
            const long long q = f / a.cand.len[k];
            dst[k] = a.cand.axes[a.cand.off[k] + (int)(f - q * a.cand.len[k])];
            f = q;
        }
    }
}

__global__ void init_best_kernel(double* best, long long* besti, int* nan_flag) {
    best[0] = -INFINITY; besti[0] = kNoIndex; nan_flag[0] = 0;
}

// bogp_result records (24 bytes: score, index, nan flag) gathered from several sweeps / ranks -> one record
__global__ void __launch_bounds__(256) reduce_results_kernel(const bogp_result* __restrict__ in, int count, bogp_result* __restrict__ out) {
    __shared__ double ws[8]; __shared__ long long wi[8]; __shared__ int wn;
    if (threadIdx.x == 0) wn = 0;
    __syncthreads();
    double s = -INFINITY; long long i = kNoIndex; int nanf = 0;
    for (int b = threadIdx.x; b < count; b += 256) {
        const double bs = in[b].score; const long long bi = in[b].index;
        nanf |= in[b].nan_flag;
        if (better(bs, bi, s, i)) { s = bs; i = bi; }
    }
    if (nanf) atomicOr(&wn, 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double os = __shfl_xor_sync(0xffffffffu, s, o);
        long long oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (better(os, oi, s, i)) { s = os; i = oi; }
    }
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wi[threadIdx.x >> 5] = i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) if (better(ws[w], wi[w], s, i)) { s = ws[w]; i = wi[w]; }
        out->score = s; out->index = i; out->nan_flag = wn; out->reserved = 0;
    }
}

// Workspace of one chunk of S candidates.  The layout is sized for the larger of the two tensor
// paths (FP64: 8 B per panel entry, 256-row blocks; INT8: 7 B per entry, 128-row blocks) so that a
// workspace is valid whichever path is selected.
struct AcqLayout { size_t panel, qpart, mupart, total; int64_t S; };
static size_t acq_bytes_per_candidate(int64_t n_pad) {
    return (size_t)n_pad * 8 + (size_t)(n_pad / 128) * 8 + (size_t)(n_pad / kAcqBM) * 8;
}
static AcqLayout acq_layout(int64_t n_pad, int64_t max_chunk) {
    AcqLayout l{}; size_t off = 0;
    int64_t S = (max_chunk + kAcqBN - 1) / kAcqBN * kAcqBN;
    if (S < kAcqBN) S = kAcqBN;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    l.panel = take((size_t)n_pad * S * 8); l.qpart = take((size_t)(n_pad / 128) * S * 8); l.mupart = take((size_t)(n_pad / kAcqBM) * S * 8);
    l.total = off; l.S = S;
    return l;
}

}  // namespace bogp

using namespace bogp;

extern "C" size_t bogp_acquire_workspace_bytes(const bogp_fit* fit, int64_t max_chunk) {
    if (!fit || max_chunk <= 0) return 0;
    const int64_t n_pad = bogp_fit_n_pad(fit);
    const size_t chunked = acq_layout(n_pad, max_chunk).total;
    const size_t fused = n_pad <= 16384 ? fused_workspace_bytes(n_pad) : 0;      // ring of the fused sweep kernel (INT8 path)
    return (chunked > fused ? chunked : fused) + grid_table_reserve(n_pad);     // + the kernel-factor tables of a grid sweep
}

// The whole sweep as one launch of the fused persistent kernel (acquire_fused.cu).  Returns false if the workspace cannot
// hold its ring (the caller then runs the separate kernels: same results); `rc` carries a launch error.
static bool fused_sweep(bogp_ctx* ctx, const bogp_fit* fit, const CandDesc& cd, int64_t c_begin, int64_t c_end, int dim, int64_t n_pad,
                        int kind, double explore, double f_best, double prior_diag, double* d_mu_out, double* d_sigma_out,
                        double* d_acq_out, void* d_workspace, size_t workspace_bytes, bogp_result* d_result, cudaStream_t st,
                        const AcqChunk& tab, int& rc) {
    AcqChunk a{};
    a.ft = tab.ft; a.tt = tab.tt;
    for (int k = 0; k < BOGP_MAX_DIM; k++) { a.toff[k] = tab.toff[k]; a.lenp[k] = tab.lenp[k]; }
    a.points = cd.points; a.axes = cd.axes; a.cross_jitter = cd.cross_jitter;
    for (int k = 0; k < BOGP_MAX_DIM; k++) { a.len[k] = cd.len[k]; a.off[k] = cd.off[k]; }
    a.x_pad = fit_xpad(fit); a.inv_ell2 = fit_inv_ell2(fit); a.alpha = fit_alpha(fit);
    a.wp = fit_wp(fit); a.wq = fit_wq(fit); a.wscale = fit_wscale(fit);
    a.c0 = c_begin; a.c_end = c_end; a.cur = c_end - c_begin; a.S = 0;
    a.n = (int)fit_n(fit); a.n_pad = (int)n_pad; a.dim = dim;
    FusedFinal f{d_mu_out, d_sigma_out, d_acq_out, nullptr, kind, explore, f_best, prior_diag, d_result};
    rc = launch_acquire_fused(ctx, a, f, d_workspace, workspace_bytes, st);
    if (rc == 1) { rc = BOGP_OK; return false; }
    return true;
}

// All device work of one sweep, enqueued on ctx->stream without any host synchronisation.  The running winner lives in
// the caller's 24-byte device record `d_result` (score, index, nan flag).
static int acquire_enqueue(bogp_ctx* ctx, const bogp_fit* fit, const bogp_candidates* cand, int64_t c_begin, int64_t c_end,
                           int kind, double explore, double f_best, double prior_diag, double* d_mu_out,
                           double* d_sigma_out, double* d_acq_out, void* d_workspace, size_t workspace_bytes,
                           bogp_result* d_result) {
    if (!ctx || !fit || !cand || !d_workspace || !d_result || c_begin < 0 || c_end <= c_begin || c_end > cand->c_total ||
        (kind != BOGP_ACQ_LCB && kind != BOGP_ACQ_EI)) {
        set_error("bogp_acquire: bad argument"); return BOGP_ERR_BAD_ARG;
    }
    NvtxRange nvtx("bogp acquisition sweep");
    const int dim = fit_dim(fit);
    const int64_t n_pad = bogp_fit_n_pad(fit);
    const int nI = (int)(n_pad / kAcqBM);
    CandDesc cd{};
    cd.points = cand->d_points; cd.axes = cand->d_axes; cd.c_total = cand->c_total; cd.cross_jitter = cand->cross_jitter;
    if (!cd.points) {
        if (!cd.axes || !cand->h_axis_len) { set_error("bogp_acquire: neither points nor grid axes given"); return BOGP_ERR_BAD_ARG; }
        int64_t prod = 1; int off = 0;
        for (int k = 0; k < dim; k++) {
            cd.len[k] = cand->h_axis_len[k]; cd.off[k] = off; off += cd.len[k];
            if (cd.len[k] <= 0) { set_error("bogp_acquire: empty grid axis %d", k); return BOGP_ERR_BAD_ARG; }
            prod *= cd.len[k];
        }
        if (prod != cand->c_total) { set_error("bogp_acquire: grid has %lld points, c_total says %lld", (long long)prod, (long long)cand->c_total); return BOGP_ERR_BAD_ARG; }
    }
    // the int32 level sums of the digit-slice product are overflow-free up to K = 16384 (7 * K * 2^14 < 2^31);
    // larger systems take the fp64 path
    const bool use_i8 = ctx->acquire_path == BOGP_PATH_INT8_TCGEN05 && n_pad <= 16384;
    // Grid sweeps on the INT8 path: per-axis kernel-factor tables, built once per sweep into the tail of the workspace
    AcqChunk tab{};
    if (use_i8 && !cd.points) {
        tab.axes = cd.axes; tab.x_pad = fit_xpad(fit); tab.inv_ell2 = fit_inv_ell2(fit); tab.dim = dim; tab.n_pad = (int)n_pad;
        for (int k = 0; k < BOGP_MAX_DIM; k++) { tab.len[k] = cd.len[k]; tab.off[k] = cd.off[k]; }
        tab.n = (int)fit_n(fit);
        const size_t tb = (grid_table_geometry(tab) + 255) / 256 * 256;
        if (tb > 0 && tb <= grid_table_reserve(n_pad) && workspace_bytes >= tb + ((size_t)64 << 10) && workspace_bytes - tb >= acq_layout(n_pad, kAcqBN).total) {
            workspace_bytes -= tb;
            tab.ft = reinterpret_cast<const double*>(static_cast<char*>(d_workspace) + workspace_bytes);
            const int trc = launch_grid_factors(ctx, tab, const_cast<double*>(tab.ft), ctx->stream);
            if (trc) return trc;
        }
    }
    // chunk capacity from the workspace size
    const size_t per_cand = acq_bytes_per_candidate(n_pad);
    int64_t S = (int64_t)(workspace_bytes / per_cand) / kAcqBN * kAcqBN;
    if (S < kAcqBN) { set_error("bogp_acquire: workspace too small"); return BOGP_ERR_WORKSPACE; }
    if (S > (int64_t)kMaxReduceBlocks * 256) S = (int64_t)kMaxReduceBlocks * 256;
    if (S > c_end - c_begin) S = (c_end - c_begin + kAcqBN - 1) / kAcqBN * kAcqBN;
    const AcqLayout l = acq_layout(n_pad, S);
    if (l.total > workspace_bytes) { set_error("bogp_acquire: workspace too small"); return BOGP_ERR_WORKSPACE; }
    char* base = static_cast<char*>(d_workspace);
    double* panel = (double*)(base + l.panel); double* qpart = (double*)(base + l.qpart); double* mupart = (double*)(base + l.mupart);

    static DeviceOnce configured;
    if (configured.need(ctx->device)) {
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(trigemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTriSmem));
    }
    cudaStream_t st = ctx->stream;
    double* best = &d_result->score; long long* besti = reinterpret_cast<long long*>(&d_result->index);
    int* nan_flag = &d_result->nan_flag;
    {   // the fit packs W for the path selected at fit time; a later switch packs on first use
        const int prc = fit_ensure_packed(ctx, fit, use_i8 ? BOGP_PATH_INT8_TCGEN05 : BOGP_PATH_FP64_DMMA);
        if (prc) return prc;
    }
    auto finish_chunk = [&](int64_t c0, int64_t cur, int64_t Sb, const double* qp, const double* mp, int nIq,
                            const int* d_count = nullptr, const long long* idx_map = nullptr) -> int {
        const int nfb = (int)((cur + 255) / 256);
        if (nfb > kMaxReduceBlocks) { set_error("bogp_acquire: chunk of %lld candidates exceeds the reducer capacity", (long long)cur); return BOGP_ERR_BAD_ARG; }
        const int64_t o = c0 - c_begin;
        FinalArgs fa{qp, mp, d_mu_out ? d_mu_out + o : nullptr, d_sigma_out ? d_sigma_out + o : nullptr,
                     d_acq_out ? d_acq_out + o : nullptr, ctx->d_block_score, ctx->d_block_index, nan_flag,
                     c0, cur, Sb, nIq, nI, kind, explore, f_best, prior_diag, d_count, idx_map};
        BOGP_PROFILED(ctx, BOGP_PROF_FINALIZE, (finalize_kernel<<<nfb, 256, 0, st>>>(fa))); BOGP_LAUNCH_CHECK(ctx);
        BOGP_PROFILED(ctx, BOGP_PROF_MERGE, (merge_kernel<<<1, 256, 0, st>>>(ctx->d_block_score, ctx->d_block_index, nfb, best, besti))); BOGP_LAUNCH_CHECK(ctx);
        return BOGP_OK;
    };
    int fused_rc = BOGP_OK;
    const bool want_out = d_mu_out || d_sigma_out || d_acq_out;
    const bool screen = use_i8 && ctx->screening && !want_out && !ctx->profile && cand->cross_jitter == 0.0 && prior_diag >= 1.0 &&
                        (kind == BOGP_ACQ_EI || explore >= 0.0) && (c_end - c_begin) >= 4 * kScreenSeed && S >= kScreenSeed;
    bool fused_done = false;
    if (!screen && use_i8 && ctx->fused) {
        fused_done = fused_sweep(ctx, fit, cd, c_begin, c_end, dim, n_pad, kind, explore, f_best, prior_diag,
                                 d_mu_out, d_sigma_out, d_acq_out, d_workspace, workspace_bytes, d_result, st, tab, fused_rc);
        if (fused_rc) return fused_rc;
    }
    if (fused_done) return BOGP_OK;
    if (screen && tab.ft) {
        // grid sweep: the means of all candidates from GEMMs, survivors scored exactly (screen_gemm.cu)
        const int grc = gemm_screen_sweep(ctx, fit, tab, c_begin, c_end, kind, explore, f_best, prior_diag, d_workspace, workspace_bytes, d_result);
        if (grc != 1) return grc;
    }
    init_best_kernel<<<1, 1, 0, st>>>(best, besti, nan_flag); BOGP_LAUNCH_CHECK(ctx);
    if (screen) {
        // One buffer set of S candidates.  The tail of the panel region (the digits take 7 of its 8 bytes per entry)
        // holds the compacted survivors: coordinates, global flat indices and the device-side count.
        const AcqLayout lb = acq_layout(n_pad, S);
        char* tail = base + lb.panel + (size_t)n_pad * S * 7;
        double* surv_pts = reinterpret_cast<double*>(tail);
        long long* surv_idx = reinterpret_cast<long long*>(tail + (size_t)S * BOGP_MAX_DIM * 8);
        int* count = reinterpret_cast<int*>(tail + (size_t)S * (BOGP_MAX_DIM + 1) * 8);
        const int64_t total = c_end - c_begin;
        auto make_chunk = [&](int64_t c0, int64_t cur_cap, bool compacted) {
            AcqChunk a{};
            // compacted survivors of a grid sweep with tables: scored from their flat indices through the same tables (same bits
            // as inside the contiguous sweep); otherwise from their gathered coordinates
            const bool by_index = compacted && tab.ft != nullptr;
            a.points = (compacted && !by_index) ? surv_pts : cd.points; a.axes = (compacted && !by_index) ? nullptr : cd.axes; a.cross_jitter = 0.0;
            a.idx_list = by_index ? surv_idx : nullptr;
            for (int k = 0; k < BOGP_MAX_DIM; k++) { a.len[k] = cd.len[k]; a.off[k] = cd.off[k]; }
            a.x_pad = fit_xpad(fit); a.inv_ell2 = fit_inv_ell2(fit); a.alpha = fit_alpha(fit);
            a.wp = fit_wp(fit); a.wq = fit_wq(fit); a.wscale = fit_wscale(fit);
            a.panel = base + lb.panel; a.qpart = (double*)(base + lb.qpart); a.mupart = (double*)(base + lb.mupart);
            a.c0 = compacted ? 0 : c0; a.c_end = compacted ? cur_cap : c_end; a.cur = cur_cap; a.S = S;
            a.n = (int)fit_n(fit); a.n_pad = (int)n_pad; a.dim = dim;
            a.d_count = compacted ? count : nullptr;
            if (!compacted || by_index) { a.ft = tab.ft; a.tt = tab.tt; for (int k = 0; k < BOGP_MAX_DIM; k++) { a.toff[k] = tab.toff[k]; a.lenp[k] = tab.lenp[k]; } }
            return a;
        };
        auto exact_pass = [&](int64_t cap) -> int {          // the survivors in surv_pts[0 .. *count), at most `cap`
            const AcqChunk a = make_chunk(0, cap, true);
            int rc = launch_panel_i8(ctx, a, st); if (rc) return rc;
            rc = launch_trigemm_i8(ctx, a, st); if (rc) return rc;
            return finish_chunk(0, cap, S, a.qpart, a.mupart, (int)(n_pad / 128), count, surv_idx);
        };
        ScreenArgs sa{};
        sa.mupart = (double*)(base + lb.mupart); sa.nImu = nI; sa.S = S; sa.cand = cd; sa.dim = dim;
        sa.kind = kind; sa.explore = explore; sa.f_best = f_best; sa.sigma_max = sqrt(prior_diag);
        sa.best = d_result; sa.surv_idx = surv_idx; sa.surv_pts = surv_pts; sa.count = count; sa.capacity = (int)S;
        sa.stats = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 16);
        // seed: a strided sample over the whole range is scored first, so that the running best is already high
        // when the first batch is screened
        BOGP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
        sa.seed = 1; sa.c0 = ctx->global_seed ? 0 : c_begin; sa.cur = kScreenSeed;
        sa.stride = ctx->global_seed ? cd.c_total : total;       // length of the sampled range (global seed: one shard of a sharded arg-max, bogp_set_global_seed)
        screen_kernel<<<(unsigned)((sa.cur + 255) / 256), 256, 0, st>>>(sa); BOGP_LAUNCH_CHECK(ctx);
        int rc = exact_pass(kScreenSeed); if (rc) return rc;
        sa.seed = 0; sa.stride = 1;
        for (int64_t c0 = c_begin; c0 < c_end; c0 += S) {
            const int64_t cur = (c_end - c0 < S) ? (c_end - c0) : S;
            rc = launch_panel_i8(ctx, make_chunk(c0, cur, false), st, /*mu_only=*/true); if (rc) return rc;
            BOGP_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
            sa.c0 = c0; sa.cur = cur;
            screen_kernel<<<(unsigned)((cur + 255) / 256), 256, 0, st>>>(sa); BOGP_LAUNCH_CHECK(ctx);
            rc = exact_pass((cur + kAcqBN - 1) / kAcqBN * kAcqBN); if (rc) return rc;
        }
    } else if (use_i8) {
        // INT8 path, separate kernels.  The panel kernel (FP64/INT pipes) and the tensor-core kernel use different
        // pipes and fit on one SM together, so with two buffer sets the panel of chunk s+1 is built
        // on a second stream while chunk s is on the tensor cores.
        const int64_t total = c_end - c_begin;
        const bool overlap = !ctx->profile && total > S / 2 && S >= 4 * kAcqBN;
        const int nbuf = overlap ? 2 : 1;
        const int64_t Sb = overlap ? (S / 2) / kAcqBN * kAcqBN : S;
        const AcqLayout lb = acq_layout(n_pad, Sb);
        if ((size_t)nbuf * lb.total > workspace_bytes) { set_error("bogp_acquire: workspace too small"); return BOGP_ERR_WORKSPACE; }
        auto make_chunk = [&](int64_t c0, int b) {
            AcqChunk a{};
            char* bb = base + (size_t)b * lb.total;
            a.points = cd.points; a.axes = cd.axes; a.cross_jitter = cd.cross_jitter;
            for (int k = 0; k < BOGP_MAX_DIM; k++) { a.len[k] = cd.len[k]; a.off[k] = cd.off[k]; }
            a.x_pad = fit_xpad(fit); a.inv_ell2 = fit_inv_ell2(fit); a.alpha = fit_alpha(fit);
            a.wp = fit_wp(fit); a.wq = fit_wq(fit); a.wscale = fit_wscale(fit);
            a.panel = bb + lb.panel; a.qpart = (double*)(bb + lb.qpart); a.mupart = (double*)(bb + lb.mupart);
            a.c0 = c0; a.c_end = c_end; a.cur = (c_end - c0 < Sb) ? (c_end - c0) : Sb; a.S = Sb;
            a.n = (int)fit_n(fit); a.n_pad = (int)n_pad; a.dim = dim;
            a.ft = tab.ft; a.tt = tab.tt; for (int k = 0; k < BOGP_MAX_DIM; k++) { a.toff[k] = tab.toff[k]; a.lenp[k] = tab.lenp[k]; }
            return a;
        };
        cudaStream_t ps = overlap ? ctx->aux_stream : st;
        if (overlap) {
            BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, st));
            BOGP_CUDA_CHECK(cudaStreamWaitEvent(ps, ctx->ev_fork, 0));
        }
        const int64_t nchunks = (total + Sb - 1) / Sb;
        int rc = launch_panel_i8(ctx, make_chunk(c_begin, 0), ps);
        if (rc) return rc;
        if (overlap) BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_panel[0], ps));
        for (int64_t s = 0; s < nchunks; s++) {
            const int b = (int)(s % nbuf);
            if (overlap && s + 1 < nchunks) {          // next panel into the other buffer set
                const int nb = (int)((s + 1) % 2);
                if (s >= 1) BOGP_CUDA_CHECK(cudaStreamWaitEvent(ps, ctx->ev_done[nb], 0));
                rc = launch_panel_i8(ctx, make_chunk(c_begin + (s + 1) * Sb, nb), ps);
                if (rc) return rc;
                BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_panel[nb], ps));
            }
            const AcqChunk a = make_chunk(c_begin + s * Sb, b);
            if (overlap) BOGP_CUDA_CHECK(cudaStreamWaitEvent(st, ctx->ev_panel[b], 0));
            rc = launch_trigemm_i8(ctx, a, st);
            if (rc) return rc;
            rc = finish_chunk(a.c0, a.cur, Sb, a.qpart, a.mupart, (int)(n_pad / 128));
            if (rc) return rc;
            if (overlap) BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_done[b], st));
            if (!overlap && s + 1 < nchunks) {
                rc = launch_panel_i8(ctx, make_chunk(c_begin + (s + 1) * Sb, 0), ps);
                if (rc) return rc;
            }
        }
    } else {
        for (int64_t c0 = c_begin; c0 < c_end; c0 += S) {
            const int64_t cur = (c_end - c0 < S) ? (c_end - c0) : S;
            const int nct = (int)((cur + kAcqBN - 1) / kAcqBN);
            PanelArgs pa{cd, fit_xpad(fit), fit_inv_ell2(fit), fit_alpha(fit), panel, mupart, c0, c_end, S, (int)fit_n(fit), (int)n_pad, dim};
#define BOGP_PANEL(D) BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_kernel<D><<<dim3(nct, nI), 256, 0, st>>>(pa)))
            if (dim <= 2) BOGP_PANEL(2); else if (dim <= 4) BOGP_PANEL(4); else if (dim <= 6) BOGP_PANEL(6);
            else if (dim <= 8) BOGP_PANEL(8); else if (dim <= 10) BOGP_PANEL(10); else if (dim <= 12) BOGP_PANEL(12);
            else BOGP_PANEL(16);
#undef BOGP_PANEL
            BOGP_LAUNCH_CHECK(ctx);
            TriArgs ta{fit_wp(fit), panel, qpart, nI, nct, (int)n_pad, S};
            BOGP_PROFILED(ctx, BOGP_PROF_TRIGEMM, (trigemm_kernel<<<nI * nct, 288, kTriSmem, st>>>(ta))); BOGP_LAUNCH_CHECK(ctx);
            int rc = finish_chunk(c0, cur, S, qpart, mupart, nI);
            if (rc) return rc;
        }
    }
    return BOGP_OK;
}

static int read_result(bogp_ctx* ctx, const bogp_result* d_result, const char* who, double* h_best_score, int64_t* h_best_index) {
    bogp_result h;
    BOGP_CUDA_CHECK(cudaMemcpyAsync(&h, d_result, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BOGP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h.nan_flag) { set_error("%s: NaN acquisition value (reference raises IndexError, point_selector.py:207)", who); return BOGP_ERR_NAN_SCORE; }
    if (h_best_score) *h_best_score = h.score;
    if (h_best_index) *h_best_index = h.index;
    return BOGP_OK;
}

extern "C" int bogp_acquire_async(bogp_ctx* ctx, const bogp_fit* fit, const bogp_candidates* cand, int64_t c_begin, int64_t c_end,
                                  int kind, double explore, double f_best, double prior_diag, double* d_mu_out,
                                  double* d_sigma_out, double* d_acq_out, void* d_workspace, size_t workspace_bytes,
                                  bogp_result* d_result) {
    return acquire_enqueue(ctx, fit, cand, c_begin, c_end, kind, explore, f_best, prior_diag, d_mu_out, d_sigma_out, d_acq_out,
                           d_workspace, workspace_bytes, d_result);
}

extern "C" int bogp_acquire(bogp_ctx* ctx, const bogp_fit* fit, const bogp_candidates* cand, int64_t c_begin, int64_t c_end,
                            int kind, double explore, double f_best, double prior_diag, double* d_mu_out,
                            double* d_sigma_out, double* d_acq_out, void* d_workspace, size_t workspace_bytes,
                            double* h_best_score, int64_t* h_best_index) {
    if (!ctx) { set_error("bogp_acquire: bad argument"); return BOGP_ERR_BAD_ARG; }
    bogp_result* slot = reinterpret_cast<bogp_result*>(ctx->d_scalars);
    const int rc = acquire_enqueue(ctx, fit, cand, c_begin, c_end, kind, explore, f_best, prior_diag, d_mu_out, d_sigma_out, d_acq_out,
                                   d_workspace, workspace_bytes, slot);
    if (rc) return rc;
    if (!h_best_score && !h_best_index) return BOGP_OK;
    return read_result(ctx, slot, "bogp_acquire", h_best_score, h_best_index);
}

extern "C" int bogp_screen_stats(bogp_ctx* ctx, int64_t* h_screened, int64_t* h_survived, int reset) {
    if (!ctx) { set_error("bogp_screen_stats: null context"); return BOGP_ERR_BAD_ARG; }
    unsigned long long h[2] = {0, 0};
    BOGP_CUDA_CHECK(cudaMemcpyAsync(h, ctx->d_scalars + 16, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) BOGP_CUDA_CHECK(cudaMemsetAsync(ctx->d_scalars + 16, 0, sizeof(h), ctx->stream));
    BOGP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h_screened) *h_screened = (int64_t)h[0];
    if (h_survived) *h_survived = (int64_t)h[1];
    return BOGP_OK;
}

extern "C" int bogp_reduce_results(bogp_ctx* ctx, const bogp_result* d_results, int count, bogp_result* d_out,
                                   double* h_best_score, int64_t* h_best_index) {
    if (!ctx || !d_results || count <= 0) { set_error("bogp_reduce_results: bad argument"); return BOGP_ERR_BAD_ARG; }
    bogp_result* out = d_out ? d_out : reinterpret_cast<bogp_result*>(ctx->d_scalars + 8);
    reduce_results_kernel<<<1, 256, 0, ctx->stream>>>(d_results, count, out); BOGP_LAUNCH_CHECK(ctx);
    if (!h_best_score && !h_best_index) return BOGP_OK;
    return read_result(ctx, out, "bogp_reduce_results", h_best_score, h_best_index);
}

extern "C" int bogp_score_argmax_async(bogp_ctx* ctx, const double* d_mu, const double* d_sigma, int64_t c, int64_t index_offset,
                                       int kind, double explore, double f_best, double* d_acq_out, bogp_result* d_result) {
    if (!ctx || !d_mu || !d_sigma || !d_result || c <= 0 || (kind != BOGP_ACQ_LCB && kind != BOGP_ACQ_EI)) { set_error("bogp_score_argmax: bad argument"); return BOGP_ERR_BAD_ARG; }
    cudaStream_t st = ctx->stream;
    double* best = &d_result->score; long long* besti = reinterpret_cast<long long*>(&d_result->index);
    int* nan_flag = &d_result->nan_flag;
    init_best_kernel<<<1, 1, 0, st>>>(best, besti, nan_flag); BOGP_LAUNCH_CHECK(ctx);
    int nb = (int)((c + 255) / 256); if (nb > 2 * ctx->sm_count) nb = 2 * ctx->sm_count;
    score_kernel<<<nb, 256, 0, st>>>(d_mu, d_sigma, c, index_offset, kind, explore, f_best, d_acq_out, ctx->d_block_score, ctx->d_block_index, nan_flag); BOGP_LAUNCH_CHECK(ctx);
    merge_kernel<<<1, 256, 0, st>>>(ctx->d_block_score, ctx->d_block_index, nb, best, besti); BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

extern "C" int bogp_score_argmax(bogp_ctx* ctx, const double* d_mu, const double* d_sigma, int64_t c, int kind, double explore,
                                 double f_best, double* d_acq_out, double* h_best_score, int64_t* h_best_index) {
    if (!ctx) { set_error("bogp_score_argmax: bad argument"); return BOGP_ERR_BAD_ARG; }
    bogp_result* slot = reinterpret_cast<bogp_result*>(ctx->d_scalars);
    const int rc = bogp_score_argmax_async(ctx, d_mu, d_sigma, c, 0, kind, explore, f_best, d_acq_out, slot);
    if (rc) return rc;
    return read_result(ctx, slot, "bogp_score_argmax", h_best_score, h_best_index);
}
