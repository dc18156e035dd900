// GP fit on B200: Gram matrix (K1), blocked fp64 Cholesky with DMMA trailing update (K2),
// recursive triangular inverse W = L^-1, alpha = K^-1 y, log det, and the fragment-packed
// copy of W that the acquisition kernel streams with bulk-TMA copies.
//
// Replaces np.linalg.inv / np.linalg.det / the quadratic form of the reference
// (point_selector.py:79,89-90,117-119).
#include "common.cuh"
#include "gemm_f64.cuh"
#include "fit.cuh"
#include "chol_diag.cuh"
#include <vector>
#include <cstdlib>

namespace bogp {

// ------------------------------------------------------------------------------------------------
// K1: ARD squared-exponential Gram matrix                                point_selector.py:166-195
// 64x64 output tile per CTA, 4x4 elements per thread, points staged transposed in shared memory,
// rows written as double2 pairs (each warp writes 512 contiguous bytes per row pair).
// ------------------------------------------------------------------------------------------------
struct GramArgs {
    const double* a; const double* b; const double* inv_ell2;
    double* k; int64_t ldk; int64_t na, nb; int dim; double jitter;
    int64_t na_valid, nb_valid;   // rows/cols beyond these are identity padding (fit) -- pass na/nb for none
    int lower_tiles_only;
    int64_t strideK; int ell_stride;   // per blockIdx.z (batched multi-restart fit)
};

__global__ void __launch_bounds__(256) gram_kernel(GramArgs g) {
    __shared__ double sa[BOGP_MAX_DIM][64];
    __shared__ double sb[BOGP_MAX_DIM][64];
    __shared__ double sl[BOGP_MAX_DIM];
    const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
    if (g.lower_tiles_only && c0 > r0) return;
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * g.dim; i += 256) {
        int p = i / g.dim, k = i % g.dim;
        sa[k][p] = (r0 + p < g.na_valid) ? g.a[(r0 + p) * g.dim + k] : 0.0;
        sb[k][p] = (c0 + p < g.nb_valid) ? g.b[(c0 + p) * g.dim + k] : 0.0;
    }
    if (tid < g.dim) sl[tid] = g.inv_ell2[(int64_t)blockIdx.z * g.ell_stride + tid];
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    double s[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s[i][j] = 0.0;
    for (int k = 0; k < g.dim; k++) {
        double av[4], bv[4];
        const double l = sl[k];
#pragma unroll
        for (int i = 0; i < 4; i++) av[i] = sa[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; j++) bv[j] = sb[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { double df = av[i] - bv[j]; s[i][j] += (df * df) * l; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t r = r0 + ty * 4 + i;
        if (r >= g.na) continue;
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t c = c0 + tx * 4 + j;
            const bool pad = (r >= g.na_valid) || (c >= g.nb_valid);
            double e = pad ? 0.0 : exp_nonpos(-0.5 * s[i][j]);
            if (r == c) e = pad ? 1.0 : e + g.jitter;
            v[j] = e;
        }
        const int64_t c = c0 + tx * 4;
        double* out = g.k + (int64_t)blockIdx.z * g.strideK + r * g.ldk + c;
        if (c + 3 < g.nb && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            reinterpret_cast<double2*>(out)[0] = make_double2(v[0], v[1]);
            reinterpret_cast<double2*>(out)[1] = make_double2(v[2], v[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (c + j < g.nb) out[j] = v[j];
        }
    }
}

__global__ void inv_ell2_kernel(const double* ell, double* out, int64_t count) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) { double l = ell[k]; out[k] = 1.0 / (l * l); }
}

int launch_gram(bogp_ctx* ctx, const double* d_a, int64_t na, int64_t na_valid, const double* d_b, int64_t nb,
                int64_t nb_valid, int dim, const double* d_inv_ell2, double jitter, double* d_k, int64_t ldk,
                bool lower_tiles_only, int batch, int64_t strideK) {
    GramArgs g{d_a, d_b, d_inv_ell2, d_k, ldk, na, nb, dim, jitter, na_valid, nb_valid, lower_tiles_only ? 1 : 0,
               strideK, batch > 1 ? dim : 0};
    dim3 grid((unsigned)((nb + 63) / 64), (unsigned)((na + 63) / 64), (unsigned)batch);
    gram_kernel<<<grid, 256, 0, ctx->stream>>>(g);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

int launch_inv_ell2(bogp_ctx* ctx, const double* d_ell, double* d_out, int64_t count) {
    inv_ell2_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d_ell, d_out, count);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

int trtri_recursive(bogp_ctx* ctx, const double* d_l, int64_t ldl, int64_t strideL, double* d_w, int64_t ldw,
                    int64_t strideW, double* d_t, int64_t strideT, int64_t n, int batch, int64_t b_start);

// ------------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky, batch of `batch` matrices (strides in doubles).
// ------------------------------------------------------------------------------------------------
static int cholesky_blocked_plain(bogp_ctx* ctx, double* d_a, int64_t n, int64_t lda, int64_t strideA, double* d_w, int64_t ldw,
                               int64_t strideW, double* d_logdet, int* d_info, int batch) {
    if (n % kDiagNB != 0) { set_error("cholesky: n=%lld is not a multiple of %d", (long long)n, kDiagNB); return BOGP_ERR_BAD_ARG; }
    // Two-level right-looking blocking: an outer panel of kOuter columns is factored with
    // NB=64 steps whose SYRK only touches the panel; the big trailing update then runs once
    // per outer panel with K = kOuter (16 pipeline k-steps per tile instead of 4).
    constexpr int kOuter = 256;
    auto syrk = [&](int64_t row0, int64_t ncols, int64_t pcol0, int K) -> int {
        // C[row0.., row0..row0+ncols) -= P P^T, P = A[row0.., pcol0..pcol0+K)   (lower part)
        const int m = (int)(n - row0);
        if (m <= 0 || ncols <= 0) return BOGP_OK;
        GemmArgs s{};
        const double* P = d_a + row0 * lda + pcol0;
        s.A = P; s.lda = lda; s.strideA = strideA;
        s.B = P; s.ldb = lda; s.strideB = strideA;
        s.C = d_a + row0 * (lda + 1); s.ldc = lda; s.strideC = strideA;
        s.M = m; s.N = (int)ncols; s.K = K; s.alpha = -1.0; s.accumulate = 1; s.lower_only = 1;
        int rc = BOGP_OK;
        BOGP_PROFILED(ctx, K == kDiagNB ? 6 : 7, (rc = launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, s, batch)));
        return rc;
    };
    for (int64_t ko = 0; ko < n; ko += kOuter) {
        const int64_t wpan = (n - ko < kOuter) ? (n - ko) : kOuter;
        for (int64_t ki = 0; ki < wpan; ki += kDiagNB) {
            const int64_t k = ko + ki;
            DiagArgs dg{d_a, lda, strideA, d_w, ldw, strideW, d_logdet, d_info, (int)(k / kDiagNB)};
            if (batch >= 2 * ctx->sm_count) { BOGP_PROFILED(ctx, 4, (chol_diag_batched_kernel<<<batch, 256, 0, ctx->stream>>>(dg))); }
            else { BOGP_PROFILED(ctx, 4, (chol_diag_kernel<<<batch, 256, 0, ctx->stream>>>(dg))); }
            BOGP_LAUNCH_CHECK(ctx);
            const int below = (int)(n - (k + kDiagNB));
            if (below <= 0) break;
            double* panel = d_a + (k + kDiagNB) * lda + k;
            const double* dinv = d_w + k * (ldw + 1);
            GemmArgs t{};   // panel <- panel * Dinv^T   (in place: one CTA owns complete rows, K fully staged before the stores)
            t.A = panel; t.lda = lda; t.strideA = strideA;
            t.B = dinv;  t.ldb = ldw; t.strideB = strideW;
            t.C = panel; t.ldc = lda; t.strideC = strideA;
            t.M = below; t.N = kDiagNB; t.K = kDiagNB; t.alpha = 1.0; t.accumulate = 0; t.lower_only = 0;
            int rc = BOGP_OK;
            BOGP_PROFILED(ctx, 5, (rc = launch_gemm<128, 64, A_MK, B_NK, K_ALL>(ctx, t, batch)));
            if (rc) return rc;
            // inner update: only the remaining columns of this outer panel
            rc = syrk(k + kDiagNB, ko + wpan - (k + kDiagNB), k, kDiagNB);
            if (rc) return rc;
        }
        // outer update of everything right of the panel
        int rc = syrk(ko + wpan, n - (ko + wpan), ko, (int)wpan);
        if (rc) return rc;
    }
    return BOGP_OK;
}

// ------------------------------------------------------------------------------------------------
// One launch for the whole serial part of a 256-column panel: factor the 256x256 diagonal block
// (4 diagonal 64-blocks with their small panel solves and updates) and invert its factor (two
// recursive-doubling levels).  8 CTAs -- one thread-block cluster -- share the tile work of each phase and
// meet at the hardware cluster barrier; all cross-CTA data is read L2-coherently (cp.async.cg / ld.cg).
// This replaces 26 dependent launches per panel by one.
// ------------------------------------------------------------------------------------------------
struct InBlockArgs {
    double* a; int64_t lda;      // diagonal block A[ko.., ko..] lives at a + ko*(lda+1)
    double* w; int64_t ldw;
    double* t;                   // scratch, >= 16384 doubles
    double* logdet; int* info;
    int kblk0;                   // first 64-block index of the panel (ko / 64)
};
constexpr int kInBlockCtas = 8;

// Hardware barrier of a thread-block cluster with release/acquire semantics: every global write made by a CTA of
// the cluster before it arrives is visible to every CTA after the wait (about 0.2 us instead of the 2 us of an
// atomic counter in L2).
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The 8 CTAs are ONE thread-block cluster: co-scheduled by the hardware, meeting at the hardware cluster barrier.
__device__ __forceinline__ void inblock256_body(const InBlockArgs& g) {
    extern __shared__ __align__(16) double smem[];
    DiagSmem& dsm = *reinterpret_cast<DiagSmem*>(smem);
    const int cta = blockIdx.x;
    constexpr int NB = kDiagNB;
    double* A = g.a + (int64_t)g.kblk0 * NB * (g.lda + 1);     // 256 x 256 diagonal block
    double* W = g.w + (int64_t)g.kblk0 * NB * (g.ldw + 1);
    auto barrier = [&]() { cluster_barrier(); };
    auto tile = [&](auto fn) { __syncthreads(); fn(); };          // shared memory is reused between tiles

    for (int k = 0; k < 4; k++) {
        if (cta == 0) {
            DiagArgs dg{g.a, g.lda, 0, g.w, g.ldw, 0, g.logdet, g.info, g.kblk0 + k};
            __syncthreads();
            chol_diag_block(dg, 0, dsm);
        }
        if (k == 3) break;
        barrier();
        // panel solve: P_i = A_ik * Dinv_k^T, i = k+1..3   (in place; one CTA owns the tile)
        if (cta < 3 - k) {
            const int i = k + 1 + cta;
            GemmArgs t{};
            t.lda = g.lda; t.ldb = g.ldw; t.ldc = g.lda; t.M = NB; t.N = NB; t.K = NB; t.alpha = 1.0;
            double* Aik = A + (int64_t)i * NB * g.lda + k * NB;
            tile([&] { gemm_tile<64, 64, A_MK, B_NK, K_ALL>(t, Aik, W + (int64_t)k * NB * (g.ldw + 1), Aik, 0, 0, smem); });
        }
        barrier();
        // update: A_ij -= P_i P_j^T, k < j <= i <= 3
        {
            int idx = 0;
            for (int i = k + 1; i < 4; i++)
                for (int j = k + 1; j <= i; j++, idx++) {
                    if (idx % kInBlockCtas != cta) continue;
                    GemmArgs s{};
                    s.lda = g.lda; s.ldb = g.lda; s.ldc = g.lda; s.M = NB; s.N = NB; s.K = NB; s.alpha = -1.0; s.accumulate = 1;
                    s.lower_only = (i == j) ? 1 : 0;
                    tile([&] { gemm_tile<64, 64, A_MK, B_NK, K_ALL>(s, A + (int64_t)i * NB * g.lda + k * NB, A + (int64_t)j * NB * g.lda + k * NB,
                                                                     A + (int64_t)i * NB * g.lda + j * NB, 0, 0, smem); });
                }
        }
        barrier();
    }
    barrier();
    // inverse, level 64: pairs (0,1), (2,3):  W21 = -W22 * (L21 * W11)
    if (cta < 2) {
        const int o = 2 * cta * NB;
        double* T = g.t + cta * NB * NB;
        GemmArgs g1{};
        g1.lda = g.lda; g1.ldb = g.ldw; g1.ldc = NB; g1.M = NB; g1.N = NB; g1.K = NB; g1.alpha = 1.0;
        tile([&] { gemm_tile<64, 64, A_MK, B_KN, K_GE_N>(g1, A + (int64_t)(o + NB) * g.lda + o, W + (int64_t)o * (g.ldw + 1), T, 0, 0, smem); });
        __threadfence();
        GemmArgs g2{};
        g2.lda = g.ldw; g2.ldb = NB; g2.ldc = g.ldw; g2.M = NB; g2.N = NB; g2.K = NB; g2.alpha = -1.0;
        tile([&] { gemm_tile<64, 64, A_MK, B_KN, K_LE_M>(g2, W + (int64_t)(o + NB) * (g.ldw + 1), T, W + (int64_t)(o + NB) * g.ldw + o, 0, 0, smem); });
    }
    barrier();
    // level 128: T = L21 * W11 (128 x 128), then W21 = -W22 * T; 4 tiles each
    double* T2 = g.t + 2 * NB * NB;
    if (cta < 4) {
        GemmArgs g1{};
        g1.lda = g.lda; g1.ldb = g.ldw; g1.ldc = 2 * NB; g1.M = 2 * NB; g1.N = 2 * NB; g1.K = 2 * NB; g1.alpha = 1.0;
        tile([&] { gemm_tile<64, 64, A_MK, B_KN, K_GE_N>(g1, A + (int64_t)(2 * NB) * g.lda, W, T2, (cta >> 1) * 64, (cta & 1) * 64, smem); });
    }
    barrier();
    if (cta < 4) {
        GemmArgs g2{};
        g2.lda = g.ldw; g2.ldb = 2 * NB; g2.ldc = g.ldw; g2.M = 2 * NB; g2.N = 2 * NB; g2.K = 2 * NB; g2.alpha = -1.0;
        tile([&] { gemm_tile<64, 64, A_MK, B_KN, K_LE_M>(g2, W + (int64_t)(2 * NB) * (g.ldw + 1), T2, W + (int64_t)(2 * NB) * g.ldw, (cta >> 1) * 64, (cta & 1) * 64, smem); });
    }
}

__global__ void __cluster_dims__(kInBlockCtas, 1, 1) __launch_bounds__(256) inblock256_cluster_kernel(InBlockArgs g) { inblock256_body(g); }

// Pipelined variant of the driver below for a single matrix: the serial chain
//   [factor + invert diagonal block] -> [panel rows of the NEXT diagonal block] -> [update that block]
// runs on the high-priority second stream (small kernels, a few CTAs each), while everything bulky --
// the rest of the tall panel and the trailing SYRK -- runs on the caller's stream, one panel behind.
// Four events carry the dependencies.  All accumulation orders are the same as in the sequential
// driver, so the factor is bit-identical.
__global__ void zero_upper_blocks_kernel(double* __restrict__ a, int64_t lda, int64_t n, int blk);

struct StreamSwap {
    bogp_ctx* c; cudaStream_t saved;
    StreamSwap(bogp_ctx* c_, cudaStream_t s) : c(c_), saved(c_->stream) { c->stream = s; }
    ~StreamSwap() { c->stream = saved; }
};

static int cholesky_pipelined(bogp_ctx* ctx, double* d_a, int64_t n, int64_t lda, double* d_w, int64_t ldw, double* d_logdet,
                              int* d_info, double* d_t, int64_t* w_level) {
    constexpr int kOuter = 256;
    cudaStream_t bs = ctx->stream, cs = ctx->aux_stream;
    cudaEvent_t e_in = ctx->ev_panel[0], e_t1 = ctx->ev_panel[1], e_dn2 = ctx->ev_done[0], e_rest = ctx->ev_done[1];
    BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, bs));
    BOGP_CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_fork, 0));
    int rc = BOGP_OK;
    // Interleaved triangular inverse (n >= 1024, caller asked for it), right-looking by block rows of 256:
    //   W[i, j] = -W_ii * Acc[i, j],   Acc[i, j] = sum_{k=j}^{i-1} L[i, k] W[k, j]      (i > j, blocks)
    // As soon as block row p of W is complete, its contribution to every later row is added in ONE rank-256
    // product  Acc[p+1:, :p+1] += L[p+1:, p] * W[p, :p+1]  (the shape of the trailing SYRK: fills idle SMs on the
    // third stream), so after the last panel only the small product of the last block row remains -- instead of the
    // serial chain of one merge per level that recursive doubling leaves.  Acc is kept TRANSPOSED in the strictly
    // upper 256-blocks of A, which the factorisation never touches: Acc[i, j] lives at A[j, i].
    // Systems beyond 4096 rows are throughput-bound, not latency-bound: there the recursive-doubling merges
    // W21 = -W22 (L21 W11) (few large products, issued as soon as their inputs are final; needs a power-of-two panel
    // count) are cheaper than 64 read-modify-write passes over the accumulator.
    const int64_t npan = n / kOuter;
    const bool interleave = w_level && n >= 1024 && n % kOuter == 0;
    const bool want_doubling = n > 4096;     // measured (right-looking vs doubling): 4096 4.01 vs 4.42 ms, 8192 21.1 vs 18.6 ms, 16384 142 vs 120 ms
    const bool doubling = interleave && want_doubling && (npan & (npan - 1)) == 0;
    cudaStream_t ts = ctx->aux2_stream;
    if (interleave && !doubling) {
        BOGP_CUDA_CHECK(cudaStreamWaitEvent(ts, ctx->ev_fork, 0));
        zero_upper_blocks_kernel<<<dim3((unsigned)(n / kOuter - 1), 64), 256, 0, ts>>>(d_a, lda, n, kOuter);   // the accumulator blocks
        BOGP_LAUNCH_CHECK(ctx);
    }
    auto schedule_rightlooking = [&](int64_t pnl) -> int {        // panel pnl (and all bulk work of its iteration) has been enqueued
        BOGP_CUDA_CHECK(cudaStreamWaitEvent(ts, e_in, 0));      // W_pp (chain)
        StreamSwap sw(ctx, ts);
        const int64_t o = pnl * kOuter;
        if (pnl > 0) {                                          // block row pnl of W: -W_pp * Acc[p, :p]
            GemmArgs g2{};
            g2.A = d_w + o * (ldw + 1); g2.lda = ldw; g2.B = d_a + o; g2.ldb = lda; g2.C = d_w + o * ldw; g2.ldc = ldw;
            g2.M = kOuter; g2.N = (int)o; g2.K = kOuter; g2.alpha = -1.0;
            int r = launch_gemm<64, 64, A_MK, B_NK, K_LE_M>(ctx, g2, 1);
            if (r) return r;
        }
        if (o + kOuter < n) {                                   // Acc^T[:p+1, p+1:] += W[p, :p+1]^T * L[p+1:, p]^T
            BOGP_CUDA_CHECK(cudaStreamWaitEvent(ts, e_dn2, 0)); // L below panel pnl is final (bulk)
            GemmArgs u{};
            u.A = d_w + o * ldw; u.lda = ldw; u.B = d_a + (o + kOuter) * lda + o; u.ldb = lda; u.C = d_a + (o + kOuter); u.ldc = lda;
            u.M = (int)(o + kOuter); u.N = (int)(n - o - kOuter); u.K = kOuter; u.alpha = 1.0; u.accumulate = 1;
            return launch_gemm<128, 128, A_KM, B_NK, K_ALL>(ctx, u, 1);
        }
        return BOGP_OK;
    };
    auto schedule_doubling = [&](int64_t i) -> int {          // panel i (and all bulk work of iteration i) has been enqueued
        // third stream: the merges fill idle SMs without delaying the bulk stream (which the chain waits on)
        BOGP_CUDA_CHECK(cudaStreamWaitEvent(ts, e_in, 0));      // W blocks of panel i (chain)
        BOGP_CUDA_CHECK(cudaStreamWaitEvent(ts, e_dn2, 0));     // L rows below panel i are final (bulk)
        StreamSwap sw(ctx, ts);
        size_t toff = 131072;
        for (int64_t b = kOuter; b < n; toff += (size_t)(b * b), b *= 2) {
            const int64_t nb = b / kOuter;
            if ((i + 1) % nb != 0) break;
            const int64_t q = (i + 1) / nb - 1;                 // the size-b block that has just been completed
            double* T = d_t + toff;
            if (q % 2 == 0) {                                   // left block of its pair: T = L21 * W11
                const int64_t o = q * b;
                GemmArgs g1{};
                g1.A = d_a + (o + b) * lda + o; g1.lda = lda; g1.B = d_w + o * (ldw + 1); g1.ldb = ldw; g1.C = T; g1.ldc = b;
                g1.M = (int)b; g1.N = (int)b; g1.K = (int)b; g1.alpha = 1.0;
                return launch_gemm<128, 128, A_MK, B_KN, K_GE_N>(ctx, g1, 1);
            }
            const int64_t o = (q - 1) * b;                      // right block: W21 = -W22 * T
            GemmArgs g2{};
            g2.A = d_w + (o + b) * (ldw + 1); g2.lda = ldw; g2.B = T; g2.ldb = b; g2.C = d_w + (o + b) * ldw + o; g2.ldc = ldw;
            g2.M = (int)b; g2.N = (int)b; g2.K = (int)b; g2.alpha = -1.0;
            int r = launch_gemm<128, 128, A_MK, B_KN, K_LE_M>(ctx, g2, 1);
            if (r) return r;
        }
        return BOGP_OK;
    };
    auto schedule_trtri = [&](int64_t pnl) -> int { return doubling ? schedule_doubling(pnl) : schedule_rightlooking(pnl); };
    constexpr size_t kInBlockSmem = GemmSmem<64, 64>::bytes > sizeof(DiagSmem) ? GemmSmem<64, 64>::bytes : sizeof(DiagSmem);
    {
        static DeviceOnce configured;
        if (configured.need(ctx->device)) {
            BOGP_CUDA_CHECK(cudaFuncSetAttribute(inblock256_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kInBlockSmem));
        }
    }
    for (int64_t ko = 0; ko < n; ko += kOuter) {
        const int64_t w = (n - ko < kOuter) ? (n - ko) : kOuter;
        double* Add = d_a + ko * (lda + 1);
        double* Wdd = d_w + ko * (ldw + 1);
        if (w == kOuter) {   // ---- chain: factor and invert the diagonal block, one launch
            InBlockArgs ia{d_a, lda, d_w, ldw, d_t, d_logdet, d_info, (int)(ko / kDiagNB)};
            inblock256_cluster_kernel<<<kInBlockCtas, 256, kInBlockSmem, cs>>>(ia);
            BOGP_LAUNCH_CHECK(ctx);
        } else {
            StreamSwap sw(ctx, cs);
            for (int64_t ki = 0; ki < w; ki += kDiagNB) {
                const int64_t k = ko + ki;
                DiagArgs dg{d_a, lda, 0, d_w, ldw, 0, d_logdet, d_info, (int)(k / kDiagNB)};
                chol_diag_kernel<<<1, 256, 0, cs>>>(dg);
                BOGP_LAUNCH_CHECK(ctx);
                const int below = (int)(w - (ki + kDiagNB));
                if (below <= 0) break;
                double* panel = d_a + (k + kDiagNB) * lda + k;
                GemmArgs t{};
                t.A = panel; t.lda = lda; t.B = d_w + k * (ldw + 1); t.ldb = ldw; t.C = panel; t.ldc = lda;
                t.M = below; t.N = kDiagNB; t.K = kDiagNB; t.alpha = 1.0;
                if ((rc = launch_gemm<64, 64, A_MK, B_NK, K_ALL>(ctx, t, 1))) return rc;
                GemmArgs s{};
                s.A = panel; s.lda = lda; s.B = panel; s.ldb = lda; s.C = d_a + (k + kDiagNB) * (lda + 1); s.ldc = lda;
                s.M = below; s.N = below; s.K = kDiagNB; s.alpha = -1.0; s.accumulate = 1; s.lower_only = 1;
                if ((rc = launch_gemm<64, 64, A_MK, B_NK, K_ALL>(ctx, s, 1))) return rc;
            }
            if ((rc = trtri_recursive(ctx, Add, lda, 0, Wdd, ldw, 0, d_t, 0, w, 1, kDiagNB))) return rc;
        }
        BOGP_CUDA_CHECK(cudaEventRecord(e_in, cs));
        const int64_t row1 = ko + w;
        const int64_t below = n - row1;
        if (below <= 0) {
            if (interleave && (rc = schedule_trtri(ko / kOuter))) return rc;
            break;
        }
        const int64_t r1 = (below < kOuter) ? below : kOuter;      // rows of the next diagonal block
        const int64_t r2 = below - r1;                              // everything under it
        double* P1 = d_a + row1 * lda + ko;                         // panel rows of the next diagonal block
        double* P2 = d_a + (row1 + r1) * lda + ko;
        double* Tp = d_t + 65536;                                  // P1 staging (r1 x w, leading dimension w), past the in-block scratch
        {   // ---- chain: panel rows R1 (out of place, small tiles), then the next diagonal block
            StreamSwap sw(ctx, cs);
            if (ko > 0) BOGP_CUDA_CHECK(cudaStreamWaitEvent(cs, e_dn2, 0));     // A[R1, panel] got its last update, and Tp is free again
            GemmArgs t{};
            t.A = P1; t.lda = lda; t.B = Wdd; t.ldb = ldw; t.C = Tp; t.ldc = w;
            t.M = (int)r1; t.N = (int)w; t.K = (int)w; t.alpha = 1.0;
            if ((rc = launch_gemm<32, 32, A_MK, B_NK, K_ALL>(ctx, t, 1))) return rc;
            BOGP_CUDA_CHECK(cudaEventRecord(e_t1, cs));
            if (ko > 0) BOGP_CUDA_CHECK(cudaStreamWaitEvent(cs, e_rest, 0));    // A[R1, R1] received the previous panel's bulk update
            GemmArgs s{};
            s.A = Tp; s.lda = w; s.B = Tp; s.ldb = w; s.C = d_a + row1 * (lda + 1); s.ldc = lda;
            s.M = (int)r1; s.N = (int)r1; s.K = (int)w; s.alpha = -1.0; s.accumulate = 1; s.lower_only = 1;
            if ((rc = launch_gemm<32, 32, A_MK, B_NK, K_ALL>(ctx, s, 1))) return rc;
        }
        {   // ---- bulk, on the caller's stream
            BOGP_CUDA_CHECK(cudaStreamWaitEvent(bs, e_t1, 0));
            // P1 belongs to L: copy it into place (off the critical path)
            BOGP_CUDA_CHECK(cudaMemcpy2DAsync(P1, lda * sizeof(double), Tp, w * sizeof(double), w * sizeof(double), r1, cudaMemcpyDeviceToDevice, bs));
            if (r2 > 0) {
                BOGP_CUDA_CHECK(cudaStreamWaitEvent(bs, e_in, 0));
                GemmArgs t{};
                t.A = P2; t.lda = lda; t.B = Wdd; t.ldb = ldw; t.C = P2; t.ldc = lda;
                t.M = (int)r2; t.N = (int)w; t.K = (int)w; t.alpha = 1.0;
                // These two products are narrow (256 columns) and sit on the critical path while the trailing update is
                // still large: small tiles, so that they spread over the whole GPU (32-row tiles for the in-place product,
                // which must own complete rows; 64 x 64 tiles for the update of the next panel's column).
                if (r2 <= 64 * 148) rc = launch_gemm<32, 256, A_MK, B_NK, K_ALL>(ctx, t, 1);
                else                rc = launch_gemm<64, 256, A_MK, B_NK, K_ALL>(ctx, t, 1);
                if (rc) return rc;
                GemmArgs d{};       // A[R2, R1 columns] -= P2 P1^T
                d.A = P2; d.lda = lda; d.B = P1; d.ldb = lda; d.C = d_a + (row1 + r1) * lda + row1; d.ldc = lda;
                d.M = (int)r2; d.N = (int)r1; d.K = (int)w; d.alpha = -1.0; d.accumulate = 1;
                if (r2 * r1 <= (int64_t)128 * 128 * 148) rc = launch_gemm<32, 64, A_MK, B_NK, K_ALL>(ctx, d, 1);
                else {
                    rc = launch_gemm_tma_nt(ctx, d);
                    if (rc == 1) rc = launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, d, 1);
                }
                if (rc) return rc;
            }
            BOGP_CUDA_CHECK(cudaEventRecord(e_dn2, bs));
            if (r2 > 0) {
                GemmArgs r{};       // A[R2, R2] -= P2 P2^T (lower)
                r.A = P2; r.lda = lda; r.B = P2; r.ldb = lda; r.C = d_a + (row1 + r1) * (lda + 1); r.ldc = lda;
                r.M = (int)r2; r.N = (int)r2; r.K = (int)w; r.alpha = -1.0; r.accumulate = 1; r.lower_only = 1;
                rc = launch_gemm_tma_nt(ctx, r);
                if (rc == 1) rc = launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, r, 1);
                if (rc) return rc;
            }
            BOGP_CUDA_CHECK(cudaEventRecord(e_rest, bs));
            if (interleave && (rc = schedule_trtri(ko / kOuter))) return rc;
        }
    }
    if (w_level) *w_level = interleave ? n : kOuter;
    if (interleave) {
        BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_aux2, ts));
        BOGP_CUDA_CHECK(cudaStreamWaitEvent(bs, ctx->ev_aux2, 0));
    }
    BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, cs));            // join: the caller's stream continues after the chain
    BOGP_CUDA_CHECK(cudaStreamWaitEvent(bs, ctx->ev_fork, 0));
    return BOGP_OK;
}

// Blocked Cholesky with explicit inverses of the 256x256 diagonal blocks.
// Per outer panel of 256 columns:
//   (a) the 256x256 diagonal block is factored (NB = 64 steps inside the block only),
//   (b) its factor is inverted in place (two recursive-doubling levels),
//   (c) the whole tall panel below is ONE product  P = A_below * W_dd^T  (64 x 256 tiles, in place),
//   (d) one trailing SYRK with K = 256.
// Compared with NB = 64 steps over the full height this removes 3 of 4 tall panel solves and all
// tall inner SYRKs from the serial chain.  Needs `d_t` (trtri_scratch_doubles(n) per matrix); when
// absent the plain two-level driver is used.  On return the aligned 256-blocks of W hold the
// inverses of the corresponding blocks of L (trtri_recursive continues from block size 256).
int cholesky_blocked(bogp_ctx* ctx, double* d_a, int64_t n, int64_t lda, int64_t strideA, double* d_w, int64_t ldw,
                     int64_t strideW, double* d_logdet, int* d_info, int batch, double* d_t, int64_t strideT, int64_t* w_level) {
    if (w_level) *w_level = kDiagNB;
    if (!d_t) return cholesky_blocked_plain(ctx, d_a, n, lda, strideA, d_w, ldw, strideW, d_logdet, d_info, batch);
    {   // same shared-memory carve-out as the GEMMs around it: no SM reconfiguration between the kernels of the chain
        static DeviceOnce configured;
        if (configured.need(ctx->device)) {
            BOGP_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
    }
    if (n % kDiagNB != 0) { set_error("cholesky: n=%lld is not a multiple of %d", (long long)n, kDiagNB); return BOGP_ERR_BAD_ARG; }
    if (batch == 1 && !ctx->profile)
        return cholesky_pipelined(ctx, d_a, n, lda, d_w, ldw, d_logdet, d_info, d_t, w_level);
    constexpr int kOuter = 256;
    for (int64_t ko = 0; ko < n; ko += kOuter) {
        const int64_t wpan = (n - ko < kOuter) ? (n - ko) : kOuter;
        double* Add = d_a + ko * (lda + 1);
        double* Wdd = d_w + ko * (ldw + 1);
        // (a) factor the diagonal block
        for (int64_t ki = 0; ki < wpan; ki += kDiagNB) {
            const int64_t k = ko + ki;
            DiagArgs dg{d_a, lda, strideA, d_w, ldw, strideW, d_logdet, d_info, (int)(k / kDiagNB)};
            if (batch >= 2 * ctx->sm_count) { BOGP_PROFILED(ctx, 4, (chol_diag_batched_kernel<<<batch, 256, 0, ctx->stream>>>(dg))); }
            else { BOGP_PROFILED(ctx, 4, (chol_diag_kernel<<<batch, 256, 0, ctx->stream>>>(dg))); }
            BOGP_LAUNCH_CHECK(ctx);
            const int below = (int)(wpan - (ki + kDiagNB));          // rows of the block under this step
            if (below <= 0) break;
            double* panel = d_a + (k + kDiagNB) * lda + k;
            GemmArgs t{};
            t.A = panel; t.lda = lda; t.strideA = strideA;
            t.B = d_w + k * (ldw + 1); t.ldb = ldw; t.strideB = strideW;
            t.C = panel; t.ldc = lda; t.strideC = strideA;
            t.M = below; t.N = kDiagNB; t.K = kDiagNB; t.alpha = 1.0; t.accumulate = 0; t.lower_only = 0;
            int rc = BOGP_OK;
            BOGP_PROFILED(ctx, 5, (rc = launch_gemm<64, 64, A_MK, B_NK, K_ALL>(ctx, t, batch)));
            if (rc) return rc;
            GemmArgs s{};
            s.A = panel; s.lda = lda; s.strideA = strideA;
            s.B = panel; s.ldb = lda; s.strideB = strideA;
            s.C = d_a + (k + kDiagNB) * (lda + 1); s.ldc = lda; s.strideC = strideA;
            s.M = below; s.N = below; s.K = kDiagNB; s.alpha = -1.0; s.accumulate = 1; s.lower_only = 1;
            BOGP_PROFILED(ctx, 6, (rc = launch_gemm<64, 64, A_MK, B_NK, K_ALL>(ctx, s, batch)));
            if (rc) return rc;
        }
        // (b) W_dd = L_dd^-1
        int rc = trtri_recursive(ctx, Add, lda, strideA, Wdd, ldw, strideW, d_t, strideT, wpan, batch, kDiagNB);
        if (rc) return rc;
        const int below = (int)(n - (ko + wpan));
        if (below <= 0) break;
        // (c) tall panel: P = A_below * W_dd^T, in place (a CTA owns 64 complete rows of the panel)
        double* tall = d_a + (ko + wpan) * lda + ko;
        GemmArgs t{};
        t.A = tall; t.lda = lda; t.strideA = strideA;
        t.B = Wdd;  t.ldb = ldw; t.strideB = strideW;
        t.C = tall; t.ldc = lda; t.strideC = strideA;
        t.M = below; t.N = (int)wpan; t.K = (int)wpan; t.alpha = 1.0; t.accumulate = 0; t.lower_only = 0;
        BOGP_PROFILED(ctx, 5, (rc = launch_gemm<64, 256, A_MK, B_NK, K_ALL>(ctx, t, batch)));
        if (rc) return rc;
        // (d) trailing update.  With look-ahead (single matrix, not profiling) only the columns of the NEXT
        //     panel are updated on the main stream; the bulk of the SYRK runs on the second stream while the
        //     next panel -- a serial chain that occupies a handful of SMs -- is being factored.
        const int64_t row0 = ko + wpan;
        const bool lookahead = (batch == 1) && !ctx->profile;
        const int64_t nextw = lookahead ? ((below < kOuter) ? below : kOuter) : below;
        if (lookahead && ko > 0) BOGP_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_done[0], 0));   // bulk of the previous panel
        GemmArgs s{};
        s.A = tall; s.lda = lda; s.strideA = strideA;
        s.B = tall; s.ldb = lda; s.strideB = strideA;
        s.C = d_a + row0 * (lda + 1); s.ldc = lda; s.strideC = strideA;
        s.M = below; s.N = (int)nextw; s.K = (int)wpan; s.alpha = -1.0; s.accumulate = 1; s.lower_only = 1;
        if (batch == 1) { BOGP_PROFILED(ctx, 7, (rc = launch_gemm_tma_nt(ctx, s))); } else rc = 1;
        if (rc == 1) { BOGP_PROFILED(ctx, 7, (rc = launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, s, batch))); }
        if (rc) return rc;
        if (lookahead && below > nextw) {
            BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_panel[0], ctx->stream));
            BOGP_CUDA_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_panel[0], 0));
            const int64_t row1 = row0 + nextw;
            const double* P1 = d_a + row1 * lda + ko;
            GemmArgs r{};
            r.A = P1; r.lda = lda; r.B = P1; r.ldb = lda;
            r.C = d_a + row1 * (lda + 1); r.ldc = lda;
            r.M = (int)(n - row1); r.N = (int)(n - row1); r.K = (int)wpan; r.alpha = -1.0; r.accumulate = 1; r.lower_only = 1;
            cudaStream_t main_stream = ctx->stream;
            ctx->stream = ctx->aux_stream;
            rc = launch_gemm_tma_nt(ctx, r);
            if (rc == 1) rc = launch_gemm<128, 128, A_MK, B_NK, K_ALL>(ctx, r, 1);
            ctx->stream = main_stream;
            if (rc) return rc;
            BOGP_CUDA_CHECK(cudaEventRecord(ctx->ev_done[0], ctx->aux_stream));
            if (row1 + kOuter >= n) BOGP_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_done[0], 0));   // last bulk: join now
        }
    }
    return BOGP_OK;
}

// ------------------------------------------------------------------------------------------------
// W = L^-1 by recursive doubling: the 64-blocks on the diagonal of W are already inverted
// (chol_diag_kernel); each level merges neighbouring inverted blocks,
//   W21 = -W22 * (L21 * W11),
// as two batched GEMMs that skip the structurally-zero k range.
// `d_t` needs trtri_scratch_doubles(n) doubles per matrix.
// ------------------------------------------------------------------------------------------------
// Scratch of cholesky_blocked + trtri_recursive per matrix: the recursion's T blocks, the fused in-block
// kernel's T blocks (2*64^2 + 128^2) and, past offset 65536, the staged panel rows P1 (256 x 256).
size_t cholesky_scratch_doubles(int64_t n) {
    const size_t t = trtri_scratch_doubles(n);
    const size_t c = n > 256 ? 131072 : 24576;
    return t > c ? t : c;
}

size_t trtri_scratch_doubles(int64_t n) {
    size_t need = 0;
    for (int64_t b = kDiagNB; b < n; b *= 2) {
        int64_t full = 0; bool ragged = false;
        for (int64_t o = 0; o + b < n; o += 2 * b) { if (o + 2 * b <= n) full++; else ragged = true; }
        size_t t = (size_t)(full + (ragged ? 1 : 0)) * b * b;
        if (t > need) need = t;
    }
    return need;
}

int trtri_recursive(bogp_ctx* ctx, const double* d_l, int64_t ldl, int64_t strideL, double* d_w, int64_t ldw,
                    int64_t strideW, double* d_t, int64_t strideT, int64_t n, int batch, int64_t b_start) {
    for (int64_t b = b_start; b < n; b *= 2) {
        int64_t full = 0; int64_t ragged_o = -1;
        for (int64_t o = 0; o + b < n; o += 2 * b) { if (o + 2 * b <= n) full++; else ragged_o = o; }
        for (int pass = 0; pass < 2; pass++) {
            int64_t npairs, o0, r;
            if (pass == 0) { npairs = full; o0 = 0; r = b; }
            else { if (ragged_o < 0) break; npairs = 1; o0 = ragged_o; r = n - ragged_o - b; }
            if (npairs == 0) continue;
            {
                // blockIdx.z = matrix * npairs + pair
                const double* L  = d_l;
                double*       W  = d_w;
                double*       T  = d_t + (pass == 1 ? (size_t)full * b * b : 0);
                GemmArgs g1{};
                g1.inner = (int)npairs; g1.strideA2 = strideL; g1.strideB2 = strideW; g1.strideC2 = strideT;
                g1.A = L + (o0 + b) * ldl + o0; g1.lda = ldl; g1.strideA = 2 * b * (ldl + 1);
                g1.B = W + o0 * (ldw + 1);      g1.ldb = ldw; g1.strideB = 2 * b * (ldw + 1);
                g1.C = T;                        g1.ldc = b;   g1.strideC = b * b;
                g1.M = (int)r; g1.N = (int)b; g1.K = (int)b; g1.alpha = 1.0; g1.accumulate = 0; g1.lower_only = 0;
                int rc = (b <= 128) ? launch_gemm<64, 64, A_MK, B_KN, K_GE_N>(ctx, g1, (int)npairs * batch)
                                    : launch_gemm<128, 128, A_MK, B_KN, K_GE_N>(ctx, g1, (int)npairs * batch);
                if (rc) return rc;
                GemmArgs g2{};
                g2.inner = (int)npairs; g2.strideA2 = strideW; g2.strideB2 = strideT; g2.strideC2 = strideW;
                g2.A = W + (o0 + b) * (ldw + 1); g2.lda = ldw; g2.strideA = 2 * b * (ldw + 1);
                g2.B = T;                         g2.ldb = b;   g2.strideB = b * b;
                g2.C = W + (o0 + b) * ldw + o0;   g2.ldc = ldw; g2.strideC = 2 * b * (ldw + 1);
                g2.M = (int)r; g2.N = (int)b; g2.K = (int)r; g2.alpha = -1.0; g2.accumulate = 0; g2.lower_only = 0;
                rc = (b <= 128) ? launch_gemm<64, 64, A_MK, B_KN, K_LE_M>(ctx, g2, (int)npairs * batch)
                                : launch_gemm<128, 128, A_MK, B_KN, K_LE_M>(ctx, g2, (int)npairs * batch);
                if (rc) return rc;
            }
        }
    }
    return BOGP_OK;
}

// ------------------------------------------------------------------------------------------------
// alpha = W^T (W y) with fixed summation order, y^T alpha, nlml.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ w, int64_t ldw, int64_t strideW,
                                                         const double* __restrict__ y, double* __restrict__ v, int n) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    w += blockIdx.y * strideW; v += (int64_t)blockIdx.y * n;     // y is shared by the batch
    const double* wr = w + (int64_t)row * ldw;
    double s = 0.0;
    for (int j = lane; j <= row; j += 32) s += wr[j] * y[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) v[row] = s;
}

// alpha_j = sum_{i >= j} W[i][j] v[i] in two stages with a fixed summation order.
// Stage 1: CTA (cb, rc) = 32 columns x 256 rows (8 warps x 32 rows, lane = column: 256-byte row segments);
//          the 8 warp sums are combined in order -> part[rc][col].  Tiles above the diagonal are skipped.
// Stage 2: alpha[col] = sum over row chunks rc >= col / 256, ascending.
constexpr int kTrmvRows = 256;
__global__ void __launch_bounds__(256) trmv_lower_t_part_kernel(const double* __restrict__ w, int64_t ldw, int64_t strideW,
                                                                const double* __restrict__ v, double* __restrict__ part, int n, int nrc) {
    __shared__ double sm[8][33];
    const int cb = blockIdx.x, rc = blockIdx.y, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int c0 = cb * 32, r0 = rc * kTrmvRows;
    if (r0 + kTrmvRows <= c0) return;                              // entirely above the diagonal
    w += blockIdx.z * strideW; v += (int64_t)blockIdx.z * n; part += (int64_t)blockIdx.z * nrc * n;
    const int col = c0 + lane;
    const int rb = r0 + wp * 32, re = min(n, rb + 32);
    double s = 0.0;
    if (col < n) {
#pragma unroll 8
        for (int i = max(rb, col); i < re; i++) s += w[(int64_t)i * ldw + col] * v[i];
    }
    sm[wp][lane] = s;
    __syncthreads();
    if (wp == 0 && col < n) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += sm[k][lane];
        part[(int64_t)rc * n + col] = t;
    }
}
__global__ void __launch_bounds__(256) trmv_lower_t_sum_kernel(const double* __restrict__ part, double* __restrict__ alpha, int n, int nrc) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= n) return;
    part += (int64_t)blockIdx.y * nrc * n; alpha += (int64_t)blockIdx.y * n;
    double t = 0.0;
    for (int rc = col / kTrmvRows; rc < nrc; rc++) t += part[(int64_t)rc * n + col];
    alpha[col] = t;
}

size_t alpha_scratch_doubles(int64_t n) { return (size_t)((n + kTrmvRows - 1) / kTrmvRows) * (size_t)n; }

// d_part: alpha_scratch_doubles(n) doubles per batch entry
int launch_alpha(bogp_ctx* ctx, const double* d_w, int64_t ldw, int64_t strideW, const double* d_y, double* d_v,
                 double* d_alpha, double* d_part, int n, int batch) {
    trmv_lower_kernel<<<dim3((unsigned)((n + 7) / 8), batch), 256, 0, ctx->stream>>>(d_w, ldw, strideW, d_y, d_v, n);
    BOGP_LAUNCH_CHECK(ctx);
    const int nrc = (n + kTrmvRows - 1) / kTrmvRows;
    trmv_lower_t_part_kernel<<<dim3((unsigned)((n + 31) / 32), nrc, batch), 256, 0, ctx->stream>>>(d_w, ldw, strideW, d_v, d_part, n, nrc);
    BOGP_LAUNCH_CHECK(ctx);
    trmv_lower_t_sum_kernel<<<dim3((unsigned)((n + 255) / 256), batch), 256, 0, ctx->stream>>>(d_part, d_alpha, n, nrc);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

// scalars[1] = y . alpha ; scalars[2] = 0.5*(y.alpha + logdet + n log 2pi)   (point_selector.py:119)
__global__ void __launch_bounds__(256) nlml_finish_kernel(const double* __restrict__ y, const double* __restrict__ alpha, int n_pad,
                                                          int n, double* scalars) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n_pad; i += 256) s += y[i] * alpha[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) {
        scalars[1] = red[0];
        scalars[2] = 0.5 * (red[0] + scalars[0] + (double)n * log(2.0 * 3.14159265358979323846));
    }
}

// ------------------------------------------------------------------------------------------------
// Pack W (row-major, lower) into DMMA-fragment order for the acquisition kernel:
// tile (ib, kt) = rows [256 ib, +256) x k [16 kt, +16), 0 <= kt < 16 (ib+1); inside a tile
// element (m, k) sits at ((k/4 % 4) * 32 + m/8) * 32 + (m % 8) * 4 + k % 4, so that a warp's
// A fragment is one contiguous 256-byte line and a whole tile is one 32 KB bulk copy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_w_kernel(const double* __restrict__ w, int64_t ldw, double* __restrict__ wp) {
    __shared__ double t[kAcqBM][kAcqKB + 1];
    const int ib = blockIdx.y, kt = blockIdx.x;
    if (kt >= (ib + 1) * (kAcqBM / kAcqKB)) return;
    const int64_t tile = (int64_t)ib * (ib + 1) / 2 * (kAcqBM / kAcqKB) + kt;
    const double* src = w + (int64_t)ib * kAcqBM * ldw + (int64_t)kt * kAcqKB;
    for (int i = threadIdx.x; i < kAcqBM * kAcqKB; i += 256) {
        int m = i / kAcqKB, k = i % kAcqKB;
        t[m][k] = src[(int64_t)m * ldw + k];
    }
    __syncthreads();
    double* dst = wp + tile * (kAcqBM * kAcqKB);
    for (int i = threadIdx.x; i < kAcqBM * kAcqKB; i += 256) {
        int lane = i & 31, m8 = (i >> 5) & 31, kk = i >> 10;
        dst[i] = t[m8 * 8 + (lane >> 2)][kk * 4 + (lane & 3)];
    }
}

size_t packed_w_doubles(int64_t n_pad) {
    const int64_t nI = n_pad / kAcqBM;
    return (size_t)(nI * (nI + 1) / 2) * kAcqBM * kAcqBM;
}

__global__ void pad_copy_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n, int64_t n_pad, int width) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * width) return;
    dst[i] = (i / width < n) ? src[i] : 0.0;
}

// One launch at the start of a fit: zero-padded copies of X and y, 1/ell^2, cleared status words.
__global__ void __launch_bounds__(256) fit_prepare_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ ell,
                                                          double* __restrict__ x_pad, double* __restrict__ y_pad, double* __restrict__ inv_ell2,
                                                          double* __restrict__ scalars, int* __restrict__ info, int64_t n, int64_t n_pad, int dim) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad * dim) x_pad[i] = (i / dim < n) ? x[i] : 0.0;
    if (i < n_pad) y_pad[i] = (i < n) ? y[i] : 0.0;
    if (i < dim) { const double l = ell[i]; inv_ell2[i] = 1.0 / (l * l); }
    if (i < 64) { scalars[i] = 0.0; info[i] = 0; }
}

// zero the strictly upper blk x blk blocks of a row-major matrix (accumulators of the interleaved triangular inverse)
__global__ void __launch_bounds__(256) zero_upper_blocks_kernel(double* __restrict__ a, int64_t lda, int64_t n, int blk) {
    const int64_t j = blockIdx.x;                      // block row
    const int64_t c0 = (j + 1) * blk, w2 = (n - c0) / 2;   // columns [c0, n), as double2 (blk and n are even)
    for (int64_t r = blockIdx.y; r < blk; r += gridDim.y) {
        double2* row = reinterpret_cast<double2*>(a + (j * blk + r) * lda + c0);
        for (int64_t c = threadIdx.x; c < w2; c += 256) row[c] = make_double2(0.0, 0.0);
    }
}

__global__ void zero_kernel(double* p, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = 0.0;
}

}  // namespace bogp

using namespace bogp;

// ================================================================================================
// C ABI
// ================================================================================================
struct bogp_fit {
    int64_t n, n_pad; int dim;
    double *x_pad, *y_pad, *inv_ell2, *a, *w, *wp, *t, *alpha, *v, *scalars, *wscale;
    uint8_t* wq; int* wexp;
    int* info;
    double jitter;
    bogp_ctx* ctx;
    double* apart;                 // partial sums of the transposed triangular product (launch_alpha)
    bool wp_ready, wq_ready;       // which operand packings of W exist (fit_ensure_packed)
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct FitLayout { size_t x, y, ell, a, w, wp, alpha, v, apart, scal, info, wq, wexp, wscale, total; };
static FitLayout fit_layout(int64_t n, int dim) {
    const int64_t np = (n + kPad - 1) / kPad * kPad;
    FitLayout l{}; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    l.x = take(np * dim * 8); l.y = take(np * 8); l.ell = take(BOGP_MAX_DIM * 8);
    l.a = take((size_t)np * np * 8); l.w = take((size_t)np * np * 8);
    size_t shared = packed_w_doubles(np); size_t tneed = cholesky_scratch_doubles(np);
    // the trtri scratch and the packed W are never live at the same time -> but keep them
    // separate when the scratch is the larger one (ragged block counts)
    l.wp = take((shared > tneed ? shared : tneed) * 8);
    l.alpha = take(np * 8); l.v = take(np * 8); l.apart = take(alpha_scratch_doubles(np) * 8); l.scal = take(64 * 8); l.info = take(64 * 4);
    l.wq = take(i8_wq_bytes(np)); l.wexp = take(np * 4); l.wscale = take(np * 8);
    l.total = off;
    return l;
}

extern "C" size_t bogp_fit_workspace_bytes(int64_t n, int dim) {
    if (n <= 0 || dim <= 0 || dim > BOGP_MAX_DIM) return 0;
    return fit_layout(n, dim).total;
}

extern "C" int bogp_kernel_matrix(bogp_ctx* ctx, const double* d_a, int64_t na, const double* d_b, int64_t nb, int dim,
                                  const double* d_ell, double jitter, double* d_k, int64_t ldk) {
    if (!ctx || !d_a || !d_b || !d_ell || !d_k || na <= 0 || nb <= 0 || dim <= 0 || dim > BOGP_MAX_DIM || ldk < nb) {
        set_error("bogp_kernel_matrix: bad argument"); return BOGP_ERR_BAD_ARG;
    }
    int rc = launch_inv_ell2(ctx, d_ell, ctx->d_scalars + 32, dim);
    if (rc) return rc;
    return launch_gram(ctx, d_a, na, na, d_b, nb, nb, dim, ctx->d_scalars + 32, jitter, d_k, ldk, false, 1, 0);
}

extern "C" int bogp_cholesky(bogp_ctx* ctx, double* d_a, int64_t n, int64_t lda, double* d_linv, double* d_logdet, int* d_info) {
    if (!ctx || !d_a || !d_linv || !d_logdet || !d_info || n <= 0 || lda < n) { set_error("bogp_cholesky: bad argument"); return BOGP_ERR_BAD_ARG; }
    BOGP_CUDA_CHECK(cudaMemsetAsync(d_logdet, 0, sizeof(double), ctx->stream));
    BOGP_CUDA_CHECK(cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
    return cholesky_blocked(ctx, d_a, n, lda, 0, d_linv, lda, 0, d_logdet, d_info, 1, nullptr, 0, nullptr);
}

static int fit_enqueue(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim, const double* d_ell,
                       double jitter, void* d_workspace, size_t workspace_bytes, bogp_fit** out);

// Read back the status of an enqueued fit (synchronises the stream).
extern "C" int bogp_fit_status(bogp_fit* f, double* h_nlml_out) {
    if (!f) { set_error("bogp_fit_status: null fit"); return BOGP_ERR_BAD_ARG; }
    // scalars (64 doubles) and info (64 ints) are adjacent in the workspace (fit_layout): one copy
    struct { double sc[64]; int info; } h;
    cudaStream_t st = f->ctx->stream;
    static_assert(sizeof(double) * 64 == 512, "layout");
    if (reinterpret_cast<const char*>(f->info) != reinterpret_cast<const char*>(f->scalars) + 512) { set_error("bogp_fit_status: unexpected workspace layout"); return BOGP_ERR_BAD_ARG; }
    BOGP_CUDA_CHECK(cudaMemcpyAsync(&h, f->scalars, 512 + sizeof(int), cudaMemcpyDeviceToHost, st));
    BOGP_CUDA_CHECK(cudaStreamSynchronize(st));
    const int info = h.info; const double* sc = h.sc;
    if (info != 0) {
        set_error("bogp_fit: matrix not positive definite (pivot %d of %lld)", info, (long long)f->n);
        return BOGP_ERR_NOT_POSDEF;
    }
    if (h_nlml_out) *h_nlml_out = sc[2];
    return BOGP_OK;
}

// Enqueue all device work of a fit without synchronising (stream-capturable: the host side can
// record it into a CUDA graph and replay it; bogp_fit_status reads the result).
extern "C" int bogp_fit_enqueue(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim, const double* d_ell,
                                double jitter, void* d_workspace, size_t workspace_bytes, bogp_fit** out) {
    return fit_enqueue(ctx, d_x, d_y, n, dim, d_ell, jitter, d_workspace, workspace_bytes, out);
}

extern "C" int bogp_fit_create(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim, const double* d_ell,
                               double jitter, void* d_workspace, size_t workspace_bytes, bogp_fit** out, double* h_nlml_out) {
    bogp_fit* f = nullptr;
    int rc = fit_enqueue(ctx, d_x, d_y, n, dim, d_ell, jitter, d_workspace, workspace_bytes, &f);
    if (rc) return rc;
    rc = bogp_fit_status(f, h_nlml_out);
    if (rc) { delete f; return rc; }
    *out = f;
    return BOGP_OK;
}

static int fit_enqueue(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim, const double* d_ell,
                       double jitter, void* d_workspace, size_t workspace_bytes, bogp_fit** out) {
    if (!ctx || !d_x || !d_y || !d_ell || !d_workspace || !out || n <= 0 || dim <= 0 || dim > BOGP_MAX_DIM) {
        set_error("bogp_fit_create: bad argument"); return BOGP_ERR_BAD_ARG;
    }
    NvtxRange nvtx("bogp fit: Gram + Cholesky + L^-1 + alpha + nlml");
    const FitLayout l = fit_layout(n, dim);
    if (workspace_bytes < l.total) { set_error("bogp_fit_create: workspace %zu < %zu bytes", workspace_bytes, l.total); return BOGP_ERR_WORKSPACE; }
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) != 0) { set_error("bogp_fit_create: workspace must be 256-byte aligned"); return BOGP_ERR_BAD_ARG; }
    char* base = static_cast<char*>(d_workspace);
    bogp_fit* f = new bogp_fit();
    f->ctx = ctx; f->n = n; f->dim = dim; f->jitter = jitter;
    const int64_t np = (n + kPad - 1) / kPad * kPad; f->n_pad = np;
    f->x_pad = (double*)(base + l.x); f->y_pad = (double*)(base + l.y); f->inv_ell2 = (double*)(base + l.ell);
    f->a = (double*)(base + l.a); f->w = (double*)(base + l.w); f->wp = (double*)(base + l.wp); f->t = f->wp;
    f->wq = (uint8_t*)(base + l.wq); f->wexp = (int*)(base + l.wexp); f->wscale = (double*)(base + l.wscale);
    f->alpha = (double*)(base + l.alpha); f->v = (double*)(base + l.v); f->apart = (double*)(base + l.apart); f->scalars = (double*)(base + l.scal); f->info = (int*)(base + l.info);
    cudaStream_t st = ctx->stream;
    int rc;
#define FIT_TRY(e) do { rc = (e); if (rc) { delete f; return rc; } } while (0)
#define FIT_CUDA(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { set_error("%s: %s", #e, cudaGetErrorString(_e)); delete f; return BOGP_ERR_CUDA; } } while (0)
    // W is cleared on the third stream while the Gram matrix is built
    FIT_CUDA(cudaEventRecord(ctx->ev_aux2, st));
    FIT_CUDA(cudaStreamWaitEvent(ctx->aux2_stream, ctx->ev_aux2, 0));
    FIT_CUDA(cudaMemsetAsync(f->w, 0, (size_t)np * np * 8, ctx->aux2_stream));
    FIT_CUDA(cudaEventRecord(ctx->ev_aux2, ctx->aux2_stream));
    fit_prepare_kernel<<<(unsigned)((np * dim + 255) / 256), 256, 0, st>>>(d_x, d_y, d_ell, f->x_pad, f->y_pad, f->inv_ell2, f->scalars, f->info,
                                                                         n, np, dim); ctx->launches++;
    // K1 (lower tiles; identity in the padding)
    FIT_TRY(launch_gram(ctx, f->x_pad, np, n, f->x_pad, np, n, dim, f->inv_ell2, jitter, f->a, np, true, 1, 0));
    FIT_CUDA(cudaStreamWaitEvent(st, ctx->ev_aux2, 0));
    // K2
    int64_t w_level = 0;     // block size up to which W = L^-1 is already complete
    FIT_TRY(cholesky_blocked(ctx, f->a, np, np, 0, f->w, np, 0, f->scalars, f->info, 1, f->t, 0, &w_level));
    if (w_level < np) FIT_TRY(trtri_recursive(ctx, f->a, np, 0, f->w, np, 0, f->t, 0, np, 1, w_level < 256 ? 256 : w_level));
    // W is final.  Its operand packing for the selected tensor path (which overwrites the trtri scratch) goes to
    // the third stream, next to alpha = W^T W y and the marginal likelihood on this one.
    f->wp_ready = f->wq_ready = false;
    FIT_CUDA(cudaEventRecord(ctx->ev_aux2, st));
    FIT_CUDA(cudaStreamWaitEvent(ctx->aux2_stream, ctx->ev_aux2, 0));
    {
        StreamSwap sw(ctx, ctx->aux2_stream);
        const int path = (ctx->acquire_path == BOGP_PATH_INT8_TCGEN05 && np <= 16384) ? BOGP_PATH_INT8_TCGEN05 : BOGP_PATH_FP64_DMMA;
        FIT_TRY(fit_ensure_packed(ctx, f, path));
    }
    FIT_CUDA(cudaEventRecord(ctx->ev_aux2, ctx->aux2_stream));
    FIT_TRY(launch_alpha(ctx, f->w, np, 0, f->y_pad, f->v, f->alpha, f->apart, (int)np, 1));
    nlml_finish_kernel<<<1, 256, 0, st>>>(f->y_pad, f->alpha, (int)np, (int)n, f->scalars); ctx->launches++;
    FIT_CUDA(cudaStreamWaitEvent(st, ctx->ev_aux2, 0));
    FIT_CUDA(cudaGetLastError());
    *out = f;
    return BOGP_OK;
#undef FIT_TRY
#undef FIT_CUDA
}

extern "C" void bogp_fit_destroy(bogp_fit* fit) { delete fit; }
extern "C" int64_t bogp_fit_n_pad(const bogp_fit* fit) { return fit ? fit->n_pad : 0; }
extern "C" const double* bogp_fit_chol(const bogp_fit* fit) { return fit ? fit->a : nullptr; }
extern "C" const double* bogp_fit_linv(const bogp_fit* fit) { return fit ? fit->w : nullptr; }
extern "C" const double* bogp_fit_alpha(const bogp_fit* fit) { return fit ? fit->alpha : nullptr; }
extern "C" double bogp_fit_logdet(const bogp_fit* fit) {
    if (!fit) return 0.0;
    double v = 0.0;
    cudaMemcpyAsync(&v, fit->scalars, sizeof(double), cudaMemcpyDeviceToHost, fit->ctx->stream);
    cudaStreamSynchronize(fit->ctx->stream);
    return v;
}

// accessors used by acquire.cu
namespace bogp {
int fit_ensure_packed(bogp_ctx* ctx, const bogp_fit* cf, int path) {
    bogp_fit* f = const_cast<bogp_fit*>(cf);
    const int64_t np = f->n_pad;
    if (path == BOGP_PATH_INT8_TCGEN05) {
        if (f->wq_ready) return BOGP_OK;
        int rc = launch_slice_w(ctx, f->w, np, f->wexp, f->wscale, f->wq);      // digit tiles of W for the INT8 tensor path
        if (rc) return rc;
        f->wq_ready = true;
    } else {
        if (f->wp_ready) return BOGP_OK;
        const int nI = (int)(np / kAcqBM);
        pack_w_kernel<<<dim3(nI * (kAcqBM / kAcqKB), nI), 256, 0, ctx->stream>>>(f->w, np, f->wp);   // fragment-packed W for the DMMA path
        BOGP_LAUNCH_CHECK(ctx);
        f->wp_ready = true;
    }
    return BOGP_OK;
}
const double* fit_wp(const bogp_fit* f) { return f->wp; }
const uint8_t* fit_wq(const bogp_fit* f) { return f->wq; }
const double* fit_wscale(const bogp_fit* f) { return f->wscale; }
const double* fit_xpad(const bogp_fit* f) { return f->x_pad; }
const double* fit_inv_ell2(const bogp_fit* f) { return f->inv_ell2; }
const double* fit_alpha(const bogp_fit* f) { return f->alpha; }
int64_t fit_n(const bogp_fit* f) { return f->n; }
double fit_jitter(const bogp_fit* f) { return f->jitter; }
int fit_dim(const bogp_fit* f) { return f->dim; }
}
