// K3: batched negative log marginal likelihood (+ analytic gradient) for multi-restart
// ARD length-scale fitting.                                      point_selector.py:111-138
//
// The reference evaluates its 50x50 (or 1-D) length-scale grid in a Python double loop, each
// iteration building K, inverting it and taking a determinant (`eval_log_marginal`).  Here all
// R length-scale vectors are evaluated at once:
//   * n <= 64  : one CTA per restart, everything in shared memory (the reference's native
//                sizes, M <= 21 measured points x 2500 grid points);
//   * n  > 64  : batched Gram build, batched blocked Cholesky (DMMA trailing updates), batched
//                recursive triangular inverse, K^-1 = W^T W on the tensor path, and a fused
//                gradient contraction.
// nlml = 0.5*(y^T K^-1 y + log det K + n log 2pi) with log det = 2 sum log L_ii.  Deliberate deviation: the
// reference's np.log(np.linalg.det(.)) underflows to -inf as soon as det K < 2^-1074 -- with the 1e-4 jitter from about
// M = 80-90 measured points on its own grids (earlier for long length scales), not only at N >~ 1000 -- and its float32
// table then selects the first -inf cell; this path selects the true minimiser (tests/test_gpu_configs.py pins both at M = 120).
// Gradient (extension; jitter not differentiated), with alpha = K^-1 y:
//   d nlml/d ell_k = -0.5/ell_k^3 * sum_ij (alpha_i alpha_j - K^-1_ij) k_ij (x_ik - x_jk)^2
#include "common.cuh"
#include "gemm_f64.cuh"
#include "fit.cuh"

namespace bogp {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;

// ------------------------------------------------------------------------------------------------
// small path: n <= 64
// ------------------------------------------------------------------------------------------------
struct SmallArgs {
    const double* x; const double* y; const double* ell;
    double* nlml; double* grad; int* info;
    int n, dim; double jitter;
};

constexpr int SLD = 65;
constexpr size_t kSmallSmem = (size_t)(2 * 64 * SLD + BOGP_MAX_DIM * 64 + 4 * 64 + 2 * BOGP_MAX_DIM) * sizeof(double);

__global__ void __launch_bounds__(256) nlml_small_kernel(SmallArgs g) {
    extern __shared__ __align__(16) double sm[];
    double* a  = sm;                    // K, then L            [64][65]
    double* w  = a + 64 * SLD;          // W = L^-1             [64][65]
    double* xs = w + 64 * SLD;          // points, transposed   [dim][64]
    double* ys = xs + BOGP_MAX_DIM * 64;
    double* dg = ys + 64;
    double* v  = dg + 64;
    double* al = v + 64;
    double* il = al + 64;               // 1/ell^2
    double* l3 = il + BOGP_MAX_DIM;     // 1/ell^3
    const int tid = threadIdx.x, n = g.n, dim = g.dim;
    const int64_t r = blockIdx.x;

    for (int i = tid; i < n * dim; i += 256) xs[(i % dim) * 64 + i / dim] = g.x[i];
    if (tid < n) ys[tid] = g.y[tid];
    if (tid < dim) { double l = g.ell[r * dim + tid]; il[tid] = 1.0 / (l * l); l3[tid] = 1.0 / (l * l * l); }
    for (int i = tid; i < 64 * SLD; i += 256) { a[i] = 0.0; w[i] = 0.0; }
    __syncthreads();
    for (int e = tid; e < n * n; e += 256) {
        const int i = e / n, j = e % n;
        if (j > i) continue;
        double s = 0.0;
        for (int k = 0; k < dim; k++) { const double df = xs[k * 64 + i] - xs[k * 64 + j]; s += (df * df) * il[k]; }
        double kv = exp_nonpos(-0.5 * s);
        if (i == j) kv += g.jitter;
        a[i * SLD + j] = kv;
    }
    bool bad = false;
    for (int j = 0; j < n; j++) {
        __syncthreads();
        const double ajj = a[j * SLD + j];
        if (!(ajj > 0.0) || isinf(ajj)) bad = true;
        const double d = sqrt(ajj);
        if (tid == j) dg[j] = d;
        if (tid > j && tid < n) a[tid * SLD + j] = a[tid * SLD + j] / d;
        __syncthreads();
        const int rem = n - 1 - j;
        for (int idx = tid; idx < rem * rem; idx += 256) {
            const int i = j + 1 + idx / rem, l = j + 1 + idx % rem;
            if (l <= i) a[i * SLD + l] -= a[i * SLD + j] * a[l * SLD + j];
        }
    }
    __syncthreads();
    if (bad && tid == 0) g.info[r] = 1;
    {   // W = L^-1 (columns; 4 lanes per column, partial sums combined by shuffles)
        const int c = tid >> 2, q = tid & 3;
        if (q == 0 && c < n) w[c * SLD + c] = 1.0 / dg[c];
        __syncwarp();
        for (int i = 1; i < n; i++) {
            double s = 0.0;
            if (i > c && c < n) for (int k = c + q; k < i; k += 4) s += a[i * SLD + k] * w[k * SLD + c];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (i > c && c < n && q == 0) w[i * SLD + c] = -s / dg[i];
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid < n) { double s = 0.0; for (int j = 0; j <= tid; j++) s += w[tid * SLD + j] * ys[j]; v[tid] = s; }
    __syncthreads();
    if (tid < n) { double s = 0.0; for (int i = tid; i < n; i++) s += w[i * SLD + tid] * v[i]; al[tid] = s; }
    __syncthreads();
    if (tid == 0) {
        double quad = 0.0, logdet = 0.0;
        for (int i = 0; i < n; i++) { quad += ys[i] * al[i]; logdet += 2.0 * log(dg[i]); }
        g.nlml[r] = 0.5 * (quad + logdet + (double)n * kLog2Pi);
    }
    if (!g.grad) return;
    // gradient: pairs i > j (the diagonal contributes nothing), doubled
    double acc[BOGP_MAX_DIM];
#pragma unroll
    for (int k = 0; k < BOGP_MAX_DIM; k++) acc[k] = 0.0;
    for (int e = tid; e < n * n; e += 256) {
        const int i = e / n, j = e % n;
        if (j >= i) continue;
        double kinv = 0.0;
        for (int k = i; k < n; k++) kinv += w[k * SLD + i] * w[k * SLD + j];
        double s = 0.0;
        for (int k = 0; k < dim; k++) { const double df = xs[k * 64 + i] - xs[k * 64 + j]; s += (df * df) * il[k]; }
        const double c = (al[i] * al[j] - kinv) * exp_nonpos(-0.5 * s);
#pragma unroll
        for (int k = 0; k < BOGP_MAX_DIM; k++) if (k < dim) { const double df = xs[k * 64 + i] - xs[k * 64 + j]; acc[k] += c * (df * df); }
    }
    __syncthreads();            // L (array a) is dead: reuse it for the per-thread partials [256][dim]
#pragma unroll
    for (int k = 0; k < BOGP_MAX_DIM; k++) if (k < dim) a[tid * dim + k] = acc[k];
    __syncthreads();
    if (tid < dim) {
        double s = 0.0;
        for (int t = 0; t < 256; t++) s += a[t * dim + tid];
        g.grad[r * dim + tid] = -(s * l3[tid]);     // -0.5 * 2 * sum_{i>j}
    }
}

// ------------------------------------------------------------------------------------------------
// large path pieces
// ------------------------------------------------------------------------------------------------
// gpart[r][bi][k] = sum over the pairs (i > j) of block row bi (64 rows, tiles bj = 0..bi) of
//   (alpha_i alpha_j - Kinv_ij) k_ij (x_ik - x_jk)^2.
// grid (64-row blocks, restarts); 256 threads, 4x4 pairs per thread and tile, the CTA walks the tiles of its block
// row (row coordinates staged once, column coordinates per tile), ONE fixed-order block reduction at the end.
// Templated on the padded feature count: no predicated work on absent dimensions.
struct GradArgs {
    const double* x_pad; const double* inv_ell2; const double* alpha; const double* kinv;
    double* gpart; int n, n_pad, dim, nrb;
};

template <int DIMP>
__global__ void __launch_bounds__(256, 2) grad_contract_kernel(GradArgs g) {
    __shared__ __align__(16) double xi[DIMP][64], xj[DIMP][64];
    __shared__ __align__(16) double ai[64], aj[64], ils[DIMP], etab[64];
    __shared__ double red[8][DIMP];
    const int tid = threadIdx.x, dim = g.dim;
    const int bi = blockIdx.x;
    const int64_t r = blockIdx.y;
    const double* kinv = g.kinv + r * (int64_t)g.n_pad * g.n_pad;
    const double* al = g.alpha + r * (int64_t)g.n_pad;
    for (int e = tid; e < 64 * DIMP; e += 256) {
        const int p = e / DIMP, k = e % DIMP;
        xi[k][p] = k < dim ? g.x_pad[(int64_t)(bi * 64 + p) * dim + k] : 0.0;
    }
    if (tid < 64) { ai[tid] = al[bi * 64 + tid]; etab[tid] = kExp2Tab[tid]; }
    if (tid < DIMP) ils[tid] = tid < dim ? g.inv_ell2[r * dim + tid] : 0.0;
    // thread = 8 rows (its warp's) x 2 adjacent columns (its lane's): K^-1 is read as one double2 per row and thread
    // (a warp reads 512 contiguous bytes), column coordinates as conflict-free double2, row coordinates as broadcasts
    const int lane = tid & 31, wrp = tid >> 5;
    double acc[DIMP];
#pragma unroll
    for (int k = 0; k < DIMP; k++) acc[k] = 0.0;
    for (int bj = 0; bj <= bi; bj++) {
        __syncthreads();                                   // the previous tile is done with xj / aj
        for (int e = tid; e < 64 * DIMP; e += 256) {
            const int p = e / DIMP, k = e % DIMP;
            xj[k][p] = k < dim ? g.x_pad[(int64_t)(bj * 64 + p) * dim + k] : 0.0;
        }
        if (tid < 64) aj[tid] = al[bj * 64 + tid];
        __syncthreads();
        double xj0[DIMP], xj1[DIMP];
#pragma unroll
        for (int k = 0; k < DIMP; k++) {
            const double2 v = *reinterpret_cast<const double2*>(&xj[k][2 * lane]);
            xj0[k] = v.x; xj1[k] = v.y;
        }
        const double2 ajv = *reinterpret_cast<const double2*>(&aj[2 * lane]);
        const int j0 = bj * 64 + 2 * lane;
#pragma unroll 2
        for (int a = 0; a < 8; a++) {
            const int li = wrp * 8 + a, i = bi * 64 + li;
            if (i >= g.n || j0 >= i) continue;                     // rows of real points; columns strictly left of the diagonal
            const double2 kv = *reinterpret_cast<const double2*>(kinv + (int64_t)i * g.n_pad + j0);
            const double aiv = ai[li];
            {
                double s = 0.0, d2[DIMP];
#pragma unroll
                for (int k = 0; k < DIMP; k++) { const double df = xi[k][li] - xj0[k]; d2[k] = df * df; s += d2[k] * ils[k]; }
                const double c = (aiv * ajv.x - kv.x) * exp_nonpos(-0.5 * s, etab);
#pragma unroll
                for (int k = 0; k < DIMP; k++) acc[k] += c * d2[k];
            }
            if (j0 + 1 < i) {
                double s = 0.0, d2[DIMP];
#pragma unroll
                for (int k = 0; k < DIMP; k++) { const double df = xi[k][li] - xj1[k]; d2[k] = df * df; s += d2[k] * ils[k]; }
                const double c = (aiv * ajv.y - kv.y) * exp_nonpos(-0.5 * s, etab);
#pragma unroll
                for (int k = 0; k < DIMP; k++) acc[k] += c * d2[k];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < DIMP; k++) {                               // fixed xor tree over the warp, then the 8 warps in order
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][k] = v;
    }
    __syncthreads();
    if (tid < dim) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; q++) s += red[q][tid];
        g.gpart[(r * g.nrb + bi) * dim + tid] = s;
    }
}

// nlml[r] and grad[r][k] from the pieces; one warp per restart
__global__ void __launch_bounds__(256) lml_finish_kernel(const double* __restrict__ y_pad, const double* __restrict__ alpha,
                                                         const double* __restrict__ logdet, const double* __restrict__ gpart,
                                                         const double* __restrict__ ell, double* nlml, double* grad,
                                                         int n, int n_pad, int dim, int nrb, int64_t R) {
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += y_pad[i] * alpha[r * n_pad + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) nlml[r] = 0.5 * (s + logdet[r] + (double)n * kLog2Pi);
    if (grad && lane < dim) {
        double t = 0.0;
        for (int b = 0; b < nrb; b++) t += gpart[(r * nrb + b) * dim + lane];
        const double l = ell[r * dim + lane];
        grad[r * dim + lane] = -t / (l * l * l);
    }
}

__global__ void pad_copy2_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n, int64_t n_pad, int width) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * width) return;
    dst[i] = (i / width < n) ? src[i] : 0.0;
}

struct LmlLayout { size_t x, y, il, a, w, t, alpha, v, apart, logdet, info, gpart, total; int64_t n_pad; size_t tper; };
static LmlLayout lml_layout(int64_t n, int dim, int64_t R, int want_grad) {
    LmlLayout l{}; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    if (n <= 64) { l.info = take(R * 4); l.total = off; l.n_pad = n; return l; }
    const int64_t np = (n + kDiagNB - 1) / kDiagNB * kDiagNB; l.n_pad = np;
    l.tper = cholesky_scratch_doubles(np);
    l.x = take(np * dim * 8); l.y = take(np * 8); l.il = take(R * dim * 8);
    l.a = take((size_t)R * np * np * 8); l.w = take((size_t)R * np * np * 8); l.t = take((size_t)R * l.tper * 8);
    l.alpha = take(R * np * 8); l.v = take(R * np * 8); l.apart = take((size_t)R * alpha_scratch_doubles(np) * 8);
    l.logdet = take(R * 8); l.info = take(R * 4);
    l.gpart = take(want_grad ? (size_t)R * ((np / 64) * (np / 64 + 1) / 2) * dim * 8 : 8);
    l.total = off;
    return l;
}

}  // namespace bogp

using namespace bogp;

extern "C" size_t bogp_nlml_batched_workspace_bytes(int64_t n, int dim, int64_t r, int want_grad) {
    if (n <= 0 || dim <= 0 || dim > BOGP_MAX_DIM || r <= 0) return 0;
    return lml_layout(n, dim, r, want_grad).total;
}

extern "C" int bogp_nlml_batched(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim, const double* d_ell,
                                 int64_t r, double jitter, double* d_nlml_out, double* d_grad_out, void* d_workspace,
                                 size_t workspace_bytes) {
    if (!ctx || !d_x || !d_y || !d_ell || !d_nlml_out || !d_workspace || n <= 0 || dim <= 0 || dim > BOGP_MAX_DIM || r <= 0) {
        set_error("bogp_nlml_batched: bad argument"); return BOGP_ERR_BAD_ARG;
    }
    NvtxRange nvtx("bogp batched nlml (+ gradient)");
    const LmlLayout l = lml_layout(n, dim, r, d_grad_out != nullptr);
    if (workspace_bytes < l.total) { set_error("bogp_nlml_batched: workspace %zu < %zu bytes", workspace_bytes, l.total); return BOGP_ERR_WORKSPACE; }
    char* base = static_cast<char*>(d_workspace);
    cudaStream_t st = ctx->stream;
    int* info = (int*)(base + l.info);
    BOGP_CUDA_CHECK(cudaMemsetAsync(info, 0, r * 4, st));
    if (n <= 64) {
        static DeviceOnce configured;
        if (configured.need(ctx->device)) {
            BOGP_CUDA_CHECK(cudaFuncSetAttribute(nlml_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem));
        }
        SmallArgs a{d_x, d_y, d_ell, d_nlml_out, d_grad_out, info, (int)n, dim, jitter};
        nlml_small_kernel<<<(unsigned)r, 256, kSmallSmem, st>>>(a);
        BOGP_LAUNCH_CHECK(ctx);
        return BOGP_OK;   // a non-positive-definite restart yields NaN in nlml_out (the reference's np.log(det<=0))
    }
    const int64_t np = l.n_pad;
    double* x_pad = (double*)(base + l.x); double* y_pad = (double*)(base + l.y); double* il = (double*)(base + l.il);
    double* A = (double*)(base + l.a); double* W = (double*)(base + l.w); double* T = (double*)(base + l.t);
    double* alpha = (double*)(base + l.alpha); double* v = (double*)(base + l.v); double* logdet = (double*)(base + l.logdet);
    double* gpart = (double*)(base + l.gpart); double* apart = (double*)(base + l.apart);
    const int64_t mat = np * np;
    pad_copy2_kernel<<<(unsigned)((np * dim + 255) / 256), 256, 0, st>>>(d_x, x_pad, n, np, dim); BOGP_LAUNCH_CHECK(ctx);
    pad_copy2_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(d_y, y_pad, n, np, 1); BOGP_LAUNCH_CHECK(ctx);
    BOGP_CUDA_CHECK(cudaMemsetAsync(W, 0, (size_t)r * mat * 8, st));
    BOGP_CUDA_CHECK(cudaMemsetAsync(logdet, 0, r * 8, st));
    int rc = launch_inv_ell2(ctx, d_ell, il, r * dim); if (rc) return rc;
    // the batch index rides on blockIdx.z (<= 65535 per launch)
    for (int64_t r0 = 0; r0 < r; r0 += 32768) {
        const int rb = (int)((r - r0 < 32768) ? (r - r0) : 32768);
        double* Ab = A + r0 * mat; double* Wb = W + r0 * mat;
        rc = launch_gram(ctx, x_pad, np, n, x_pad, np, n, dim, il + r0 * dim, jitter, Ab, np, true, rb, mat); if (rc) return rc;
        rc = cholesky_blocked(ctx, Ab, np, np, mat, Wb, np, mat, logdet + r0, info + r0, rb, T + r0 * l.tper, (int64_t)l.tper); if (rc) return rc;
        rc = trtri_recursive(ctx, Ab, np, mat, Wb, np, mat, T + r0 * l.tper, (int64_t)l.tper, np, rb, 256); if (rc) return rc;
        rc = launch_alpha(ctx, Wb, np, mat, y_pad, v + r0 * np, alpha + r0 * np, apart + r0 * alpha_scratch_doubles(np), (int)np, rb); if (rc) return rc;
        if (d_grad_out) {
            GemmArgs k{};   // Kinv = W^T W (lower part) into A (L is dead)
            k.A = Wb; k.lda = np; k.strideA = mat; k.B = Wb; k.ldb = np; k.strideB = mat; k.C = Ab; k.ldc = np; k.strideC = mat;
            k.M = (int)np; k.N = (int)np; k.K = (int)np; k.alpha = 1.0; k.accumulate = 0; k.lower_only = 1;
            rc = launch_gemm<128, 128, A_KM, B_KN, K_GE_MAXMN>(ctx, k, rb); if (rc) return rc;
            const int nt = (int)(np / 64);
            GradArgs ga{x_pad, il + r0 * dim, alpha + r0 * np, Ab, gpart + r0 * nt * dim, (int)n, (int)np, dim, nt};
            const dim3 ggrid(nt, rb);
#define BOGP_GRAD(D) grad_contract_kernel<D><<<ggrid, 256, 0, st>>>(ga)
            if (dim <= 2) BOGP_GRAD(2); else if (dim <= 4) BOGP_GRAD(4); else if (dim <= 6) BOGP_GRAD(6); else if (dim <= 8) BOGP_GRAD(8);
            else if (dim <= 10) BOGP_GRAD(10); else if (dim <= 12) BOGP_GRAD(12); else BOGP_GRAD(16);
#undef BOGP_GRAD
            BOGP_LAUNCH_CHECK(ctx);
        }
    }
    lml_finish_kernel<<<(unsigned)((r + 7) / 8), 256, 0, st>>>(y_pad, alpha, logdet, gpart, d_ell, d_nlml_out, d_grad_out,
                                                             (int)n, (int)np, dim, (int)(np / 64), r);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}
