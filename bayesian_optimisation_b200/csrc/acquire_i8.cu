// K4 on the 5th-generation tensor cores: V = W k_* computed EXACTLY in integer arithmetic
// (Ozaki-style error-free splitting) with tcgen05.mma kind::i8 and TMEM accumulators.
//
// Why: on B200 the FP64 tensor path (DMMA) shares the FP64 pipe and tops out at 37 TFLOP/s
// (profiles/r01_dmma_bench.txt); the DMMA acquisition kernel already sits at 92-97 % of that.
// The int8 tensor pipe is ~120x wider, so fp64-exact products are cheaper as sliced integers:
//
//   W[i,j]  = 2^{e_i}  * sum_p a_p[i,j] 2^{-7-8p}      a_p in [-128,127]  (7 balanced base-256 digits,
//   k[j,c]  = 2^{1}    * sum_q b_q[j,c] 2^{-7-8q}      b_q in [-128,127]   fixed point with 55 fraction bits)
//   V[i,c]  = 2^{e_i+1-14} * sum_t 2^{-8t} S_t[i,c],   S_t = sum_{p+q=t} sum_j a_p[i,j] b_q[j,c]   (int32, exact)
//
// Levels t = 0..7 are kept (34 digit pairs).  The dropped levels t >= 8 amount to at most 5 K 2^-62 of the
// row scale 2^{e_i} in the worst case (2^-47.7 at K = 4096, every digit extreme and aligned) and to about
// 2^-56 for real data, where the level sums are random walks; with the 2^-55 fixed-point rounding of the
// operands the result stays inside the worst-case rounding error of an fp64 dot product of the same length
// (tests/test_digit_slices.py checks the scheme against exact rational arithmetic on the host).
// |S_t| <= 7 * K * 2^14 < 2^31 for K <= 16384, so int32 accumulation never overflows.
// Because the integer sums are exact, V does not depend on any summation order: the result is
// bit-reproducible across tiles, chunks and GPUs by construction.
//
// Kernel structure (one CTA = 128 rows of W x 64 candidates, 320 threads):
//   warp 0   : bulk-TMA producer  (W digit tile 28 KB + panel digit tile 14 KB per K=32 stage, mbarrier ring)
//   warp 1   : tcgen05.mma issuer (11 MMAs per stage: slice p of W against slices 0..min(6,7-p) of the
//              panel concatenated along N, landing on TMEM columns 64(p+q)..: the 8 levels fill all 512 columns)
//   warps 2-9: epilogue -- tcgen05.ld the 8 int32 levels, Horner-combine them in fp64, scale by the row
//              exponent, square and reduce over the 128 rows (transposing warp butterfly + shared memory).
#include "acquire_i8.cuh"

namespace bogp {

// per-row exponent of W: |W[i,j]| * 2^-e_i < 1/2   (one warp per row)
__global__ void __launch_bounds__(256) row_exp_kernel(const double* __restrict__ w, int64_t ldw, int n, int* __restrict__ wexp,
                                                      double* __restrict__ wscale) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    double m = 0.0;
    for (int j = lane; j <= row; j += 32) m = fmax(m, fabs(w[(int64_t)row * ldw + j]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        const int e = (m > 0.0 ? ilogb(m) : 0) + 2;
        wexp[row] = e;
        wscale[row] = ldexp(1.0, e + 1 - 14);      // 2^{e_i} (W) * 2^{1} (panel) * 2^{-7-7}
    }
}

// W (row-major, lower) -> digit tiles.  Tile (ib, kt): rows [128 ib, +128) x k [32 kt, +32), kt < 4 (ib+1);
// layout [slice p][k chunk of 16][row][16 B] so that each slice is a K-major UMMA operand with
// LBO = 2048 B (next k chunk) and SBO = 128 B (next 8 rows), and a whole tile is one bulk copy.
__global__ void __launch_bounds__(256) slice_w_kernel(const double* __restrict__ w, int64_t ldw, const int* __restrict__ wexp,
                                                      uint8_t* __restrict__ wq) {
    const int ib = blockIdx.y, kt = blockIdx.x;
    if (kt >= (ib + 1) * (kI8BM / kI8KB)) return;
    const int64_t tile = (int64_t)ib * (ib + 1) / 2 * (kI8BM / kI8KB) + kt;
    const int r = threadIdx.x >> 1, h = threadIdx.x & 1;
    const int row = ib * kI8BM + r;
    const double* src = w + (int64_t)row * ldw + kt * kI8KB + h * 16;
    const int sh = 55 - wexp[row];
    uint32_t pk[kI8Slices][4];
#pragma unroll
    for (int p = 0; p < kI8Slices; p++) pk[p][0] = pk[p][1] = pk[p][2] = pk[p][3] = 0u;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        int d[kI8Slices];
        balanced_digits(__double2ll_rn(ldexp(src[e], sh)), d);
#pragma unroll
        for (int m = 0; m < kI8Slices; m++) pk[kI8Slices - 1 - m][e >> 2] |= (uint32_t)(d[m] & 0xFF) << (8 * (e & 3));
    }
    uint8_t* dst = wq + tile * kI8ATile + h * (kI8BM * 16) + r * 16;
#pragma unroll
    for (int p = 0; p < kI8Slices; p++)
        *reinterpret_cast<uint4*>(dst + p * (kI8BM * kI8KB)) = make_uint4(pk[p][0], pk[p][1], pk[p][2], pk[p][3]);
}

// Per-axis kernel-factor tables of a grid sweep: ft[toff[k] + j * lenp[k] + g] = exp(-0.5 (axis_k[g] - x_jk)^2 / ell_k^2),
// the operations of one term of every other kernel-function site.  grid (ceil(n_pad * lenp_k / 256), dim).
__global__ void __launch_bounds__(256) grid_factor_kernel(GridTabArgs a) {
    const int k = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)a.n_pad * a.lenp[k]) return;
    const int64_t j = i / a.lenp[k];
    const int g = (int)(i - j * a.lenp[k]);
    double v = 0.0;
    if (g < a.len[k]) {
        const double d = a.axes[a.off[k] + g] - a.x_pad[j * a.dim + k];
        v = exp_nonpos(-0.5 * ((d * d) * a.inv_ell2[k]));
    }
    a.ft[a.toff[k] + i] = v;
    if (g < a.len[k]) a.ft[a.toffT[k] + (int64_t)g * a.n_pad + j] = v;      // transposed copy: rows over j
}

size_t grid_table_reserve(int64_t n_pad) {
    const size_t full = (size_t)n_pad * BOGP_MAX_DIM * 64 * 8;      // tables, both layouts
    return full < ((size_t)64 << 20) ? full : ((size_t)64 << 20);
}

size_t grid_table_geometry(AcqChunk& a) {
    a.ft = nullptr;
    if (a.points || !a.axes) return 0;
    // trailing axes multiplied per entry: the fewest whose grid points number at least one tile (64), so that a tile sees
    // at most two settings of the leading axes; their table rows are staged in shared memory, 16 grid points per row
    int tt = 0; int64_t prod = 1;
    while (tt < a.dim && prod < kI8BN) { tt++; prod *= a.len[a.dim - tt]; }
    if (tt < 1 || tt > 3) return 0;
    for (int t = 0; t < tt; t++) if (a.len[a.dim - 1 - t] > 16) return 0;
    a.tt = tt;
    int64_t off = 0;
    for (int k = 0; k < a.dim; k++) {
        a.lenp[k] = (a.len[k] + 15) / 16 * 16;          // rows of whole 128-byte lines (the trailing axes: exactly one)
        if (off > (int64_t)0x7fffffff) return 0;
        a.toff[k] = (int)off;
        off += (int64_t)a.n_pad * a.lenp[k];
    }
    for (int k = 0; k < a.dim; k++) {                   // transposed copies
        if (off > (int64_t)0x7fffffff) return 0;
        a.toffT[k] = (int)off;
        off += (int64_t)a.n_pad * a.len[k];
    }
    return (size_t)off * 8;
}

int launch_grid_factors(bogp_ctx* ctx, const AcqChunk& a, double* d_ft, cudaStream_t stream) {
    GridTabArgs g{};
    g.axes = a.axes; g.x_pad = a.x_pad; g.inv_ell2 = a.inv_ell2; g.ft = d_ft; g.dim = a.dim; g.n_pad = a.n_pad;
    int maxlen = 0;
    for (int k = 0; k < BOGP_MAX_DIM; k++) { g.len[k] = a.len[k]; g.off[k] = a.off[k]; g.toff[k] = a.toff[k]; g.lenp[k] = a.lenp[k]; g.toffT[k] = a.toffT[k]; if (k < a.dim && a.lenp[k] > maxlen) maxlen = a.lenp[k]; }
    const dim3 grid((unsigned)(((int64_t)a.n_pad * maxlen + 255) / 256), a.dim);
    grid_factor_kernel<<<grid, 256, 0, stream>>>(g);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

// stand-alone panel kernel: grid (ceil(cur/64), n_pad/256), 256 threads, one panel tile per CTA
template <int DIMP, bool UB, bool MUONLY>
__global__ void __launch_bounds__(256, 3) panel_i8_kernel(PanelI8Args p) {
    extern __shared__ __align__(128) unsigned char panel_smem_raw[];
    PanelSmem<DIMP>& sm = *reinterpret_cast<PanelSmem<DIMP>*>(panel_smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t phase = 0;
    panel_tile<DIMP, UB, MUONLY, false>(p, blockIdx.x, blockIdx.x, blockIdx.y, sm, tid, phase);
}

__global__ void __launch_bounds__(kI8Threads, 1) trigemm_i8_kernel(TriI8Args g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full   = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kI8Stages * kI8Stage);
    uint64_t* empt   = full + kI8Stages;
    uint64_t* accbar = empt + kI8Stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accbar + 1);
    double*   red    = reinterpret_cast<double*>(smem_raw + (size_t)kI8Stages * kI8Stage + 256);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTA order: groups of `group` candidate tiles; inside a group the heaviest row blocks first.  CTAs
    // that run together then share the group's panel tiles (and the W row block) in L2.
    const int per_group = g.nI * g.group;
    const int grp = (int)(blockIdx.x / per_group), rem = (int)(blockIdx.x % per_group);
    const int gc = min(g.group, g.nct - grp * g.group);
    const int ib = g.nI - 1 - rem / gc;
    const int ct = grp * g.group + rem % gc;
    if (ib < 0) return;                                           // padding CTAs of the last (partial) group
    if (g.d_count && (int64_t)ct * kI8BN >= (int64_t)*g.d_count - g.count_c0) return;   // candidate tile beyond the device-side count
    const int nk = (ib + 1) * (kI8BM / kI8KB);
    const uint8_t* wsrc = g.wq + (int64_t)ib * (ib + 1) / 2 * (kI8BM / kI8KB) * kI8ATile;
    const uint8_t* psrc = g.panel + (int64_t)ct * (g.n_pad / kI8KB) * kI8BTile;
    const int npre = nk < kI8Stages ? nk : kI8Stages;             // stages filled before anybody has to free one

    if (tid == 0) {
        for (int s = 0; s < kI8Stages; s++) { mbar_init(&full[s], 1); mbar_init(&empt[s], 1); }
        mbar_init(accbar, 1);
        fence_mbar_init();
        // the first loads leave now, so that their latency overlaps the TMEM allocation and the barrier below
        for (int kt = 0; kt < npre; kt++) {
            unsigned char* dst = smem_raw + (size_t)kt * kI8Stage;
            mbar_expect_tx(&full[kt], kI8Stage);
            bulk_g2s(dst, wsrc + (int64_t)kt * kI8ATile, kI8ATile, &full[kt]);
            bulk_g2s(dst + kI8ATile, psrc + (int64_t)kt * kI8BTile, kI8BTile, &full[kt]);
        }
    }
    if (warp == 1) {                                              // TMEM: all 512 columns (8 levels x 64 candidates)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kt = npre; kt < nk; kt++) {
                const int s = kt % kI8Stages;
                mbar_wait(&empt[s], ((kt / kI8Stages) - 1) & 1);
                unsigned char* dst = smem_raw + (size_t)s * kI8Stage;
                mbar_expect_tx(&full[s], kI8Stage);
                bulk_g2s(dst, wsrc + (int64_t)kt * kI8ATile, kI8ATile, &full[s]);
                bulk_g2s(dst + kI8ATile, psrc + (int64_t)kt * kI8BTile, kI8BTile, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kt = 0; kt < nk; kt++) {
                const int s = kt % kI8Stages;
                mbar_wait(&full[s], (kt / kI8Stages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(smem_raw + (size_t)s * kI8Stage);
                const uint32_t b0 = a0 + kI8ATile;
#pragma unroll
                for (int p = 0; p < kI8Slices; p++) {
                    const int nq = (8 - p) < kI8Slices ? (8 - p) : kI8Slices;     // panel slices q = 0..nq-1 -> levels p..p+nq-1
                    const int ntot = nq * kI8BN;
                    const uint64_t da = umma_desc_kmajor(a0 + p * (kI8BM * kI8KB), kI8BM * 16, 128);
#pragma unroll
                    for (int n0 = 0; n0 < ntot; n0 += 256) {
                        const int nn = (ntot - n0) < 256 ? (ntot - n0) : 256;
                        const uint64_t db = umma_desc_kmajor(b0 + n0 * 16, kI8Slices * kI8BN * 16, 128);
                        if (p == 1 && n0 == 256 && kt == 0) {
                            // Levels 0..6 are initialised by the non-accumulating MMAs of slice 0; level 7 is first
                            // written here (slice 1 x panel slice 6), so in the very first stage that part is split off
                            // and does not accumulate.  No TMEM zeroing pass is needed.
                            umma_i8(tmem + (uint32_t)(p * kI8BN + n0), da, db, umma_idesc_i8(kI8BM, nn - kI8BN, g.b_signed), 1u);
                            const uint64_t db7 = umma_desc_kmajor(b0 + (n0 + nn - kI8BN) * 16, kI8Slices * kI8BN * 16, 128);
                            umma_i8(tmem + (uint32_t)(p * kI8BN + n0 + nn - kI8BN), da, db7, umma_idesc_i8(kI8BM, kI8BN, g.b_signed), 0u);
                        } else {
                            umma_i8(tmem + (uint32_t)(p * kI8BN + n0), da, db, umma_idesc_i8(kI8BM, nn, g.b_signed), (p > 0 || kt > 0) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(&empt[s]);                            // stage reusable once these MMAs have read it
            }
            umma_commit(accbar);                                  // all MMAs done -> epilogue
        }
    } else {
        // Epilogue, 8 warps: warp w reads TMEM lane quarter w % 4 (hardware rule) and half (w-2)/4 of the 64 candidates.
        // Two warps per lane quarter keep TMEM loads of one in flight while the other one computes.
        mbar_wait(accbar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const double scale = g.wscale[ib * kI8BM + q4 * 32 + lane];
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        const int csel = (b4 ? 4 : 0) | (b3 ? 2 : 0) | (b2 ? 1 : 0);
        for (int c0 = half * (kI8BN / 2); c0 < (half + 1) * (kI8BN / 2); c0 += 8) {
            uint32_t r[8][8];
#pragma unroll
            for (int t = 0; t < 8; t++) tmem_ld8(tmem + lane_base + t * kI8BN + c0, r[t]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double a = i32_bits_to_f64(r[7][j]);
#pragma unroll
                for (int t = 6; t >= 0; t--) a = fma(a, 0.00390625, i32_bits_to_f64(r[t][j]));      // Horner in 2^-8 (exact products)
                a *= scale;
                v[j] = a * a;
            }
            // sum over the 32 rows of this warp for 8 columns at once: a transposing butterfly (each step halves the
            // number of columns a lane carries), then two plain steps.  Fixed order, same on every tile.
            double w4[4], w2[2], w1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
                w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double keep = b3 ? w4[i + 2] : w4[i], send = b3 ? w4[i] : w4[i + 2];
                w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            {
                const double keep = b2 ? w2[1] : w2[0], send = b2 ? w2[0] : w2[1];
                w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
            w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
            if ((lane & 3) == 0) red[q4 * kI8BN + c0 + csel] = w1;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int e = tid - 64;
        if (e < kI8BN) {
            const double q = ((red[e] + red[kI8BN + e]) + red[2 * kI8BN + e]) + red[3 * kI8BN + e];
            g.qpart[(int64_t)ib * g.S + (int64_t)ct * kI8BN + e] = q;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ------------------------------------------------------------------------------------------------
size_t i8_wq_bytes(int64_t n_pad) {
    const int64_t nI = n_pad / kI8BM;
    return (size_t)(nI * (nI + 1) / 2) * (kI8BM / kI8KB) * kI8ATile;
}

int launch_slice_w(bogp_ctx* ctx, const double* d_w, int64_t n_pad, int* d_wexp, double* d_wscale, uint8_t* d_wq) {
    row_exp_kernel<<<(unsigned)((n_pad + 7) / 8), 256, 0, ctx->stream>>>(d_w, n_pad, (int)n_pad, d_wexp, d_wscale);
    BOGP_LAUNCH_CHECK(ctx);
    const int nI = (int)(n_pad / kI8BM);
    slice_w_kernel<<<dim3(nI * (kI8BM / kI8KB), nI), 256, 0, ctx->stream>>>(d_w, n_pad, d_wexp, d_wq);
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

size_t i8_panel_bytes(int64_t n_pad, int64_t S) { return (size_t)n_pad * S * kI8Slices; }

int launch_panel_i8(bogp_ctx* ctx, const AcqChunk& a, cudaStream_t stream, bool mu_only) {
    const int nct = (int)((a.cur + kI8BN - 1) / kI8BN);
    PanelI8Args pa{};
    pa.d_count = a.d_count; pa.idx_list = a.idx_list;
    fill_cand(pa.cand, a);
    pa.x_pad = a.x_pad; pa.inv_ell2 = a.inv_ell2; pa.alpha = a.alpha; pa.panel = (uint8_t*)a.panel; pa.mupart = a.mupart;
    pa.c0 = a.c0; pa.c_end = a.c_end; pa.S = a.S; pa.n = a.n; pa.n_pad = a.n_pad; pa.dim = a.dim;
    const bool ub = a.n_pad <= 8192;      // unsigned panel digits while the int32 level sums cannot overflow
    const dim3 pgrid(nct, a.n_pad / kAcqBM);
#define BOGP_PANEL_I8(D)                                                                                                 \
    do {                                                                                                                 \
        static DeviceOnce once;                                                                                          \
        const size_t psm = sizeof(PanelSmem<D>);                                                                         \
        if (once.need(ctx->device)) {                                                                                    \
            BOGP_CUDA_CHECK(cudaFuncSetAttribute(panel_i8_kernel<D, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));   \
            BOGP_CUDA_CHECK(cudaFuncSetAttribute(panel_i8_kernel<D, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));  \
            BOGP_CUDA_CHECK(cudaFuncSetAttribute(panel_i8_kernel<D, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm)); \
        }                                                                                                                \
        if (mu_only) { BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_i8_kernel<D, true, true><<<pgrid, 256, psm, stream>>>(pa))); }  \
        else if (ub) { BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_i8_kernel<D, true, false><<<pgrid, 256, psm, stream>>>(pa))); } \
        else         { BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_i8_kernel<D, false, false><<<pgrid, 256, psm, stream>>>(pa))); }\
    } while (0)
    if (a.dim <= 2) BOGP_PANEL_I8(2); else if (a.dim <= 4) BOGP_PANEL_I8(4); else if (a.dim <= 6) BOGP_PANEL_I8(6);
    else if (a.dim <= 8) BOGP_PANEL_I8(8); else if (a.dim <= 10) BOGP_PANEL_I8(10); else if (a.dim <= 12) BOGP_PANEL_I8(12);
    else BOGP_PANEL_I8(16);
#undef BOGP_PANEL_I8
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

int launch_trigemm_i8(bogp_ctx* ctx, const AcqChunk& a, cudaStream_t stream) {
    static DeviceOnce configured;
    if (configured.need(ctx->device)) {
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(trigemm_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kI8Smem));
    }
    const int nct = (int)((a.cur + kI8BN - 1) / kI8BN);
    const bool ub = a.n_pad <= 8192;
    const int nI = a.n_pad / kI8BM;
    int group = (int)((32u << 20) / ((size_t)(a.n_pad / kI8KB) * kI8BTile));     // ~32 MB of panel per group
    if (ctx->fused_group > 0) group = ctx->fused_group;          // bogp_set_fused(ctx, enable, group): the work-group size of either variant
    group = group < 1 ? 1 : (group > 64 ? 64 : group);
    const int ngroups = (nct + group - 1) / group;
    TriI8Args ta{a.wq, (const uint8_t*)a.panel, a.wscale, a.qpart, nI, nct, a.n_pad, a.S, ub ? 0 : 1, group, a.d_count, a.c0};
    BOGP_PROFILED(ctx, BOGP_PROF_TRIGEMM, (trigemm_i8_kernel<<<ngroups * nI * group, kI8Threads, kI8Smem, stream>>>(ta)));
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

}  // namespace bogp
