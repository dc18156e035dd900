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
#include "common.cuh"
#include "fit.cuh"

namespace bogp {

constexpr int kI8Slices   = 7;
constexpr int kI8BM       = 128;                 // rows of W per CTA (UMMA M)
constexpr int kI8BN       = 64;                  // candidates per CTA
constexpr int kI8KB       = 32;                  // k per stage (one UMMA K for 8-bit operands)
constexpr int kI8ATile    = kI8Slices * kI8BM * kI8KB;   // 28672 B
constexpr int kI8BTile    = kI8Slices * kI8BN * kI8KB;   // 14336 B
constexpr int kI8Stage    = kI8ATile + kI8BTile;         // 43008 B
constexpr int kI8Stages   = 4;
constexpr size_t kI8Smem  = (size_t)kI8Stages * kI8Stage + 256 + 4 * kI8BN * 8;

// ------------------------------------------------------------------------------------------------
// digit extraction: fx = sum_m d_m 256^m with d_m in [-128,127]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void balanced_digits(long long fx, int (&d)[kI8Slices]) {
#pragma unroll
    for (int m = 0; m < kI8Slices; m++) {
        const int v = (int)(signed char)(fx & 0xFF);
        d[m] = v;
        fx = (fx - v) >> 8;
    }
}

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

// ------------------------------------------------------------------------------------------------
// panel: k_*[j, c] as digits + partial posterior means.  grid (ceil(cur/64), n_pad/256), 256 threads.
// Thread = (candidate, 16 consecutive j): 16-byte digit vectors are written per slice; tile layout
// per (candidate tile, k tile of 32): [k chunk of 16][slice q][candidate][16 B]  (N index = 64 q + cand).
// ------------------------------------------------------------------------------------------------
struct CandDescI8 {
    const double* points; const double* axes;
    int len[BOGP_MAX_DIM]; int off[BOGP_MAX_DIM];
    double cross_jitter;
};
struct PanelI8Args {
    CandDescI8 cand;
    const double* x_pad; const double* inv_ell2; const double* alpha;
    uint8_t* panel; double* mupart;
    int64_t c0, c_end, S;
    int n, n_pad, dim;
    const int* d_count;       // screened sweeps: number of valid candidates of the (compacted) array lives on the device
};

// 4x4 byte transpose: out[m] = (byte m of w0, byte m of w1, byte m of w2, byte m of w3)
__device__ __forceinline__ void transpose4x4_bytes(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t (&out)[4]) {
    const uint32_t a = __byte_perm(w0, w1, 0x5140);   // w0.b0 w1.b0 w0.b1 w1.b1
    const uint32_t b = __byte_perm(w2, w3, 0x5140);
    const uint32_t c = __byte_perm(w0, w1, 0x7362);   // w0.b2 w1.b2 w0.b3 w1.b3
    const uint32_t d = __byte_perm(w2, w3, 0x7362);
    out[0] = __byte_perm(a, b, 0x5410);
    out[1] = __byte_perm(a, b, 0x7632);
    out[2] = __byte_perm(c, d, 0x5410);
    out[3] = __byte_perm(c, d, 0x7632);
}

// UB = true : k >= 0, so the 7 low bytes of the fixed-point value ARE its base-256 digits (unsigned
//             operand for the MMA; no digit arithmetic at all, just byte transposes).  Allowed while
//             7 * n_pad * 128 * 255 < 2^31, i.e. n_pad <= 8192.
// UB = false: balanced signed digits (|digit| <= 128), safe up to n_pad = 16384.
// DIMP = feature count rounded up to an instantiated size; the extra dimensions carry zeros
// (coordinate 0, 1/ell^2 = 0) and add exactly +0 to the squared distance.
// k_*(x_j, p) for one row of the shared x block ([row][DIMP], read with 16-byte broadcast loads); same operation
// order as every other kernel-function site (Gram, FP64 panel, LML): s += ((p_k - x_k)^2) / ell_k^2, k ascending.
// T = number of trailing dimensions summed here; the leading DIMP - T dimensions (coordinates shared by all candidates
// of the tile) arrive as the partial sum `s` of the same operations in the same order, so the result is bit-identical
// to T = DIMP with s = 0.
template <int DIMP, int T>
__device__ __forceinline__ double kstar_row(double s, const double (&pc)[DIMP], const double (&il)[DIMP], const double* xrow, const double* etab) {
    constexpr int K0 = DIMP - T;
    if (K0 & 1) { const double d = pc[K0] - xrow[K0]; s += (d * d) * il[K0]; }
    constexpr int K1 = K0 + (K0 & 1);
    const double2* xr = reinterpret_cast<const double2*>(xrow);
#pragma unroll
    for (int k2 = K1 / 2; k2 < DIMP / 2; k2++) {
        const double2 xv = xr[k2];
        const double d0 = pc[2 * k2] - xv.x;     s += (d0 * d0) * il[2 * k2];
        const double d1 = pc[2 * k2 + 1] - xv.y; s += (d1 * d1) * il[2 * k2 + 1];
    }
    return exp_nonpos(-0.5 * s, etab);
}

// the row groups of one panel tile for one thread (= one candidate): digits + partial posterior mean
struct PanelRowCtx {
    const double* xs; const double* al; const double* etab; const double* pre;     // shared memory
    uint8_t* panel; int64_t tile_base;    // (ct * (n_pad / 32) + jb * 8) * kI8BTile
    int nvr, jq, nl, tid; double jit;
};

// MUONLY: the screening pass of an arg-max-only sweep -- the same k_* values and the same partial posterior means
// (same operations, same order: bit-identical mu), but no digits are formed or stored.
template <int DIMP, bool UB, int T, bool MUONLY>
__device__ __forceinline__ double panel_rows(const PanelRowCtx& c, const double (&pc)[DIMP], const double (&il)[DIMP]) {
    double mu = 0.0;      // this thread's 4 row groups, ascending
    // 16 groups of 16 rows; this thread takes groups g = tid/64, +4, +8, +12
    for (int g = c.tid >> 6; g < kAcqBM / 16; g += 4) {
        uint32_t pk[kI8Slices][4];
        double mug = 0.0;
        if (MUONLY) {
#pragma unroll 4
            for (int e = 0; e < 16; e++) {
                const int jl = g * 16 + e;
                double v = kstar_row<DIMP, T>(T < DIMP ? c.pre[jl] : 0.0, pc, il, c.xs + jl * DIMP, c.etab);
                v = jl < c.nvr ? v : 0.0;
                if (jl == c.jq) v += c.jit;
                mug += c.al[jl] * v;
            }
            mu += mug;
            continue;
        }
        if (UB) {
#pragma unroll
            for (int e4 = 0; e4 < 4; e4++) {
                uint32_t lo[4], hi[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int jl = g * 16 + e4 * 4 + i;
                    double v = kstar_row<DIMP, T>(T < DIMP ? c.pre[jl] : 0.0, pc, il, c.xs + jl * DIMP, c.etab);
                    v = jl < c.nvr ? v : 0.0;
                    if (jl == c.jq) v += c.jit;
                    mug += c.al[jl] * v;
                    const unsigned long long fx = __double2ull_rn(v * 18014398509481984.0);   // t = v / 2, fx = t * 2^55 = v * 2^54 (exact scaling)
                    lo[i] = (uint32_t)fx; hi[i] = (uint32_t)(fx >> 32);
                }
                uint32_t tl[4], th[4];
                transpose4x4_bytes(lo[0], lo[1], lo[2], lo[3], tl);      // digits 0..3 (least significant first)
                transpose4x4_bytes(hi[0], hi[1], hi[2], hi[3], th);      // digits 4..6
                pk[6][e4] = tl[0]; pk[5][e4] = tl[1]; pk[4][e4] = tl[2]; pk[3][e4] = tl[3];
                pk[2][e4] = th[0]; pk[1][e4] = th[1]; pk[0][e4] = th[2];
            }
        } else {
#pragma unroll
            for (int q = 0; q < kI8Slices; q++) pk[q][0] = pk[q][1] = pk[q][2] = pk[q][3] = 0u;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int jl = g * 16 + e;
                double v = kstar_row<DIMP, T>(T < DIMP ? c.pre[jl] : 0.0, pc, il, c.xs + jl * DIMP, c.etab);
                v = jl < c.nvr ? v : 0.0;
                if (jl == c.jq) v += c.jit;
                mug += c.al[jl] * v;
                int d[kI8Slices];
                balanced_digits(__double2ll_rn(v * 18014398509481984.0), d);   // t = v / 2, fx = t * 2^55 = v * 2^54
#pragma unroll
                for (int m = 0; m < kI8Slices; m++) pk[kI8Slices - 1 - m][e >> 2] |= (uint32_t)(d[m] & 0xFF) << (8 * (e & 3));
            }
        }
        mu += mug;
        // k tile (32 rows) = jb*8 + g/2, k chunk = g & 1
        uint8_t* dst = c.panel + (c.tile_base + (g >> 1)) * kI8BTile + (g & 1) * (kI8Slices * kI8BN * 16) + c.nl * 16;
#pragma unroll
        for (int q = 0; q < kI8Slices; q++)
            *reinterpret_cast<uint4*>(dst + q * (kI8BN * 16)) = make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
    }
    return mu;
}

template <int DIMP, bool UB, bool MUONLY>
__global__ void __launch_bounds__(256, 3) panel_i8_kernel(PanelI8Args p) {
    static_assert(DIMP % 2 == 0, "rows of the x block are read as double2");
    __shared__ __align__(128) double ps_raw[kI8BN * BOGP_MAX_DIM];
    __shared__ __align__(16) double xs[kAcqBM][DIMP];
    __shared__ double al[kAcqBM];
    __shared__ double sl[BOGP_MAX_DIM];
    __shared__ double etab[64];
    __shared__ double mured[4][kI8BN];
    __shared__ double pre[kAcqBM];
    __shared__ int kshare_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int ct = blockIdx.x, jb = blockIdx.y;
    const int64_t cbase = p.c0 + (int64_t)ct * kI8BN;
    const int dim = p.dim;
    const bool explicit_mode = p.cand.points != nullptr;
    int64_t c_end = p.c_end;
    if (p.d_count) { const int64_t dc = *p.d_count; c_end = dc < c_end ? dc : c_end; }
    const int64_t remain = c_end - cbase;
    if (remain <= 0) return;                       // tile beyond the device-side count (uniform for the CTA)
    const int nvalid = remain >= kI8BN ? kI8BN : (int)remain;
    bool used_tma = false;
    if (explicit_mode) {
        const double* src = p.cand.points + cbase * dim;
        const uint32_t bytes = (uint32_t)nvalid * dim * 8;
        if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // bulk-TMA staging of the candidate block
            used_tma = true;
            if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
            __syncthreads();
            if (tid == 0) { mbar_expect_tx(&bar, bytes); bulk_g2s(ps_raw, src, bytes, &bar); }
        }
    }
    for (int i = tid; i < kAcqBM * DIMP; i += 256) {
        int r = i / DIMP, k = i % DIMP;
        xs[r][k] = k < dim ? p.x_pad[((int64_t)jb * kAcqBM + r) * dim + k] : 0.0;
    }
    al[tid] = p.alpha[jb * kAcqBM + tid];
    if (tid < BOGP_MAX_DIM) sl[tid] = tid < dim ? p.inv_ell2[tid] : 0.0;
    if (tid < 64) etab[tid] = kExp2Tab[tid];
    if (explicit_mode) {
        if (used_tma) mbar_wait(&bar, 0);
        else for (int i = tid; i < nvalid * dim; i += 256) ps_raw[i] = p.cand.points[cbase * dim + i];
    } else if (tid < nvalid) {
        int64_t f = cbase + tid;
        for (int k = dim - 1; k >= 0; k--) {
            const int64_t q = f / p.cand.len[k];
            ps_raw[tid * dim + k] = p.cand.axes[p.cand.off[k] + (int)(f - q * p.cand.len[k])];
            f = q;
        }
    }
    __syncthreads();

    // Leading coordinates that ALL candidates of the tile share (a grid sweep: all but the last two or three axes):
    // their part of the squared distance is computed once per row instead of once per entry.  Detected at run time on
    // the staged candidate block, so explicit candidate arrays that happen to be grid-ordered benefit as well.
    if (tid == 0) kshare_s = DIMP;
    __syncthreads();
    if (tid > 0 && tid < nvalid) {
        int k = 0;
        for (; k < dim; k++)
            if (__double_as_longlong(ps_raw[tid * dim + k]) != __double_as_longlong(ps_raw[k])) break;
        if (k < dim) atomicMin(&kshare_s, k);
    }
    __syncthreads();
    const int trail = DIMP - kshare_s;             // dimensions that differ inside the tile (padding dimensions count as shared only when leading)
    const int T = (DIMP > 4 && trail <= 2) ? 2 : ((DIMP > 4 && trail == 3) ? 3 : ((DIMP > 4 && trail == 4) ? 4 : DIMP));
    if (T < DIMP) {                                // row prefixes over the first DIMP - T dimensions, candidate 0's coordinates
        double s0 = 0.0;
        for (int k = 0; k < DIMP - T; k++) {
            const double d0 = (k < dim ? ps_raw[k] : 0.0) - xs[tid][k];
            s0 += (d0 * d0) * sl[k];
        }
        pre[tid] = s0;
    }
    __syncthreads();

    const int nl = tid & 63;                       // candidate within the tile
    const int ncl = nl < nvalid ? nl : nvalid - 1;
    double pc[DIMP], il[DIMP];
#pragma unroll
    for (int k = 0; k < DIMP; k++) { pc[k] = k < dim ? ps_raw[ncl * dim + k] : 0.0; il[k] = sl[k]; }
    PanelRowCtx rc;
    rc.xs = &xs[0][0]; rc.al = al; rc.etab = etab; rc.pre = pre;
    rc.panel = p.panel; rc.tile_base = (int64_t)ct * (p.n_pad / kI8KB) + (int64_t)jb * (kAcqBM / kI8KB);
    rc.nvr = p.n - jb * kAcqBM;                    // rows of this block that are real measurements (the rest is padding: k_* = 0)
    // row of this block on which the reference's shape-equality jitter falls for this candidate (-1: none)
    const int64_t jq64 = (cbase + nl) - (int64_t)jb * kAcqBM;
    rc.jq = (p.cand.cross_jitter != 0.0 && jq64 >= 0 && jq64 < kAcqBM) ? (int)jq64 : -1;
    rc.jit = p.cand.cross_jitter; rc.nl = nl; rc.tid = tid;
    double mu;
    if (DIMP > 4 && T == 2)      mu = panel_rows<DIMP, UB, (DIMP > 4 ? 2 : DIMP), MUONLY>(rc, pc, il);
    else if (DIMP > 4 && T == 3) mu = panel_rows<DIMP, UB, (DIMP > 4 ? 3 : DIMP), MUONLY>(rc, pc, il);
    else if (DIMP > 4 && T == 4) mu = panel_rows<DIMP, UB, (DIMP > 4 ? 4 : DIMP), MUONLY>(rc, pc, il);
    else                         mu = panel_rows<DIMP, UB, DIMP, MUONLY>(rc, pc, il);
    mured[tid >> 6][nl] = mu;
    __syncthreads();
    if (tid < kI8BN) {                              // fixed-order sum of the 4 thread groups
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) s += mured[g][tid];
        p.mupart[(int64_t)jb * p.S + (int64_t)ct * kI8BN + tid] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // next 16-byte k chunk
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;      // next group of 8 rows
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    return d;                                              // no swizzle, base offset 0
}
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N, int b_signed) {
    // D = s32, A = s8, B = s8 / u8, both K-major
    return (2u << 4) | (1u << 7) | ((uint32_t)(b_signed ? 1 : 0) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}

struct TriI8Args {
    const uint8_t* wq; const uint8_t* panel; const double* wscale; double* qpart;
    int nI, nct, n_pad; int64_t S; int b_signed; int group;   // group = candidate tiles scheduled together (L2 reuse of the panel)
    const int* d_count; int64_t count_c0;                     // screened sweeps: valid candidates = *d_count - count_c0 (device side)
};

// int32 (held as raw bits) -> double without the conversion unit: 2^52 + 2^31 + r is exact in the
// low mantissa bits, and subtracting the constant is exact.
__device__ __forceinline__ double i32_bits_to_f64(uint32_t r) {
    return __hiloint2double(0x43300000, (int)(r ^ 0x80000000u)) - 4503601774854144.0;
}

constexpr int kI8Threads = 320;      // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue

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
    pa.d_count = a.d_count;
    pa.cand.points = a.points; pa.cand.axes = a.axes; pa.cand.cross_jitter = a.cross_jitter;
    for (int k = 0; k < BOGP_MAX_DIM; k++) { pa.cand.len[k] = a.len[k]; pa.cand.off[k] = a.off[k]; }
    pa.x_pad = a.x_pad; pa.inv_ell2 = a.inv_ell2; pa.alpha = a.alpha; pa.panel = (uint8_t*)a.panel; pa.mupart = a.mupart;
    pa.c0 = a.c0; pa.c_end = a.c_end; pa.S = a.S; pa.n = a.n; pa.n_pad = a.n_pad; pa.dim = a.dim;
    const bool ub = a.n_pad <= 8192;      // unsigned panel digits while the int32 level sums cannot overflow
    const dim3 pgrid(nct, a.n_pad / kAcqBM);
#define BOGP_PANEL_I8(D)                                                                                                 \
    do {                                                                                                                 \
        if (mu_only) { BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_i8_kernel<D, true, true><<<pgrid, 256, 0, stream>>>(pa))); }  \
        else if (ub) { BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_i8_kernel<D, true, false><<<pgrid, 256, 0, stream>>>(pa))); } \
        else         { BOGP_PROFILED(ctx, BOGP_PROF_PANEL, (panel_i8_kernel<D, false, false><<<pgrid, 256, 0, stream>>>(pa))); }\
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
    group = group < 1 ? 1 : (group > 64 ? 64 : group);
    const int ngroups = (nct + group - 1) / group;
    TriI8Args ta{a.wq, (const uint8_t*)a.panel, a.wscale, a.qpart, nI, nct, a.n_pad, a.S, ub ? 0 : 1, group, a.d_count, a.c0};
    BOGP_PROFILED(ctx, BOGP_PROF_TRIGEMM, (trigemm_i8_kernel<<<ngroups * nI * group, kI8Threads, kI8Smem, stream>>>(ta)));
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

}  // namespace bogp
