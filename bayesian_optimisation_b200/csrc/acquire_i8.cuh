// Shared device code of the INT8 / tcgen05 acquisition path: the k_* digit builder (one 256-row x 64-candidate
// panel tile), the tcgen05 helpers, and the operand geometry.  Included by acquire_i8.cu (separate panel and
// tensor-core kernels: screening passes, profiling) and acquire_fused.cu (the persistent fused sweep kernel).
#pragma once
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

// ------------------------------------------------------------------------------------------------
// panel: k_*[j, c] as digits + partial posterior means.  grid (ceil(cur/64), n_pad/256), 256 threads.
// Thread = (candidate, 16 consecutive j): 16-byte digit vectors are written per slice; tile layout
// per (candidate tile, k tile of 32): [k chunk of 16][slice q][candidate][16 B]  (N index = 64 q + cand).
// ------------------------------------------------------------------------------------------------
struct CandDescI8 {
    const double* points; const double* axes;
    int len[BOGP_MAX_DIM]; int off[BOGP_MAX_DIM];
    double cross_jitter;
    // Grid sweeps: per-axis kernel-factor tables, ft[toff[k] + j * lenp[k] + g] =
    // exp(-0.5 (axis_k[g] - x_jk)^2 / ell_k^2), built once per sweep by grid_factor_kernel.  A grid entry is then the
    // product k_*[j, c] = (((1 f_0) f_1) ... f_{d-1}) of d table values, always in this order, so its bits do not depend on
    // the tile, chunk or GPU that computes it.  The product over the leading d - tt axes is formed once per row and
    // setting of those axes (a 64-candidate tile sees at most two), the trailing tt factors come from table rows staged
    // in shared memory.  Null: every entry is exp(-0.5 * squared distance) (explicit candidate arrays, long axes).
    const double* ft; int toff[BOGP_MAX_DIM]; int lenp[BOGP_MAX_DIM];
    int tt;     // trailing axes multiplied per entry (the smallest count whose grid points number >= 64: a tile then sees at most two settings of the leading axes)
};
struct PanelI8Args {
    CandDescI8 cand;
    const double* x_pad; const double* inv_ell2; const double* alpha;
    uint8_t* panel; double* mupart;
    int64_t c0, c_end, S;
    int n, n_pad, dim;
    const int* d_count;       // screened sweeps: number of valid candidates of the (compacted) array lives on the device
    const long long* idx_list; // screened grid sweeps: slot -> flat grid index of the compacted survivors (table mode, scattered candidates)
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

// Value sources of panel_rows: k_*(x_jl, candidate of this thread) for row jl of the 256-row block.
template <int DIMP, int T>
struct ExpSrc {                      // exp of the squared distance; prefix sums of the shared leading dimensions in `pre`
    const double (&pc)[DIMP]; const double (&il)[DIMP]; const double* xs; const double* etab; const double* pre;
    __device__ __forceinline__ double eval(int jl) const { return kstar_row<DIMP, T>(T < DIMP ? pre[jl] : 0.0, pc, il, xs + jl * DIMP, etab); }
};
template <int TT>
struct TabSrc {                      // prefix product of the leading axes (per row) x TT staged table rows at this candidate's grid points
    const double* pre; const double* st; int dig[TT]; int row0;      // st: [t][64 rows][16]; row0 = first row of the staged piece
    __device__ __forceinline__ double eval(int jl) const {
        double v = pre[jl];
#pragma unroll
        for (int t = 0; t < TT; t++) v *= st[(t * 64 + (jl - row0)) * 16 + dig[t]];
        return v;
    }
};

// the row groups of one panel tile for one thread (= one candidate): digits + partial posterior mean
struct PanelRowCtx {
    const double* al;                 // shared memory
    uint8_t* panel; int64_t tile_base;    // (ct * (n_pad / 32) + jb * 8) * kI8BTile
    int nvr, jq, nl, tid; double jit;
};

// MUONLY: the screening pass of an arg-max-only sweep -- the same k_* values and the same partial posterior means
// (same operations, same order: bit-identical mu), but no digits are formed or stored.
template <bool UB, bool MUONLY, class SRC>
__device__ __forceinline__ double panel_group(const PanelRowCtx& c, const SRC& src, int g) {
    {
        uint32_t pk[kI8Slices][4];
        double mug = 0.0;
        if (MUONLY) {
#pragma unroll 4
            for (int e = 0; e < 16; e++) {
                const int jl = g * 16 + e;
                double v = src.eval(jl);
                v = jl < c.nvr ? v : 0.0;
                if (jl == c.jq) v += c.jit;
                mug += c.al[jl] * v;
            }
            return mug;
        }
        if (UB) {
#pragma unroll
            for (int e4 = 0; e4 < 4; e4++) {
                uint32_t lo[4], hi[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int jl = g * 16 + e4 * 4 + i;
                    double v = src.eval(jl);
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
                double v = src.eval(jl);
                v = jl < c.nvr ? v : 0.0;
                if (jl == c.jq) v += c.jit;
                mug += c.al[jl] * v;
                int d[kI8Slices];
                balanced_digits(__double2ll_rn(v * 18014398509481984.0), d);   // t = v / 2, fx = t * 2^55 = v * 2^54
#pragma unroll
                for (int m = 0; m < kI8Slices; m++) pk[kI8Slices - 1 - m][e >> 2] |= (uint32_t)(d[m] & 0xFF) << (8 * (e & 3));
            }
        }
        // k tile (32 rows) = jb*8 + g/2, k chunk = g & 1
        uint8_t* dst = c.panel + (c.tile_base + (g >> 1)) * kI8BTile + (g & 1) * (kI8Slices * kI8BN * 16) + c.nl * 16;
#pragma unroll
        for (int q = 0; q < kI8Slices; q++)
            *reinterpret_cast<uint4*>(dst + q * (kI8BN * 16)) = make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
        return mug;
    }
}

// 16 groups of 16 rows; this thread takes groups g = tid/64, +4, +8, +12, ascending
template <bool UB, bool MUONLY, class SRC>
__device__ __forceinline__ double panel_rows(const PanelRowCtx& c, const SRC& src) {
    double mu = 0.0;
    for (int g = c.tid >> 6; g < kAcqBM / 16; g += 4) mu += panel_group<UB, MUONLY>(c, src, g);
    return mu;
}

static __device__ const double kOneTab[1] = {1.0};

template <int DIMP>
struct PanelSmem {
    alignas(128) double ps_raw[kI8BN * BOGP_MAX_DIM];   // candidate block, [cand][dim] as in HBM
    union {
        alignas(16) double xs[kAcqBM][DIMP];                // exp mode: the 256 measured points of the row block
        alignas(16) double stage[3 * 64 * 16];              // table mode: 64 rows of up to three trailing axes' tables
    };
    double al[kAcqBM];
    double sl[BOGP_MAX_DIM];
    double etab[64];
    double mured[4][kI8BN];
    double pre[2][kAcqBM];                                  // exp mode: [0] prefix sums; table mode: prefix products of the two leading-axis settings
    int kshare;
    alignas(8) uint64_t bar;                            // bulk-TMA staging of explicit candidate blocks (caller initialises, count 1)
};

// barrier among the 256 threads that build a panel tile: the whole CTA (stand-alone kernel) or the builder
// warps of the fused kernel (named barrier 2)
template <bool NAMED>
__device__ __forceinline__ void panel_sync() {
    if (NAMED) asm volatile("bar.sync 2, 256;" ::: "memory"); else __syncthreads();
}

// Table mode of panel_tile (grid sweeps): called after the grid digits of the tile's candidates are in shared memory.
// Rows go in four pieces of 64; the table rows of the trailing axes of a piece (64 rows x 16 grid points x tt axes, one
// contiguous 8 KB block per axis) are staged in shared memory -- the next piece travels to registers while the current one
// is worked on -- so an entry costs tt shared-memory loads and multiplications instead of a squared distance and an exp.
template <int DIMP, bool UB, bool MUONLY, bool NAMED, int TT>
__device__ __forceinline__ double panel_tile_table_rows(const PanelI8Args& p, const PanelRowCtx& rc, int jb, PanelSmem<DIMP>& sm, int tid,
                                                        const double* pre, const int* mydig) {
    const int kl = p.dim - TT;
    TabSrc<TT> src; src.pre = pre; src.st = sm.stage;
#pragma unroll
    for (int t = 0; t < TT; t++) src.dig[t] = mydig[kl + t];
    // piece r of axis t: 64 rows x 16 doubles = 512 x 16 B; thread tid moves vectors tid and tid + 256
    uint4 nx[TT][2];
    auto fetch = [&](int r) {
#pragma unroll
        for (int t = 0; t < TT; t++) {
            const uint4* src4 = reinterpret_cast<const uint4*>(p.cand.ft + p.cand.toff[kl + t] + ((int64_t)jb * kAcqBM + r * 64) * 16);
            nx[t][0] = __ldg(src4 + tid); nx[t][1] = __ldg(src4 + tid + 256);
        }
    };
    fetch(0);
    double mu = 0.0;
    for (int r = 0; r < 4; r++) {
        uint4* st4 = reinterpret_cast<uint4*>(sm.stage);
#pragma unroll
        for (int t = 0; t < TT; t++) { st4[t * 512 + tid] = nx[t][0]; st4[t * 512 + tid + 256] = nx[t][1]; }
        panel_sync<NAMED>();
        if (r < 3) fetch(r + 1);
        src.row0 = r * 64;
        mu += panel_group<UB, MUONLY>(rc, src, 4 * r + (tid >> 6));
        panel_sync<NAMED>();                       // everybody is done with the staged piece
    }
    return mu;
}

template <int DIMP, bool UB, bool MUONLY, bool NAMED>
__device__ __forceinline__ bool panel_tile_table(const PanelI8Args& p, int64_t cbase, int64_t ct_store, int jb, int nvalid,
                                                 PanelSmem<DIMP>& sm, int tid) {
    const int dim = p.dim, tt = p.cand.tt, kl = dim - tt;
    const int* dg = reinterpret_cast<const int*>(sm.ps_raw);
    const int last = nvalid - 1;
    {   // prefix products over the leading axes, (((1 f_0) f_1) ... f_{kl-1}), for the settings of the first and the last candidate
        const int64_t j = (int64_t)jb * kAcqBM + tid;
        double f0[DIMP], f1[DIMP];
#pragma unroll
        for (int k = 0; k < DIMP; k++) {           // all loads in flight together
            const bool lead = k < kl;
            const double* row = p.cand.ft + (lead ? p.cand.toff[k] + j * p.cand.lenp[k] : 0);
            f0[k] = lead ? __ldg(row + dg[k]) : 1.0; f1[k] = lead ? __ldg(row + dg[last * BOGP_MAX_DIM + k]) : 1.0;
        }
        double v0 = 1.0, v1 = 1.0;
#pragma unroll
        for (int k = 0; k < DIMP; k++) { v0 *= f0[k]; v1 *= f1[k]; }      // x 1.0 is exact
        sm.pre[0][tid] = v0; sm.pre[1][tid] = v1;
    }
    const int nl = tid & 63;
    const int ncl = nl < nvalid ? nl : last;
    const int* mydig = dg + ncl * BOGP_MAX_DIM;
    bool second = false;                           // this candidate sits on the last candidate's setting of the leading axes
    for (int k = 0; k < kl; k++) second |= mydig[k] != dg[k];
    PanelRowCtx rc;
    rc.al = sm.al;
    rc.panel = p.panel; rc.tile_base = ct_store * (p.n_pad / kI8KB) + (int64_t)jb * (kAcqBM / kI8KB);
    rc.nvr = p.n - jb * kAcqBM;
    const int64_t jq64 = (cbase + nl) - (int64_t)jb * kAcqBM;
    rc.jq = (p.cand.cross_jitter != 0.0 && jq64 >= 0 && jq64 < kAcqBM) ? (int)jq64 : -1;
    rc.jit = p.cand.cross_jitter; rc.nl = nl; rc.tid = tid;
    const double* pre = sm.pre[second ? 1 : 0];
    double mu;
    if (tt == 1)      mu = panel_tile_table_rows<DIMP, UB, MUONLY, NAMED, 1>(p, rc, jb, sm, tid, pre, mydig);
    else if (tt == 2) mu = panel_tile_table_rows<DIMP, UB, MUONLY, NAMED, 2>(p, rc, jb, sm, tid, pre, mydig);
    else              mu = panel_tile_table_rows<DIMP, UB, MUONLY, NAMED, 3>(p, rc, jb, sm, tid, pre, mydig);
    sm.mured[tid >> 6][nl] = mu;
    panel_sync<NAMED>();
    if (tid < kI8BN) {                              // fixed-order sum of the 4 thread groups
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) s += sm.mured[g][tid];
        p.mupart[(int64_t)jb * p.S + ct_store * kI8BN + tid] = s;
    }
    return true;
}

// Table mode for candidates scattered over the grid (the compacted survivors of a screened sweep): the same ordered product
// (((1 f_0) f_1) ... f_{d-1}) with every factor read from the tables in L2 -- the same bits as panel_tile_table gives the
// candidate inside a contiguous sweep, at the price of d dependent-latency loads per entry (survivors are few).
template <int DIMP>
struct TabSrcAll {
    const double* tp[DIMP]; int stride[DIMP];      // axes beyond the real dimension point at a constant 1.0 (stride 0)
    __device__ __forceinline__ double eval(int jl) const {
        double f[DIMP];
#pragma unroll
        for (int k = 0; k < DIMP; k++) f[k] = __ldg(tp[k] + (int64_t)jl * stride[k]);      // all loads in flight together
        double v = 1.0;
#pragma unroll
        for (int k = 0; k < DIMP; k++) v *= f[k];                                          // x 1.0 is exact
        return v;
    }
};
template <int DIMP, bool UB, bool MUONLY, bool NAMED>
__device__ __forceinline__ bool panel_tile_scattered(const PanelI8Args& p, int64_t cbase, int64_t ct_store, int jb, int nvalid,
                                                     PanelSmem<DIMP>& sm, int tid) {
    const int* dg = reinterpret_cast<const int*>(sm.ps_raw);
    const int nl = tid & 63;
    const int ncl = nl < nvalid ? nl : nvalid - 1;
    TabSrcAll<DIMP> src;
#pragma unroll
    for (int k = 0; k < DIMP; k++) {
        const bool real = k < p.dim;
        src.stride[k] = real ? p.cand.lenp[k] : 0;
        src.tp[k] = real ? p.cand.ft + p.cand.toff[k] + (int64_t)jb * kAcqBM * p.cand.lenp[k] + dg[ncl * BOGP_MAX_DIM + k] : kOneTab;
    }
    PanelRowCtx rc;
    rc.al = sm.al;
    rc.panel = p.panel; rc.tile_base = ct_store * (p.n_pad / kI8KB) + (int64_t)jb * (kAcqBM / kI8KB);
    rc.nvr = p.n - jb * kAcqBM;
    rc.jq = -1; rc.jit = 0.0; rc.nl = nl; rc.tid = tid;
    sm.mured[tid >> 6][nl] = panel_rows<UB, MUONLY>(rc, src);
    panel_sync<NAMED>();
    if (tid < kI8BN) {                              // fixed-order sum of the 4 thread groups
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) s += sm.mured[g][tid];
        p.mupart[(int64_t)jb * p.S + ct_store * kI8BN + tid] = s;
    }
    return true;
}

// One panel tile: candidates [c0 + 64 ct, +64) x rows [256 jb, +256): digits into tile slot `ct_store` of p.panel,
// partial posterior means into p.mupart[jb * S + 64 ct_store + e].  256 threads (tid 0..255).  `tma_phase` is the
// running parity of sm.bar.  Returns false when the tile lies beyond the (device-side) candidate count.
template <int DIMP, bool UB, bool MUONLY, bool NAMED>
__device__ __forceinline__ bool panel_tile(const PanelI8Args& p, int64_t ct, int64_t ct_store, int jb, PanelSmem<DIMP>& sm, int tid,
                                           uint32_t& tma_phase) {
    static_assert(DIMP % 2 == 0, "rows of the x block are read as double2");
    const int64_t cbase = p.c0 + ct * kI8BN;
    const int dim = p.dim;
    const bool explicit_mode = p.cand.points != nullptr;
    int64_t c_end = p.c_end;
    if (p.d_count) { const int64_t dc = *p.d_count; c_end = dc < c_end ? dc : c_end; }
    const int64_t remain = c_end - cbase;
    if (remain <= 0) return false;                 // tile beyond the device-side count (uniform for the 256 threads)
    const int nvalid = remain >= kI8BN ? kI8BN : (int)remain;
    bool used_tma = false;
    if (explicit_mode) {
        const double* src = p.cand.points + cbase * dim;
        const uint32_t bytes = (uint32_t)nvalid * dim * 8;
        if ((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // bulk-TMA staging of the candidate block
            used_tma = true;
            if (tid == 0) { mbar_expect_tx(&sm.bar, bytes); bulk_g2s(sm.ps_raw, src, bytes, &sm.bar); }
        }
    }
    const bool table = !explicit_mode && p.cand.ft != nullptr;
    if (!table) {
        for (int i = tid; i < kAcqBM * DIMP; i += 256) {
            int r = i / DIMP, k = i % DIMP;
            sm.xs[r][k] = k < dim ? p.x_pad[((int64_t)jb * kAcqBM + r) * dim + k] : 0.0;
        }
    }
    sm.al[tid] = p.alpha[jb * kAcqBM + tid];
    if (tid < BOGP_MAX_DIM) sm.sl[tid] = tid < dim ? p.inv_ell2[tid] : 0.0;
    if (tid < 64) sm.etab[tid] = kExp2Tab[tid];
    if (explicit_mode) {
        if (used_tma) { mbar_wait(&sm.bar, tma_phase); tma_phase ^= 1u; }
        else for (int i = tid; i < nvalid * dim; i += 256) sm.ps_raw[i] = p.cand.points[cbase * dim + i];
    } else if (tid < nvalid) {
        int64_t f = p.idx_list ? p.idx_list[cbase + tid] : cbase + tid;
        int* dg = reinterpret_cast<int*>(sm.ps_raw) + tid * BOGP_MAX_DIM;     // table mode: grid digits instead of coordinates
        for (int k = dim - 1; k >= 0; k--) {
            const int64_t q = f / p.cand.len[k];
            const int g = (int)(f - q * p.cand.len[k]);
            if (table) dg[k] = g; else sm.ps_raw[tid * dim + k] = p.cand.axes[p.cand.off[k] + g];
            f = q;
        }
    }
    if (tid == 0) sm.kshare = DIMP;
    panel_sync<NAMED>();
    if (table && p.idx_list) return panel_tile_scattered<DIMP, UB, MUONLY, NAMED>(p, cbase, ct_store, jb, nvalid, sm, tid);
    if (table) return panel_tile_table<DIMP, UB, MUONLY, NAMED>(p, cbase, ct_store, jb, nvalid, sm, tid);

    // Leading coordinates that ALL candidates of the tile share (a grid sweep: all but the last two or three axes):
    // their part of the squared distance is computed once per row instead of once per entry.  Detected at run time on
    // the staged candidate block, so explicit candidate arrays that happen to be grid-ordered benefit as well.
    if (tid > 0 && tid < nvalid) {
        int k = 0;
        for (; k < dim; k++)
            if (__double_as_longlong(sm.ps_raw[tid * dim + k]) != __double_as_longlong(sm.ps_raw[k])) break;
        if (k < dim) atomicMin(&sm.kshare, k);
    }
    panel_sync<NAMED>();
    const int trail = DIMP - sm.kshare;            // dimensions that differ inside the tile (padding dimensions count as shared only when leading)
    const int T = (DIMP > 4 && trail <= 2) ? 2 : ((DIMP > 4 && trail == 3) ? 3 : ((DIMP > 4 && trail == 4) ? 4 : DIMP));
    if (T < DIMP) {                                // row prefixes over the first DIMP - T dimensions, candidate 0's coordinates
        double s0 = 0.0;
        for (int k = 0; k < DIMP - T; k++) {
            const double d0 = (k < dim ? sm.ps_raw[k] : 0.0) - sm.xs[tid][k];
            s0 += (d0 * d0) * sm.sl[k];
        }
        sm.pre[0][tid] = s0;
    }
    panel_sync<NAMED>();

    const int nl = tid & 63;                       // candidate within the tile
    const int ncl = nl < nvalid ? nl : nvalid - 1;
    double pc[DIMP], il[DIMP];
#pragma unroll
    for (int k = 0; k < DIMP; k++) { pc[k] = k < dim ? sm.ps_raw[ncl * dim + k] : 0.0; il[k] = sm.sl[k]; }
    PanelRowCtx rc;
    rc.al = sm.al;
    rc.panel = p.panel; rc.tile_base = ct_store * (p.n_pad / kI8KB) + (int64_t)jb * (kAcqBM / kI8KB);
    rc.nvr = p.n - jb * kAcqBM;                    // rows of this block that are real measurements (the rest is padding: k_* = 0)
    // row of this block on which the reference's shape-equality jitter falls for this candidate (-1: none)
    const int64_t jq64 = (cbase + nl) - (int64_t)jb * kAcqBM;
    rc.jq = (p.cand.cross_jitter != 0.0 && jq64 >= 0 && jq64 < kAcqBM) ? (int)jq64 : -1;
    rc.jit = p.cand.cross_jitter; rc.nl = nl; rc.tid = tid;
    double mu;
    const double* xs0 = &sm.xs[0][0];
    if (DIMP > 4 && T == 2)      mu = panel_rows<UB, MUONLY>(rc, ExpSrc<DIMP, (DIMP > 4 ? 2 : DIMP)>{pc, il, xs0, sm.etab, sm.pre[0]});
    else if (DIMP > 4 && T == 3) mu = panel_rows<UB, MUONLY>(rc, ExpSrc<DIMP, (DIMP > 4 ? 3 : DIMP)>{pc, il, xs0, sm.etab, sm.pre[0]});
    else if (DIMP > 4 && T == 4) mu = panel_rows<UB, MUONLY>(rc, ExpSrc<DIMP, (DIMP > 4 ? 4 : DIMP)>{pc, il, xs0, sm.etab, sm.pre[0]});
    else                         mu = panel_rows<UB, MUONLY>(rc, ExpSrc<DIMP, DIMP>{pc, il, xs0, sm.etab, sm.pre[0]});
    sm.mured[tid >> 6][nl] = mu;
    panel_sync<NAMED>();
    if (tid < kI8BN) {                              // fixed-order sum of the 4 thread groups
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++) s += sm.mured[g][tid];
        p.mupart[(int64_t)jb * p.S + ct_store * kI8BN + tid] = s;
    }
    return true;
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

// per-axis kernel-factor tables of a grid sweep; one thread per (axis, row, padded grid point)
struct GridTabArgs {
    const double* axes; const double* x_pad; const double* inv_ell2; double* ft;
    int len[BOGP_MAX_DIM]; int off[BOGP_MAX_DIM]; int toff[BOGP_MAX_DIM]; int lenp[BOGP_MAX_DIM]; int toffT[BOGP_MAX_DIM];
    int dim, n_pad;
};
// PanelI8Args.cand from a chunk descriptor
__host__ inline void fill_cand(CandDescI8& c, const AcqChunk& a) {
    c.points = a.points; c.axes = a.axes; c.cross_jitter = a.cross_jitter; c.ft = a.points ? nullptr : a.ft; c.tt = a.tt;
    for (int k = 0; k < BOGP_MAX_DIM; k++) { c.len[k] = a.len[k]; c.off[k] = a.off[k]; c.toff[k] = a.toff[k]; c.lenp[k] = a.lenp[k]; }
}

constexpr int kI8Threads = 320;      // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue

}  // namespace bogp
