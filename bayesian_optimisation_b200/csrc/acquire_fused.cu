// K4, fused: ONE persistent kernel per acquisition sweep on the INT8 / tcgen05 path.      point_selector.py:81,90-98,204-207
//
//   flat grid index -> coordinates -> k_* digits -> V = L^-1 k_* on tcgen05 (exact digit slices) -> |V|^2 -> sigma^2, mu
//   -> LCB / EI -> (score, index) max-loc,
//
// all inside one launch: no per-chunk panel / product / finalize / merge kernels, and the k_* panel never exists as a
// whole -- it lives in a ring of a few candidate tiles (a few tens of MB, L2-resident) between the warps that build it
// and the warps that consume it.
//
// One CTA per SM, 18 warps:
//   warp 0      bulk-TMA producer: draws (candidate tile, row block) work items from a global counter, waits until the
//               tile's digits are in the ring, streams W-digit and panel-digit tiles through a 4-stage mbarrier ring
//   warp 1      tcgen05.mma issuer (11 MMAs per K=32 stage into all 512 TMEM columns, see acquire_i8.cu)
//   warps 2-9   epilogue: TMEM -> Horner -> scale -> square -> column sums over the 128 rows -> qpart ring; the CTA that
//               delivers the LAST row block of a candidate tile also finalises that tile: sums the row-block partials
//               in fixed order, sigma^2 = prior - q, mu, score, optional outputs, and folds the tile's winner into the
//               CTA's running (score, index)
//   warps 10-17 panel builders: draw (candidate tile, 256-row block) build items from a second counter, compute the
//               k_* entries (the same device function as the stand-alone panel kernel: bit-identical digits) and the
//               partial posterior means, and publish the tile when its last block is written
//
// The two item streams are decoupled by per-slot counters in global memory (release / acquire at gpu scope):
//   ready[slot]  build items finished        consumers wait for (generation + 1) * nJ
//   done[slot]   row blocks finished         the arrival that completes (generation + 1) * nI finalises the tile
//   freed[slot]  tiles finalised             builders wait for `generation` before overwriting the slot
// Items are drawn in order, every wait is on strictly earlier items, and a CTA that is not resident has drawn nothing:
// the scheme cannot deadlock whatever the residency.  Work items go by groups of G candidate tiles, heaviest row block
// first inside a group (the order of the stand-alone kernel: co-running CTAs share W tiles in L2); the ring holds
// R = 2 G tiles, so the builders fill the next group while the tensor cores work on this one.
//
// Every reduction keeps the order of the separate kernels (integer level sums are order-free, the row-block and
// 256-row partials are summed ascending, the arg-max rule is a total order), so mu, sigma, the scores and the winner
// are bit-identical to the stand-alone kernels -- tests/test_gpu_fused.py compares them with ==.
#include "acquire_i8.cuh"

namespace bogp {

constexpr int kFusedThreads  = 576;
constexpr int kFusedStages   = 4;                                // stages of the operand ring (a power of two)
constexpr int kFusedMiscOff  = kFusedStages * kI8Stage;          // barriers, tile ring, reduction scratch
constexpr int kFusedMiscSize = 2560;
constexpr long long kFusedNoIndex = 0x7fffffffffffffffLL;

struct FusedTile { long long ct; int ib; int pad; };

struct FusedArgs {
    PanelI8Args pa;               // candidates, x, 1/ell^2, alpha; panel = ring base, mupart = ring [jb][R*64]; c0 = sweep begin; S = R*64
    const uint8_t* wq; const double* wscale; double* qpart;      // qpart ring [ib][R*64]
    unsigned long long* ctr64;    // [0] next work item, [1] next build item
    int* ctr32;                   // [0] CTAs finished, [1] NaN flag, [16 + slot] ready, [16 + R + slot] done, [16 + 2R + slot] freed
    int nI, nJ, R, G, b_signed;
    long long nct;                // candidate tiles of the sweep (capacity; a device-side count may cut it)
    int fold_prev;                // fold the record already in `result` into the winner (running best of a screened sweep)
    double* mu_out; double* sigma_out; double* acq_out; const long long* idx_map;
    int kind; double explore, f_best, prior;
    bogp_result* cta_best; bogp_result* result;
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel_gpu(int* p, int v) {
    int o; asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); return o;
}
// candidates of the sweep that exist: the requested range, cut by the device-side count of a compacted array
__device__ __forceinline__ long long fused_valid(const PanelI8Args& p) {
    long long e = p.c_end;
    if (p.d_count) { const long long dc = *p.d_count; e = dc < e ? dc : e; }
    return e - p.c0;
}

template <int DIMP, bool UB>
__global__ void __launch_bounds__(kFusedThreads, 1) acquire_fused_i8_kernel(const __grid_constant__ FusedArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* misc = smem_raw + kFusedMiscOff;
    uint64_t* full       = reinterpret_cast<uint64_t*>(misc);            // [4]
    uint64_t* empt       = full + kFusedStages;                              // [4]
    uint64_t* accbar     = empt + kFusedStages;                              // MMAs of a tile done -> epilogue
    uint64_t* tmem_empty = accbar + 1;                                    // epilogue has drained TMEM -> next tile's MMAs
    uint64_t* tile_full  = tmem_empty + 1;                                // [4] tile descriptor published
    uint32_t* tmem_slot  = reinterpret_cast<uint32_t*>(tile_full + 4);    // +112
    int*      fin_flag   = reinterpret_cast<int*>(misc + 120);
    FusedTile* tile_ring = reinterpret_cast<FusedTile*>(misc + 128);      // [4] x 16 B
    unsigned long long* build_item = reinterpret_cast<unsigned long long*>(misc + 192);
    double*   fin_s      = reinterpret_cast<double*>(misc + 200);         // [2]
    long long* fin_i     = reinterpret_cast<long long*>(misc + 216);      // [2]
    double*   red        = reinterpret_cast<double*>(misc + 256);         // [4][64]
    PanelSmem<DIMP>& psm = *reinterpret_cast<PanelSmem<DIMP>*>(misc + kFusedMiscSize);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* ready = g.ctr32 + 16;
    int* done  = ready + g.R;
    int* freed = done + g.R;
    const long long per_group = (long long)g.nI * g.G;
    long long nct = g.nct;                                       // tiles that exist: all work lists end at the device-side count
    { const long long tv = (fused_valid(g.pa) + kI8BN - 1) / kI8BN; nct = tv < nct ? (tv > 0 ? tv : 0) : nct; }
    const unsigned long long total_items = (unsigned long long)((nct + g.G - 1) / g.G) * (unsigned long long)per_group;

    if (tid == 0) {
        for (int s = 0; s < kFusedStages; s++) { mbar_init(&full[s], 1); mbar_init(&empt[s], 1); mbar_init(&tile_full[s], 1); }
        mbar_init(accbar, 1);
        mbar_init(tmem_empty, 8);
        mbar_init(&psm.bar, 1);
        fence_mbar_init();
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
        // ---------------------------------------------------------------- work scheduler + bulk-TMA producer
        if (lane == 0) {
            const long long nvalid = fused_valid(g.pa);
            long long t = 0; uint32_t it = 0;
            // work item -> (candidate tile, row block); ib = -1: nothing to do (padding of the last group, tile beyond the
            // device-side count), ib = -2: the list is exhausted
            auto decode = [&](unsigned long long w, long long& ct, int& ib) {
                ct = -1; ib = -2;
                if (w >= total_items) return;
                const long long grp = (long long)(w / (unsigned long long)per_group);
                const int rem = (int)(w % (unsigned long long)per_group);
                const long long left = nct - grp * g.G;
                const int gc = left < g.G ? (int)left : g.G;
                ib = g.nI - 1 - rem / gc;
                ct = grp * g.G + rem % gc;
                if (ib < 0 || ct * kI8BN >= nvalid) ib = -1;
            };
            // Two items are kept in flight: w1 is worked on, w2 was drawn a whole tile earlier (so its value has long
            // arrived), and the state of w2's ring slot is read BEFORE w1's stages are issued: at the tile boundary the
            // producer normally holds everything it needs and the next tile's loads follow without a round trip to L2.
            unsigned long long w1 = atomicAdd(&g.ctr64[0], 1ull);
            unsigned long long w2 = atomicAdd(&g.ctr64[0], 1ull);
            int seen = 0; long long seen_ct = -1;                         // ready[] value last read for tile seen_ct
            for (;;) {
                long long ct; int ib;
                decode(w1, ct, ib);
                if (ib == -1) { w1 = w2; w2 = atomicAdd(&g.ctr64[0], 1ull); continue; }
                if (ib >= 0) {
                    const int slot = (int)(ct % g.R);
                    const int target = ((int)(ct / g.R) + 1) * g.nJ;
                    if (!(seen_ct == ct && seen >= target))
                        while (ld_acquire_gpu(&ready[slot]) < target) __nanosleep(64);
                    asm volatile("fence.proxy.async;" ::: "memory");      // the digits were written with generic stores; the bulk copies below read them through the async proxy
                }
                const int e = (int)(t & 3);
                tile_ring[e].ct = ct; tile_ring[e].ib = ib;
                mbar_arrive(&tile_full[e]);
                if (ib < 0) break;
                {   // look ahead: ring state of the next item, next-but-one item
                    long long ct2; int ib2;
                    decode(w2, ct2, ib2);
                    if (ib2 >= 0) { seen = ld_acquire_gpu(&ready[(int)(ct2 % g.R)]); seen_ct = ct2; }
                }
                const unsigned long long w3 = atomicAdd(&g.ctr64[0], 1ull);
                const int nk = (ib + 1) * (kI8BM / kI8KB);
                const uint8_t* wsrc = g.wq + (int64_t)ib * (ib + 1) / 2 * (kI8BM / kI8KB) * kI8ATile;
                const uint8_t* psrc = g.pa.panel + (ct % g.R) * (int64_t)(g.pa.n_pad / kI8KB) * kI8BTile;
                for (int kt = 0; kt < nk; kt++, it++) {
                    const int s = (int)(it & (kFusedStages - 1));
                    if (it >= (uint32_t)kFusedStages) mbar_wait(&empt[s], ((it / kFusedStages) - 1) & 1);
                    unsigned char* dst = smem_raw + (size_t)s * kI8Stage;
                    mbar_expect_tx(&full[s], kI8Stage);
                    bulk_g2s(dst, wsrc + (int64_t)kt * kI8ATile, kI8ATile, &full[s]);
                    bulk_g2s(dst + kI8ATile, psrc + (int64_t)kt * kI8BTile, kI8BTile, &full[s]);
                }
                w1 = w2; w2 = w3;
                t++;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------------------------------------------------------- tcgen05.mma issuer
        if (lane == 0) {
            long long t = 0; uint32_t it = 0;
            for (;;) {
                const int e = (int)(t & 3);
                mbar_wait(&tile_full[e], (uint32_t)(t >> 2) & 1);
                const int ib = tile_ring[e].ib;
                if (ib < 0) break;
                const int nk = (ib + 1) * (kI8BM / kI8KB);
                if (t > 0) mbar_wait(tmem_empty, (uint32_t)(t - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kt = 0; kt < nk; kt++, it++) {
                    const int s = (int)(it & (kFusedStages - 1));
                    mbar_wait(&full[s], (it / kFusedStages) & 1);
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
                                // level 7 is first written here (slice 1 x panel slice 6): split off, not accumulating -- no TMEM zeroing pass
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
                umma_commit(accbar);                                  // all MMAs of the tile done -> epilogue
                t++;
            }
        }
        __syncwarp();
    } else if (warp < 10) {
        // ---------------------------------------------------------------- epilogue + finalisation of completed candidate tiles
        const int q4 = warp & 3;                                      // TMEM lane quarter (hardware rule: warp id % 4)
        const int half = (warp - 2) >> 2;                             // which 32 of the 64 candidates
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        const int csel = (b4 ? 4 : 0) | (b3 ? 2 : 0) | (b2 ? 1 : 0);
        const int ep = tid - 64;                                      // 0..255; 0..63 = candidate of the tile in the 64-thread steps
        const long long nvalid = fused_valid(g.pa);
        double cbest = -INFINITY; long long cbi = kFusedNoIndex;      // running winner of this CTA (thread ep == 0)
        for (long long t = 0;; t++) {
            const int e = (int)(t & 3);
            mbar_wait(&tile_full[e], (uint32_t)(t >> 2) & 1);
            const int ib = tile_ring[e].ib;
            const long long ct = tile_ring[e].ct;
            if (ib < 0) break;
            const int slot = (int)(ct % g.R);
            const double scale = g.wscale[ib * kI8BM + q4 * 32 + lane];
            mbar_wait(accbar, (uint32_t)t & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c0 = half * (kI8BN / 2); c0 < (half + 1) * (kI8BN / 2); c0 += 8) {
                uint32_t r[8][8];
#pragma unroll
                for (int l = 0; l < 8; l++) tmem_ld8(tmem + lane_base + l * kI8BN + c0, r[l]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    double a = i32_bits_to_f64(r[7][j]);
#pragma unroll
                    for (int l = 6; l >= 0; l--) a = fma(a, 0.00390625, i32_bits_to_f64(r[l][j]));      // Horner in 2^-8 (exact products)
                    a *= scale;
                    v[j] = a * a;
                }
                // sum over the 32 rows of this warp for 8 columns at once (transposing butterfly; fixed order, as in trigemm_i8_kernel)
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
            // TMEM is drained: the next tile's MMAs may start while this one is reduced and (possibly) finalised
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ep < kI8BN) {
                const double q = ((red[ep] + red[kI8BN + ep]) + red[2 * kI8BN + ep]) + red[3 * kI8BN + ep];
                __stcg(&g.qpart[(int64_t)ib * g.pa.S + (int64_t)slot * kI8BN + ep], q);
                __threadfence();
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");            // `red` may be rewritten; the 64 partial sums are published
            if (ep < kI8BN) {
                if (ep == 0) {
                    const int target = ((int)(ct / g.R) + 1) * g.nI;
                    *fin_flag = atom_add_acq_rel_gpu(&done[slot], 1) == target - 1;
                }
                asm volatile("bar.sync 4, 64;" ::: "memory");
                if (*fin_flag) {
                    // this CTA delivered the last row block of candidate tile ct: finalise it (finalize_kernel's arithmetic and order)
                    __threadfence();
                    const long long cl = ct * kI8BN + ep;
                    double score = -INFINITY; long long idx = kFusedNoIndex;
                    if (cl < nvalid) {
                        double q = 0.0, mu = 0.0;
                        for (int b = 0; b < g.nI; b++) q += __ldcg(&g.qpart[(int64_t)b * g.pa.S + (int64_t)slot * kI8BN + ep]);
                        for (int b = 0; b < g.nJ; b++) mu += __ldcg(&g.pa.mupart[(int64_t)b * g.pa.S + (int64_t)slot * kI8BN + ep]);
                        const double var = g.prior - q;
                        const double sigma = sqrt(fabs(var));                   // np.sqrt(np.abs(.)), point_selector.py:98
                        score = acquisition_value(g.kind, mu, sigma, g.explore, g.f_best);
                        idx = g.idx_map ? g.idx_map[g.pa.c0 + cl] : g.pa.c0 + cl;
                        if (g.mu_out) g.mu_out[cl] = mu;
                        if (g.sigma_out) g.sigma_out[cl] = sigma;
                        if (g.acq_out) g.acq_out[cl] = score;
                        if (score != score) { atomicExch(&g.ctr32[1], 1); score = -INFINITY; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double os = __shfl_xor_sync(0xffffffffu, score, o);
                        const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
                        if (better(os, oi, score, idx)) { score = os; idx = oi; }
                    }
                    if (lane == 0) { fin_s[warp - 2] = score; fin_i[warp - 2] = idx; }
                    asm volatile("bar.sync 4, 64;" ::: "memory");   // both warps have read the partials and posted their winners
                    if (ep == 0) {
                        if (better(fin_s[0], fin_i[0], cbest, cbi)) { cbest = fin_s[0]; cbi = fin_i[0]; }
                        if (better(fin_s[1], fin_i[1], cbest, cbi)) { cbest = fin_s[1]; cbi = fin_i[1]; }
                        red_release_gpu(&freed[slot], 1);             // the builders may overwrite the slot
                    }
                }
            }
        }
        // CTA winner -> global; the last CTA folds all of them into the sweep's record
        if (ep == 0) {
            bogp_result* mine = g.cta_best + blockIdx.x;
            mine->score = cbest; mine->index = cbi; mine->nan_flag = 0; mine->reserved = 0;
            __threadfence();
            if (atom_add_acq_rel_gpu(&g.ctr32[0], 1) == (int)gridDim.x - 1) {
                __threadfence();
                double s = -INFINITY; long long i = kFusedNoIndex;
                for (unsigned b = 0; b < gridDim.x; b++) {
                    const double bs = __ldcg(&g.cta_best[b].score);
                    const long long bi = __ldcg(reinterpret_cast<const long long*>(&g.cta_best[b].index));
                    if (better(bs, bi, s, i)) { s = bs; i = bi; }
                }
                int nanf = __ldcg(&g.ctr32[1]);
                if (g.fold_prev) {
                    const double ps = g.result->score; const long long pi = g.result->index;
                    if (better(ps, pi, s, i)) { s = ps; i = pi; }
                    nanf |= g.result->nan_flag;
                }
                g.result->score = s; g.result->index = i; g.result->nan_flag = nanf; g.result->reserved = 0;
            }
        }
    } else {
        // ---------------------------------------------------------------- panel builders (8 warps)
        const int bt = tid - 320;
        const long long nvalid = fused_valid(g.pa);
        const unsigned long long total_build = (unsigned long long)nct * (unsigned long long)g.nJ;
        uint32_t phase = 0;
        for (;;) {
            if (bt == 0) {
                const unsigned long long b = atomicAdd(&g.ctr64[1], 1ull);
                if (b < total_build) {
                    const long long ct = (long long)(b / (unsigned long long)g.nJ);
                    if (ct * kI8BN < nvalid) {                        // wait until the tile that held this slot has been finalised
                        const int slot = (int)(ct % g.R), gen = (int)(ct / g.R);
                        while (ld_acquire_gpu(&freed[slot]) < gen) __nanosleep(128);
                    }
                }
                *build_item = b;
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            const unsigned long long b = *build_item;
            if (b >= total_build) break;
            const long long ct = (long long)(b / (unsigned long long)g.nJ);
            const int jb = (int)(b % (unsigned long long)g.nJ);
            const int slot = (int)(ct % g.R);
            const bool ok = panel_tile<DIMP, UB, false, true>(g.pa, ct, slot, jb, psm, bt, phase);
            if (ok) __threadfence();
            asm volatile("bar.sync 2, 256;" ::: "memory");            // every thread's digits are out (and `build_item` has been read)
            if (ok && bt == 0) red_release_gpu(&ready[slot], 1);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (warp < 10) {
        asm volatile("bar.sync 3, 320;" ::: "memory");
        if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------
struct FusedLayout { size_t ctr, best, qpart, mupart, panel, total; };
static FusedLayout fused_layout(int64_t n_pad, int R) {
    FusedLayout l{}; size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    l.ctr = take(64 + (16 + 3 * (size_t)R) * 4);
    l.best = take(256 * sizeof(bogp_result));
    l.qpart = take((size_t)(n_pad / kI8BM) * R * kI8BN * 8);
    l.mupart = take((size_t)(n_pad / kAcqBM) * R * kI8BN * 8);
    l.panel = take((size_t)n_pad * kI8BN * kI8Slices * R);
    l.total = off;
    return l;
}

// group size: ~32 MB of panel digits per group of candidate tiles (the L2 working set of the stand-alone kernel)
static int fused_group(int64_t n_pad) {
    int G = (int)((32u << 20) / ((size_t)n_pad * kI8BN * kI8Slices));
    return G < 1 ? 1 : (G > 32 ? 32 : G);
}

size_t fused_workspace_bytes(int64_t n_pad) { return fused_layout(n_pad, 2 * fused_group(n_pad)).total; }

template <int DIMP>
static size_t fused_smem() { return (size_t)kFusedMiscOff + kFusedMiscSize + sizeof(PanelSmem<DIMP>); }

template <int DIMP, bool UB>
static int launch_fused_t(bogp_ctx* ctx, const FusedArgs& fa, int grid, cudaStream_t st) {
    static DeviceOnce configured;
    if (configured.need(ctx->device))
        BOGP_CUDA_CHECK(cudaFuncSetAttribute(acquire_fused_i8_kernel<DIMP, UB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem<DIMP>()));
    BOGP_PROFILED(ctx, BOGP_PROF_TRIGEMM, (acquire_fused_i8_kernel<DIMP, UB><<<grid, kFusedThreads, fused_smem<DIMP>(), st>>>(fa)));
    BOGP_LAUNCH_CHECK(ctx);
    return BOGP_OK;
}

// The whole sweep [a.c0, a.c_end) (a.cur = its length, or the capacity of a compacted array with a.d_count) in one launch.
// Returns 1 if the workspace cannot hold a ring of at least two candidate tiles (the caller uses the separate kernels).
int launch_acquire_fused(bogp_ctx* ctx, const AcqChunk& a, const FusedFinal& f, void* d_workspace, size_t workspace_bytes, cudaStream_t st) {
    int G = ctx->fused_group > 0 ? ctx->fused_group : fused_group(a.n_pad);
    while (G > 1 && fused_layout(a.n_pad, 2 * G).total > workspace_bytes) G--;
    const int R = 2 * G;
    const FusedLayout l = fused_layout(a.n_pad, R);
    if (l.total > workspace_bytes) return 1;
    char* base = static_cast<char*>(d_workspace);
    BOGP_CUDA_CHECK(cudaMemsetAsync(base + l.ctr, 0, l.best - l.ctr, st));
    FusedArgs fa{};
    fill_cand(fa.pa.cand, a);
    fa.pa.x_pad = a.x_pad; fa.pa.inv_ell2 = a.inv_ell2; fa.pa.alpha = a.alpha;
    fa.pa.panel = reinterpret_cast<uint8_t*>(base + l.panel); fa.pa.mupart = reinterpret_cast<double*>(base + l.mupart);
    fa.pa.c0 = a.c0; fa.pa.c_end = a.c_end; fa.pa.S = (int64_t)R * kI8BN; fa.pa.n = a.n; fa.pa.n_pad = a.n_pad; fa.pa.dim = a.dim;
    fa.pa.d_count = a.d_count; fa.pa.idx_list = a.idx_list;
    fa.wq = a.wq; fa.wscale = a.wscale; fa.qpart = reinterpret_cast<double*>(base + l.qpart);
    fa.ctr64 = reinterpret_cast<unsigned long long*>(base + l.ctr); fa.ctr32 = reinterpret_cast<int*>(base + l.ctr + 64);
    fa.nI = a.n_pad / kI8BM; fa.nJ = a.n_pad / kAcqBM; fa.R = R; fa.G = G; fa.b_signed = a.n_pad <= 8192 ? 0 : 1;
    fa.nct = (a.cur + kI8BN - 1) / kI8BN;
    fa.mu_out = f.mu_out; fa.sigma_out = f.sigma_out; fa.acq_out = f.acq_out; fa.idx_map = f.idx_map;
    fa.kind = f.kind; fa.explore = f.explore; fa.f_best = f.f_best; fa.prior = f.prior;
    fa.cta_best = reinterpret_cast<bogp_result*>(base + l.best); fa.result = f.result; fa.fold_prev = f.fold_prev;
    const long long items = fa.nct * fa.nI;
    int grid = ctx->sm_count < 256 ? ctx->sm_count : 256;
    if (items < grid) grid = (int)(items > 0 ? items : 1);
    const bool ub = a.n_pad <= 8192;       // unsigned panel digits while the int32 level sums cannot overflow
#define BOGP_FUSED(D) (ub ? launch_fused_t<D, true>(ctx, fa, grid, st) : launch_fused_t<D, false>(ctx, fa, grid, st))
    if (a.dim <= 2) return BOGP_FUSED(2);
    if (a.dim <= 4) return BOGP_FUSED(4);
    if (a.dim <= 6) return BOGP_FUSED(6);
    if (a.dim <= 8) return BOGP_FUSED(8);
    if (a.dim <= 10) return BOGP_FUSED(10);
    if (a.dim <= 12) return BOGP_FUSED(12);
    return BOGP_FUSED(16);
#undef BOGP_FUSED
}

}  // namespace bogp
