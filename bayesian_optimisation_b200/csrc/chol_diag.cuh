// K2 diagonal block kernel (shared by fit.cu and tools/diag_bench.cu).
#pragma once
#include "common.cuh"

namespace bogp {

// ------------------------------------------------------------------------------------------------
// K2 diagonal block: 64x64 Cholesky + inverse of the factor, one CTA per matrix of the batch.
// Register-tiled (4x4 per thread, for the matrix and for the inverse), square-root-free inner loop,
// columns in groups of four with one barrier per group (see chol_diag_block).  The serial chain is
// latency-bound: a dependent fp64 op, a 64-bit shuffle and rcp.approx.f64 each cost ~24 cycles on
// B200 (profiles/r01_latency_bench.txt).
// ------------------------------------------------------------------------------------------------
struct DiagArgs {
    double* a; int64_t lda; int64_t strideA;
    double* w; int64_t ldw; int64_t strideW;
    double* logdet; int* info; int kblk;
};

// Thread (ty, tx) of a 16 x 16 grid owns the 4x4 sub-blocks A[4ty.., 4tx..] and R[4ty.., 4tx..] in
// REGISTERS.  Right-looking elimination, one __syncthreads per column:
//   * the 16 owners of column j sit in one half-warp: the pivot a_jj is broadcast by a warp shuffle;
//     the unscaled column u = A[:,j] and u / a_jj go to shared memory (trailing update
//     A[i,c] -= u_i u_c / a_jj, so only ONE reciprocal sits on the critical path; the square root
//     for the stored factor L[:,j] = u / sqrt(a_jj) is computed off it);
//   * the inverse is built in the same sweep: with R = I initially, row j of X = L^-1 is R[j,:]/d_j
//     and every later row gets R[i,:] -= L[i,j] X[j,:] = (u_i / a_jj) R[j,:]  (same rank-1 shape);
//   * after the barrier every thread applies both rank-1 updates to its registers (32 DFMA).
// Column/row buffers are double-buffered so that one barrier per column suffices.
#ifndef BOGP_DIAG_GROUPS
#define BOGP_DIAG_GROUPS (kDiagNB / 4)
#endif
struct DiagSmem {
    // column buffers: the value of row 4t + 2h + e sits at [32 h + 2 t + e], so that the 16 threads of a half-warp read and
    // write consecutive 16-byte pieces (conflict-free 128-bit accesses)
    double colU[2][4][kDiagNB];   // u_p[i] = A[i][j0+p] after the in-group updates (rows > j0+p, else 0)
    double colC[2][4][kDiagNB];   // u_p[i] / a_pp
    double rows[2][4][kDiagNB];   // R[j0+p][c] as it was when the group started
    double mul[2][4][4];          // in-group multipliers M[p][q] = colC[q][row j0+p], q < p
    double dg[kDiagNB];
};

#ifdef BOGP_DIAG_TRACE
__device__ long long g_diag_trace[64];
__device__ long long g_diag_trace2[16][8];
#define BOGP_DIAG_STAMP2(cond, g, k) do { if (cond) g_diag_trace2[g][k] = clock64(); } while (0)
#define BOGP_DIAG_STAMP(k) do { if (threadIdx.x == 0) g_diag_trace[k] = clock64(); } while (0)
#else
#define BOGP_DIAG_STAMP(k) do {} while (0)
#define BOGP_DIAG_STAMP2(cond, g, k) do {} while (0)
#endif

__device__ __forceinline__ double rcp_newton(double x) {   // hardware seed + 2 Newton steps (~1 ulp)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
}

// Device body: factor + invert block `kblk` of matrix `mat`; 256 threads; `sm` in shared memory.
//
// Columns are processed in groups of 4 with ONE barrier per group:
//   * the 16 owners of the group's columns (one half-warp) eliminate the 64 x 4 panel among themselves --
//     pivot and the three in-group multipliers travel by warp shuffles, one hardware-seeded reciprocal
//     per column is the only long-latency step;
//   * after the barrier every thread applies the rank-4 update to its 4x4 registers.  For the inverse
//     (R starts as I and becomes L^-1 up to a row scaling) the four pivot rows are published as they
//     were when the group started; the in-group dependencies are folded into the coefficients with the
//     4x4 unit-triangular multiplier matrix (6 FMAs per row) instead of a second barrier.
__device__ __forceinline__ void chol_diag_block(const DiagArgs& g, int mat, DiagSmem& sm) {
    constexpr int NB = kDiagNB;
    const int tid = threadIdx.x, tx = tid >> 4, ty = tid & 15, lane = tid & 31;
    double* A = g.a + mat * g.strideA + (int64_t)g.kblk * NB * (g.lda + 1);
    const bool active = ty >= tx;
    double a[4][4], r[4][4];
    const bool vec_in = ((g.lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = 4 * ty + i;
        const double* arow = A + (int64_t)row * g.lda + 4 * tx;
        if (ty > tx && vec_in) {                              // one 32-byte sector per thread and row: two 16-byte loads
            const double2 v01 = __ldcg(reinterpret_cast<const double2*>(arow));      // L2-coherent: other CTAs may have written it
            const double2 v23 = __ldcg(reinterpret_cast<const double2*>(arow + 2));
            a[i][0] = v01.x; a[i][1] = v01.y; a[i][2] = v23.x; a[i][3] = v23.y;
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) a[i][c] = (active && 4 * tx + c <= row) ? __ldcg(arow + c) : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) r[i][c] = (row == 4 * tx + c) ? 1.0 : 0.0;
    }
    const unsigned half_mask = 0xFFFFu << (lane & 16);
    for (int jb = 0; jb < NB / 4; jb++) {
        const int j0 = 4 * jb, buf = jb & 1;
        BOGP_DIAG_STAMP(3 * jb);
        if (tx == jb) {                                       // the half-warp that owns columns j0..j0+3 (all 16 lanes)
            const int dl = (lane & 16) + jb;                  // lane of the diagonal thread (ty == jb)
            BOGP_DIAG_STAMP2(ty == jb, jb, 0);
            double up[4][4], cp[4][4];
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const double app = __shfl_sync(half_mask, a[p][p], dl);
                const double rinv = rcp_newton(app);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    up[p][i] = (4 * ty + i > j0 + p) ? a[i][p] : 0.0;      // rows <= j0+p are exact zeros: readers need no masks
                    cp[p][i] = up[p][i] * rinv;
                }
#pragma unroll
                for (int q = p + 1; q < 4; q++) {             // in-group update of the later columns
                    const double s = __shfl_sync(half_mask, a[q][p], dl);   // u_p at row j0+q
#pragma unroll
                    for (int i = 0; i < 4; i++) a[i][q] -= cp[p][i] * s;
                }
                if (ty == jb) sm.dg[j0 + p] = app;
            }
            BOGP_DIAG_STAMP2(ty == jb, jb, 1);
#pragma unroll
            for (int p = 0; p < 4; p++) {
                *reinterpret_cast<double2*>(&sm.colU[buf][p][2 * ty])      = make_double2(up[p][0], up[p][1]);
                *reinterpret_cast<double2*>(&sm.colU[buf][p][32 + 2 * ty]) = make_double2(up[p][2], up[p][3]);
                *reinterpret_cast<double2*>(&sm.colC[buf][p][2 * ty])      = make_double2(cp[p][0], cp[p][1]);
                *reinterpret_cast<double2*>(&sm.colC[buf][p][32 + 2 * ty]) = make_double2(cp[p][2], cp[p][3]);
            }
            if (ty == jb) {
#pragma unroll
                for (int p = 0; p < 4; p++)
#pragma unroll
                    for (int q = 0; q < 4; q++) sm.mul[buf][p][q] = (q < p) ? cp[q][p] : 0.0;
            }
        }
        if (ty == jb && active) {                             // owners of rows j0..j0+3 of R (zero right of the diagonal)
#pragma unroll
            for (int p = 0; p < 4; p++) {
                *reinterpret_cast<double2*>(&sm.rows[buf][p][4 * tx])     = make_double2(r[p][0], r[p][1]);
                *reinterpret_cast<double2*>(&sm.rows[buf][p][4 * tx + 2]) = make_double2(r[p][2], r[p][3]);
            }
        }
        BOGP_DIAG_STAMP2(tx == jb && ty == jb, jb, 2);
        __syncthreads();
        BOGP_DIAG_STAMP2(tx == jb + 1 && ty == jb + 1, jb, 3);
        BOGP_DIAG_STAMP(3 * jb + 1);
        if (active && 4 * ty + 3 > j0) {
            double c[4][4];                                   // c[p][i] = colC_p[row i]
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const double2 c01 = *reinterpret_cast<const double2*>(&sm.colC[buf][p][2 * ty]);
                const double2 c23 = *reinterpret_cast<const double2*>(&sm.colC[buf][p][32 + 2 * ty]);
                c[p][0] = c01.x; c[p][1] = c01.y; c[p][2] = c23.x; c[p][3] = c23.y;
            }
            if (tx > jb) {                                    // trailing columns (the group's own columns were updated by their owners)
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const double2 k01 = *reinterpret_cast<const double2*>(&sm.colU[buf][p][2 * tx]);
                    const double2 k23 = *reinterpret_cast<const double2*>(&sm.colU[buf][p][32 + 2 * tx]);
                    const double cc[4] = {k01.x, k01.y, k23.x, k23.y};
#pragma unroll
                    for (int cidx = 0; cidx < 4; cidx++)
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i][cidx] -= c[p][i] * cc[cidx];
                }
            }
#ifdef BOGP_DIAG_NO_INVERSE
            if (false) {
#else
            if (tx <= jb) {                                   // inverse: fold the in-group dependencies into the coefficients
#endif
                double m[4][4];
#pragma unroll
                for (int p = 1; p < 4; p++)
#pragma unroll
                    for (int q = 0; q < p; q++) m[p][q] = sm.mul[buf][p][q];
#pragma unroll
                for (int i = 0; i < 4; i++) {                 // d = M^-T c  (M unit lower triangular)
                    c[2][i] -= m[3][2] * c[3][i];
                    c[1][i] -= m[2][1] * c[2][i] + m[3][1] * c[3][i];
                    c[0][i] -= m[1][0] * c[1][i] + m[2][0] * c[2][i] + m[3][0] * c[3][i];
                }
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const double2 x01 = *reinterpret_cast<const double2*>(&sm.rows[buf][p][4 * tx]);
                    const double2 x23 = *reinterpret_cast<const double2*>(&sm.rows[buf][p][4 * tx + 2]);
                    const double xr[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
                    for (int cidx = 0; cidx < 4; cidx++)
#pragma unroll
                        for (int i = 0; i < 4; i++) r[i][cidx] -= c[p][i] * xr[cidx];
                }
            }
        }
        BOGP_DIAG_STAMP2(tx == jb + 1 && ty == jb + 1, jb, 4);
        BOGP_DIAG_STAMP2(tx == jb && ty == jb, jb, 5);
        BOGP_DIAG_STAMP(3 * jb + 2);
    }
    __syncthreads();
    BOGP_DIAG_STAMP(48);
    double (&isd_s)[NB] = sm.colU[0][0];
    double (&d_s)[NB] = sm.colU[0][1];
    if (tid < NB) {                                           // isd_j = 1/sqrt(a_jj), d_j = a_jj * isd_j
        const double ajj = sm.dg[tid];
        if (!(ajj > 0.0) || isinf(ajj)) {                     // report the first bad pivot (1-based)
            const int idx = g.kblk * NB + tid + 1;
            if (atomicCAS(g.info + mat, 0, idx) != 0) atomicMin(g.info + mat, idx);
        }
        const double isd = rsqrt(ajj);
        isd_s[tid] = isd;
        d_s[tid] = ajj * isd;
    }
    __syncthreads();
    double* W = g.w ? g.w + mat * g.strideW + (int64_t)g.kblk * NB * (g.ldw + 1) : nullptr;
    const bool vec = (((g.lda | g.ldw) & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
    const double2 isd01 = *reinterpret_cast<const double2*>(&isd_s[4 * tx]), isd23 = *reinterpret_cast<const double2*>(&isd_s[4 * tx + 2]);
    const double isd_c[4] = {isd01.x, isd01.y, isd23.x, isd23.y};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = 4 * ty + i;
        const double isd_row = isd_s[row];
        double la[4], lw[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int cc = 4 * tx + c;
            la[c] = (cc == row) ? d_s[row] : a[i][c] * isd_c[c];
            lw[c] = (active && cc <= row) ? r[i][c] * isd_row : 0.0;
        }
        // a thread owns 4 consecutive columns of the row = one 32-byte sector: two 16-byte stores when aligned
        double* arow = A + (int64_t)row * g.lda + 4 * tx;
        if (ty > tx && vec) {                                 // entirely below the diagonal
            *reinterpret_cast<double2*>(arow) = make_double2(la[0], la[1]);
            *reinterpret_cast<double2*>(arow + 2) = make_double2(la[2], la[3]);
        } else if (ty >= tx) {
#pragma unroll
            for (int c = 0; c < 4; c++) if (4 * tx + c <= row) arow[c] = la[c];
        }
        if (W) {
            double* wrow = W + (int64_t)row * g.ldw + 4 * tx;
            if (vec) {
                *reinterpret_cast<double2*>(wrow) = make_double2(lw[0], lw[1]);
                *reinterpret_cast<double2*>(wrow + 2) = make_double2(lw[2], lw[3]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) wrow[c] = lw[c];
            }
        }
    }
    if (tid < 32 && g.logdet) {   // log det = sum log a_jj (= 2 sum log d_j), fixed order
        double s = log(sm.dg[tid]) + log(sm.dg[tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) atomicAdd(g.logdet + mat, s);      // fire-and-forget (no read-modify-write round trip on the chain); one add per
                                                          // diagonal block, issued in factorisation order, so the sum is still deterministic
    }
    BOGP_DIAG_STAMP(49);
}

__global__ void __launch_bounds__(256) chol_diag_kernel(DiagArgs g) {
    __shared__ __align__(16) DiagSmem sm;
    chol_diag_block(g, (int)blockIdx.x, sm);
}

// Batched variant (one CTA per matrix, hundreds of matrices): the kernel is latency-bound, so two CTAs per SM
// nearly double the throughput; 128 registers per thread instead of 138.
__global__ void __launch_bounds__(256, 2) chol_diag_batched_kernel(DiagArgs g) {
    __shared__ __align__(16) DiagSmem sm;
    chol_diag_block(g, (int)blockIdx.x, sm);
}

}  // namespace bogp
