// Earlier variants of the Cholesky diagonal-block kernel, kept only for tools/diag_bench.cu (timing comparisons).
// Not part of libbogp.
#pragma once
#include "chol_diag.cuh"

namespace bogp {

struct DiagSmemV2 {
    double col[2][kDiagNB];    // u_i = a[i][j] before scaling (rows > j)
    double col2[2][kDiagNB];   // u_i / a_jj
    double xrow[2][kDiagNB];   // R[j][c]  (unscaled row j of the inverse)
    double dg[kDiagNB];
};

// Device body: factor + invert block `kblk` of matrix `mat`; 256 threads; `sm` in shared memory.
__device__ __forceinline__ void chol_diag_block_v2(const DiagArgs& g, int mat, DiagSmemV2& sm) {
    constexpr int NB = kDiagNB;
    double (&col)[2][NB] = sm.col; double (&col2)[2][NB] = sm.col2; double (&xrow)[2][NB] = sm.xrow; double (&dg)[NB] = sm.dg;
    const int tid = threadIdx.x, tx = tid >> 4, ty = tid & 15, lane = tid & 31;
    double* A = g.a + mat * g.strideA + (int64_t)g.kblk * NB * (g.lda + 1);
    const bool active = ty >= tx;
    double a[4][4], r[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int row = 4 * ty + i, cc = 4 * tx + c;
            a[i][c] = (active && cc <= row) ? __ldcg(A + (int64_t)row * g.lda + cc) : 0.0;   // L2-coherent: other CTAs may have written it
            r[i][c] = (row == cc) ? 1.0 : 0.0;
        }
    const unsigned half_mask = 0xFFFFu << (lane & 16);
    for (int jb = 0; jb < BOGP_DIAG_GROUPS; jb++) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int j = 4 * jb + jj, buf = j & 1;
            if (tx == jb) {                                   // the half-warp that owns column j (all 16 lanes)
                const double ajj = __shfl_sync(half_mask, a[jj][jj], (lane & 16) + jb);
                // Critical path: ONE reciprocal (hardware seed + 2 Newton steps).  The loop runs the
                // square-root-free form A[i,c] -= u_i u_c / a_jj; the 64 inverse square roots that
                // turn u into L (and R into L^-1) are taken once, in parallel, after the loop.
                double rinv;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(ajj));
                double e = fma(-ajj, rinv, 1.0); rinv = fma(rinv, e, rinv);
                e = fma(-ajj, rinv, 1.0); rinv = fma(rinv, e, rinv);
                // rows <= j are written as exact zeros, so the readers need no masks
                double u[4];
#pragma unroll
                for (int i = 0; i < 4; i++) u[i] = (4 * ty + i > j) ? a[i][jj] : 0.0;
                *reinterpret_cast<double2*>(&col[buf][4 * ty])      = make_double2(u[0], u[1]);
                *reinterpret_cast<double2*>(&col[buf][4 * ty + 2])  = make_double2(u[2], u[3]);
                *reinterpret_cast<double2*>(&col2[buf][4 * ty])     = make_double2(u[0] * rinv, u[1] * rinv);
                *reinterpret_cast<double2*>(&col2[buf][4 * ty + 2]) = make_double2(u[2] * rinv, u[3] * rinv);
                if (ty == jb) dg[j] = ajj;                    // pivot; its rsqrt is taken after the loop
            }
            if (ty == jb && active) {                         // owners of row j of R (zero right of the diagonal)
                *reinterpret_cast<double2*>(&xrow[buf][4 * tx])     = make_double2(r[jj][0], r[jj][1]);
                *reinterpret_cast<double2*>(&xrow[buf][4 * tx + 2]) = make_double2(r[jj][2], r[jj][3]);
            }
            __syncthreads();
            if (active && 4 * ty + 3 > j) {
                const double2 c01 = *reinterpret_cast<const double2*>(&col2[buf][4 * ty]);
                const double2 c23 = *reinterpret_cast<const double2*>(&col2[buf][4 * ty + 2]);
                const double c2[4] = {c01.x, c01.y, c23.x, c23.y};
                if (tx >= jb) {
                    const double2 k01 = *reinterpret_cast<const double2*>(&col[buf][4 * tx]);
                    const double2 k23 = *reinterpret_cast<const double2*>(&col[buf][4 * tx + 2]);
                    const double cc[4] = {k01.x, k01.y, k23.x, k23.y};
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i][c] -= c2[i] * cc[c];
                }
                if (tx <= jb) {
                    const double2 x01 = *reinterpret_cast<const double2*>(&xrow[buf][4 * tx]);
                    const double2 x23 = *reinterpret_cast<const double2*>(&xrow[buf][4 * tx + 2]);
                    const double xr[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int i = 0; i < 4; i++) r[i][c] -= c2[i] * xr[c];
                }
            }
        }
    }
    __syncthreads();
    if (tid < NB) {                                           // isd_j = 1/sqrt(a_jj), d_j = a_jj * isd_j
        const double ajj = dg[tid];
        if (!(ajj > 0.0) || isinf(ajj)) {                     // report the first bad pivot (1-based)
            const int idx = g.kblk * NB + tid + 1;
            if (atomicCAS(g.info + mat, 0, idx) != 0) atomicMin(g.info + mat, idx);
        }
        const double isd = rsqrt(ajj);
        col[0][tid] = isd;
        col2[0][tid] = ajj * isd;
    }
    __syncthreads();
    double* W = g.w ? g.w + mat * g.strideW + (int64_t)g.kblk * NB * (g.ldw + 1) : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = 4 * ty + i;
        const double isd_row = col[0][row];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int cc = 4 * tx + c;
            if (active && cc <= row) A[(int64_t)row * g.lda + cc] = (cc == row) ? col2[0][row] : a[i][c] * col[0][cc];
            if (W) W[(int64_t)row * g.ldw + cc] = (active && cc <= row) ? r[i][c] * isd_row : 0.0;
        }
    }
    if (tid < 32 && g.logdet) {   // log det = sum log a_jj (= 2 sum log d_j), fixed order
        double s = log(dg[tid]) + log(dg[tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) atomicAdd(g.logdet + mat, s);      // fire-and-forget (no read-modify-write round trip on the chain); one add per
                                                          // diagonal block, issued in factorisation order, so the sum is still deterministic
    }
}

__global__ void __launch_bounds__(256) chol_diag_kernel_v2(DiagArgs g) {
    __shared__ __align__(16) DiagSmemV2 sm;
    chol_diag_block_v2(g, (int)blockIdx.x, sm);
}

__global__ void __launch_bounds__(256) chol_diag_kernel_v1(DiagArgs g) {
    constexpr int NB = kDiagNB;
    __shared__ __align__(16) double col[2][NB];    // u_i = a[i][j] before scaling (rows > j)
    __shared__ __align__(16) double col2[2][NB];   // u_i / a_jj
    __shared__ __align__(16) double xrow[2][NB];   // R[j][c]  (unscaled row j of the inverse)
    __shared__ double dg[NB];
    const int tid = threadIdx.x, tx = tid >> 4, ty = tid & 15, lane = tid & 31;
    double* A = g.a + blockIdx.x * g.strideA + (int64_t)g.kblk * NB * (g.lda + 1);
    const bool active = ty >= tx;
    double a[4][4], r[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int row = 4 * ty + i, cc = 4 * tx + c;
            a[i][c] = (active && cc <= row) ? A[(int64_t)row * g.lda + cc] : 0.0;
            r[i][c] = (row == cc) ? 1.0 : 0.0;
        }
    const unsigned half_mask = 0xFFFFu << (lane & 16);
    for (int jb = 0; jb < NB / 4; jb++) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int j = 4 * jb + jj, buf = j & 1;
            if (tx == jb) {                                   // the half-warp that owns column j
                const double ajj = __shfl_sync(half_mask, a[jj][jj], (lane & 16) + jb);
                // Critical path: ONE reciprocal (hardware seed + 2 Newton steps).  The loop runs the
                // square-root-free form A[i,c] -= u_i u_c / a_jj; the 64 inverse square roots that
                // turn u into L (and R into L^-1) are taken once, in parallel, after the loop.
                double rinv;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(ajj));
                double e = fma(-ajj, rinv, 1.0); rinv = fma(rinv, e, rinv);
                e = fma(-ajj, rinv, 1.0); rinv = fma(rinv, e, rinv);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int row = 4 * ty + i;
                    if (row > j) { col[buf][row] = a[i][jj]; col2[buf][row] = a[i][jj] * rinv; }
                }
                if (ty == jb) dg[j] = ajj;                    // pivot; its rsqrt is taken after the loop
            }
            if (ty == jb && active) {                         // owners of row j of R
#pragma unroll
                for (int c = 0; c < 4; c++) xrow[buf][4 * tx + c] = r[jj][c];
            }
            __syncthreads();
            if (active && 4 * ty + 3 > j) {
                double c2[4];
#pragma unroll
                for (int i = 0; i < 4; i++) c2[i] = (4 * ty + i > j) ? col2[buf][4 * ty + i] : 0.0;
                if (tx >= jb) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const double cc = (4 * tx + c > j) ? col[buf][4 * tx + c] : 0.0;
#pragma unroll
                        for (int i = 0; i < 4; i++) a[i][c] -= c2[i] * cc;
                    }
                }
                if (tx <= jb) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const double xr = (4 * tx + c <= j) ? xrow[buf][4 * tx + c] : 0.0;
#pragma unroll
                        for (int i = 0; i < 4; i++) r[i][c] -= c2[i] * xr;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tid < NB) {                                           // isd_j = 1/sqrt(a_jj), d_j = a_jj * isd_j
        const double ajj = dg[tid];
        if (!(ajj > 0.0) || isinf(ajj)) {                     // report the first bad pivot (1-based)
            const int idx = g.kblk * NB + tid + 1;
            if (atomicCAS(g.info + blockIdx.x, 0, idx) != 0) atomicMin(g.info + blockIdx.x, idx);
        }
        const double isd = rsqrt(ajj);
        col[0][tid] = isd;
        col2[0][tid] = ajj * isd;
    }
    __syncthreads();
    double* W = g.w ? g.w + blockIdx.x * g.strideW + (int64_t)g.kblk * NB * (g.ldw + 1) : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int row = 4 * ty + i;
        const double isd_row = col[0][row];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int cc = 4 * tx + c;
            if (active && cc <= row) A[(int64_t)row * g.lda + cc] = (cc == row) ? col2[0][row] : a[i][c] * col[0][cc];
            if (W) W[(int64_t)row * g.ldw + cc] = (active && cc <= row) ? r[i][c] * isd_row : 0.0;
        }
    }
    if (tid < 32 && g.logdet) {   // log det = sum log a_jj (= 2 sum log d_j), fixed order
        double s = log(dg[tid]) + log(dg[tid + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) g.logdet[blockIdx.x] += s;
    }
}

}  // namespace bogp
