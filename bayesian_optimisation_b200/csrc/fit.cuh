// Internal interfaces between the translation units of libbogp.
#pragma once
#include "common.cuh"

struct bogp_fit;

namespace bogp {

int launch_gram(bogp_ctx* ctx, const double* d_a, int64_t na, int64_t na_valid, const double* d_b, int64_t nb,
                int64_t nb_valid, int dim, const double* d_inv_ell2, double jitter, double* d_k, int64_t ldk,
                bool lower_tiles_only, int batch, int64_t strideK);
int launch_inv_ell2(bogp_ctx* ctx, const double* d_ell, double* d_out, int64_t count);
int launch_alpha(bogp_ctx* ctx, const double* d_w, int64_t ldw, int64_t strideW, const double* d_y, double* d_v,
                 double* d_alpha, double* d_part, int n, int batch);
size_t alpha_scratch_doubles(int64_t n);

int cholesky_blocked(bogp_ctx* ctx, double* d_a, int64_t n, int64_t lda, int64_t strideA, double* d_w, int64_t ldw,
                     int64_t strideW, double* d_logdet, int* d_info, int batch, double* d_t, int64_t strideT, int64_t* w_level = nullptr);
size_t trtri_scratch_doubles(int64_t n);
size_t cholesky_scratch_doubles(int64_t n);
int trtri_recursive(bogp_ctx* ctx, const double* d_l, int64_t ldl, int64_t strideL, double* d_w, int64_t ldw,
                    int64_t strideW, double* d_t, int64_t strideT, int64_t n, int batch, int64_t b_start);
size_t packed_w_doubles(int64_t n_pad);
struct GemmArgs;
// TMA-fed NT GEMM (gemm_tma.cu); returns 1 when not applicable (caller falls back to the cp.async kernel)
int launch_gemm_tma_nt(bogp_ctx* ctx, const GemmArgs& g);

// one chunk of the acquisition sweep, as handed to either tensor path
struct AcqChunk {
    const double* points; const double* axes; int len[BOGP_MAX_DIM]; int off[BOGP_MAX_DIM]; double cross_jitter;
    const double* x_pad; const double* inv_ell2; const double* alpha;
    const double* wp;                 // FP64 path: fragment-packed W
    const uint8_t* wq; const double* wscale;   // INT8 path: digit tiles of W + row scales
    void* panel; double* qpart; double* mupart;
    int64_t c0, c_end, cur, S;
    int n, n_pad, dim;
    const int* d_count;               // screened sweeps: the candidates are a compacted array whose length lives on the device
    const long long* idx_list;        // screened grid sweeps: slot -> flat grid index of the compacted survivors
    const double* ft; int toff[BOGP_MAX_DIM]; int lenp[BOGP_MAX_DIM]; int tt;   // grid sweeps: per-axis kernel-factor tables (acquire_i8.cuh), or null
    int toffT[BOGP_MAX_DIM];          // the same tables transposed, ft[toffT[k] + g * n_pad + j] (rows over j: operands of the mean GEMM, screen_gemm.cu)
};
int launch_panel_i8(bogp_ctx* ctx, const AcqChunk& a, cudaStream_t stream, bool mu_only = false);
int launch_trigemm_i8(bogp_ctx* ctx, const AcqChunk& a, cudaStream_t stream);
// the fused persistent sweep kernel (acquire_fused.cu): finalisation parameters and outputs of one sweep
struct FusedFinal {
    double* mu_out; double* sigma_out; double* acq_out;      // indexed by candidate - c0 (or null)
    const long long* idx_map;                                // compacted sweeps: global flat index per slot
    int kind; double explore, f_best, prior;
    bogp_result* result;                                     // device record the winner is written to
    int fold_prev;                                           // 1: the record already holds a running winner, fold it in
};
int launch_acquire_fused(bogp_ctx* ctx, const AcqChunk& a, const FusedFinal& f, void* d_workspace, size_t workspace_bytes, cudaStream_t st);
size_t fused_workspace_bytes(int64_t n_pad);
// per-axis kernel-factor tables of a grid sweep: geometry (returns the size in bytes, 0 if `a` is not a grid) and build
size_t grid_table_geometry(AcqChunk& a);
int launch_grid_factors(bogp_ctx* ctx, const AcqChunk& a, double* d_ft, cudaStream_t stream);
size_t grid_table_reserve(int64_t n_pad);
// Screened arg-max-only grid sweep with the posterior means of ALL candidates from one fp64 GEMM per chunk (screen_gemm.cu).
// Returns 1 if the sweep is not eligible (the caller then screens with the mean-only panel pass).
struct CandDesc;
int gemm_screen_sweep(bogp_ctx* ctx, const bogp_fit* fit, const AcqChunk& tab, int64_t c_begin, int64_t c_end, int kind, double explore,
                      double f_best, double prior_diag, void* d_workspace, size_t workspace_bytes, bogp_result* d_result);
size_t i8_wq_bytes(int64_t n_pad);
size_t i8_panel_bytes(int64_t n_pad, int64_t S);
int launch_slice_w(bogp_ctx* ctx, const double* d_w, int64_t n_pad, int* d_wexp, double* d_wscale, uint8_t* d_wq);
const uint8_t* fit_wq(const bogp_fit* f);
const double* fit_wscale(const bogp_fit* f);

const double* fit_wp(const bogp_fit* f);
// Operand packing of W for the given tensor path, if the fit has not done it yet (the fit packs for the path that
// is selected when it runs; the other packing is made on first use).  Enqueued on ctx->stream.
int fit_ensure_packed(bogp_ctx* ctx, const bogp_fit* f, int path);
const double* fit_xpad(const bogp_fit* f);
const double* fit_inv_ell2(const bogp_fit* f);
const double* fit_alpha(const bogp_fit* f);
int64_t fit_n(const bogp_fit* f);
double fit_jitter(const bogp_fit* f);
int fit_dim(const bogp_fit* f);

}  // namespace bogp
