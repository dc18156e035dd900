/*
 * bogp.h -- C ABI of libbogp.so: the B200 (sm_100a) Gaussian-process surrogate +
 * acquisition hot path that replaces the numpy arithmetic inside the reference's
 * `PointSelector` (reference: point_selector.py:13-207, driven by
 * select_parameters.py:146-157,282-293).
 *
 * The reference is pure Python and has no FFI of its own, so there is no existing
 * binding to mirror; each entry point below cites the reference lines whose
 * arithmetic it replaces.  The Python host (bayesian_optimisation_b200/_lib.py)
 * binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every `d_*` pointer is a DEVICE pointer (fp64 row-major / int64 indices);
 *     `h_*` pointers are host pointers.  No torch types cross this boundary.
 *   - every call is asynchronous on the stream given to bogp_set_stream() unless it
 *     returns host values (those synchronise that stream before returning).
 *   - return value: BOGP_OK or a negative status; bogp_last_error() has the text.
 *   - no CPU fallback exists: without a CUDA device bogp_create() fails.
 */
#ifndef BOGP_H
#define BOGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOGP_OK                 0
#define BOGP_ERR_BAD_ARG       -1   /* bad shape / null pointer / unsupported size      */
#define BOGP_ERR_CUDA          -2   /* CUDA runtime error (text in bogp_last_error)     */
#define BOGP_ERR_NOT_POSDEF    -3   /* Cholesky pivot <= 0 or not finite  (reference: LinAlgError from np.linalg.inv, point_selector.py:89,117) */
#define BOGP_ERR_NAN_SCORE     -4   /* NaN acquisition value (reference: IndexError at point_selector.py:207) */
#define BOGP_ERR_WORKSPACE     -5   /* caller-provided workspace too small              */

#define BOGP_MAX_DIM           16   /* features per point                                */
#define BOGP_ACQ_LCB            0   /* explore*sigma - mu        point_selector.py:204  */
#define BOGP_ACQ_EI             1   /* expected improvement (extension)   */

typedef struct bogp_ctx bogp_ctx;

/* version / build info: "bogp <n> sm_100a" */
const char* bogp_version(void);
const char* bogp_last_error(void);

/* One context per process and device.  Allocates a few KB of device scratch. */
int  bogp_create(int device, bogp_ctx** out);
void bogp_destroy(bogp_ctx* ctx);
int  bogp_set_stream(bogp_ctx* ctx, void* cuda_stream);
int  bogp_sm_count(const bogp_ctx* ctx);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
int64_t bogp_launch_count(const bogp_ctx* ctx);
/* Which tensor path the acquisition product V = L^-1 k_* runs on:
 *   BOGP_PATH_FP64_DMMA    mma.sync DMMA.8x8x4 on the FP64 pipe (the straightforward fp64 product);
 *   BOGP_PATH_INT8_TCGEN05 exact integer digit slices on tcgen05.mma kind::i8 with TMEM accumulators
 *                          (error-free splitting; the truncation is at most 5 K 2^-62 of the row scale of L^-1
 *                          -- 2^-47.7 at K = 4096, about 2^-56 on real data -- i.e. inside the rounding error
 *                          of an fp64 dot product; the result is independent of any summation order).
 * Default: environment variable BOGP_ACQUIRE_PATH ("fp64" | "i8"), else BOGP_PATH_DEFAULT.            */
#define BOGP_PATH_FP64_DMMA     0
#define BOGP_PATH_INT8_TCGEN05  1
#define BOGP_PATH_DEFAULT       BOGP_PATH_INT8_TCGEN05
int bogp_set_acquire_path(bogp_ctx* ctx, int path);
int bogp_get_acquire_path(const bogp_ctx* ctx);
/* Measurement aid: when enabled, each kernel of the acquisition sweep is bracketed by CUDA
 * events on the launching stream (this serialises the stream; never enable it inside a timed
 * region).  kernel_id: 0 panel, 1 tri-GEMM, 2 finalize, 3 merge.                            */
int bogp_profile(bogp_ctx* ctx, int enable);
int bogp_profile_read(const bogp_ctx* ctx, int kernel_id, double* ms_total, int64_t* launches);

/* ---- K1: ARD squared-exponential Gram matrix ------------------- point_selector.py:166-195
 * K[i,j] = exp(-0.5 * sum_k (a_ik-b_jk)^2 / ell_k^2) (+ jitter where i == j), row-major,
 * leading dimension ldk.  The caller decides about the jitter (the reference's rule is
 * "shapes equal", point_selector.py:173-177); pass 0.0 for none.                        */
int bogp_kernel_matrix(bogp_ctx* ctx, const double* d_a, int64_t na, const double* d_b, int64_t nb,
                       int dim, const double* d_ell, double jitter, double* d_k, int64_t ldk);

/* ---- K2 + fit: Cholesky, log-det, alpha = K^-1 y, W = L^-1 ------ point_selector.py:79,89-90,117-119
 * Replaces np.linalg.inv / np.linalg.det.  The fitted state lives in a caller-provided
 * device workspace (bogp_fit_workspace_bytes) and is referred to by a small host handle. */
typedef struct bogp_fit bogp_fit;

size_t bogp_fit_workspace_bytes(int64_t n, int dim);
/* Builds K(X,X)+jitter*I, factors it, computes log det, alpha, W=L^-1 (packed for the
 * acquisition kernel).  h_nlml_out (optional) receives
 * 0.5*(y^T K^-1 y + log det K + n log 2pi) (point_selector.py:119).                     */
int  bogp_fit_create(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim,
                     const double* d_ell, double jitter, void* d_workspace, size_t workspace_bytes,
                     bogp_fit** out, double* h_nlml_out);
/* The same in two steps: bogp_fit_enqueue only enqueues device work (no synchronisation, so the
 * host may record it into a CUDA graph and replay it); bogp_fit_status synchronises and reports
 * BOGP_ERR_NOT_POSDEF / the nlml of the most recent execution.                                    */
int  bogp_fit_enqueue(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim,
                      const double* d_ell, double jitter, void* d_workspace, size_t workspace_bytes,
                      bogp_fit** out);
int  bogp_fit_status(bogp_fit* fit, double* h_nlml_out);
void bogp_fit_destroy(bogp_fit* fit);
/* inspection (device pointers into the workspace; n_pad = n rounded up to 256) */
int64_t       bogp_fit_n_pad(const bogp_fit* fit);
const double* bogp_fit_chol(const bogp_fit* fit);     /* L, lower triangle of an n_pad x n_pad row-major matrix */
const double* bogp_fit_linv(const bogp_fit* fit);     /* W = L^-1, same layout                                  */
const double* bogp_fit_alpha(const bogp_fit* fit);    /* n_pad doubles                                          */
double        bogp_fit_logdet(const bogp_fit* fit);   /* host value (synchronises)                              */

/* stand-alone blocked Cholesky of a row-major lower-stored n x n matrix, in place
 * (n multiple of 64).  d_linv (n x n, zero-initialised, required) receives the inverse of the
 * 64x64 diagonal blocks on its diagonal blocks (the panel solve multiplies with them).  d_logdet: one double.  d_info: one int
 * (0, or 1-based index of the first bad pivot).                                          */
int bogp_cholesky(bogp_ctx* ctx, double* d_a, int64_t n, int64_t lda, double* d_linv,
                  double* d_logdet, int* d_info);

/* ---- K4: fused acquisition sweep ----------------------------- point_selector.py:81,90-98,204-207
 * Candidates are either an explicit device array (d_candidates != NULL, row-major
 * c_total x dim) or a Cartesian grid (d_candidates == NULL): d_axes holds the dim axes
 * back to back, h_axis_len their lengths, flat index row-major with axis 0 slowest
 * (select_parameters.py:273-279).  Scores flat indices [c_begin, c_end).
 * Outputs (each optional, device, length c_end-c_begin): mu, sigma, acq.
 * Always produces the best (score, flat index) of the range with the reference's tie
 * rule -- largest score, then smallest flat index (np.argwhere(a == amax(a))[0]).       */
typedef struct bogp_candidates {
    const double*  d_points;      /* explicit mode: c_total x dim, or NULL               */
    const double*  d_axes;        /* grid mode: concatenated axis values                 */
    const int32_t* h_axis_len;    /* grid mode: dim lengths (host)                       */
    int64_t        c_total;       /* total number of candidates                          */
    double         cross_jitter;  /* added to k(x_i, p_c) where i == c: reproduces the
                                     reference's shape-equality jitter quirk when M == C
                                     (point_selector.py:173-177 applied at :81); else 0   */
} bogp_candidates;

size_t bogp_acquire_workspace_bytes(const bogp_fit* fit, int64_t max_chunk);
int bogp_acquire(bogp_ctx* ctx, const bogp_fit* fit, const bogp_candidates* cand,
                 int64_t c_begin, int64_t c_end, int kind, double explore, double f_best,
                 double prior_diag, double* d_mu_out, double* d_sigma_out, double* d_acq_out,
                 void* d_workspace, size_t workspace_bytes,
                 double* h_best_score, int64_t* h_best_index);

/* acquisition + arg-max only, on mu/sigma already on the device
 * (lower_confidence_bound(), point_selector.py:197-207).                                */
int bogp_score_argmax(bogp_ctx* ctx, const double* d_mu, const double* d_sigma, int64_t c,
                      int kind, double explore, double f_best, double* d_acq_out,
                      double* h_best_score, int64_t* h_best_index);

/* ---- K3: batched negative log marginal likelihood (+ gradient) -- point_selector.py:111-138
 * R length-scale vectors d_ell[R x dim] against the same (X, y).  nlml_out[R];
 * grad_out[R x dim] or NULL.  Workspace from bogp_nlml_batched_workspace_bytes.          */
size_t bogp_nlml_batched_workspace_bytes(int64_t n, int dim, int64_t r, int want_grad);
int bogp_nlml_batched(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim,
                      const double* d_ell, int64_t r, double jitter, double* d_nlml_out,
                      double* d_grad_out, void* d_workspace, size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* BOGP_H */
