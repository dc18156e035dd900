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

/* number of CUDA devices visible to this process (0 and BOGP_ERR_CUDA without a driver) */
int  bogp_device_count(int* out);

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
/* Arg-max-only sweeps (no mu / sigma / acquisition outputs requested) are screened by default: the posterior mean of
 * every candidate gives an exact upper bound of its score (both acquisitions grow with sigma, and sigma^2 <= prior);
 * candidates whose bound is below the best exact score so far are dropped, the rest is scored exactly.  The returned
 * (score, index) is the one of the unscreened sweep (csrc/acquire.cu, screen_kernel).  0 switches it off.          */
int bogp_set_screening(bogp_ctx* ctx, int enable);
int bogp_get_screening(const bogp_ctx* ctx);
/* INT8 path: run a sweep as ONE persistent fused kernel (enable = 1) -- grid index -> k_* digits -> tcgen05 product ->
 * sigma^2, mu -> acquisition -> max-loc, the k_* panel living only in an L2-sized ring (csrc/acquire_fused.cu; no k_*
 * traffic to HBM, a 60 MB workspace instead of 2 GB) -- or (enable = 0, the default: measured ~6 % faster at N = 4096
 * because the panel kernel already overlaps the product on a second stream) as per-chunk panel / product / finalize /
 * merge kernels.  Outputs are bit-identical either way.
 * `group` = candidate tiles that share L2 (work group of the fused kernel, CTA ordering of the separate tri-GEMM), 0 = automatic
 * (~32 MB of panel digits; measured 8 … 64 at N = 4096: no setting beats it, tools/group_probe.py).  Replaces point_selector.py:81,90-98,204-207. */
int bogp_set_fused(bogp_ctx* ctx, int enable, int group);
int bogp_get_fused(const bogp_ctx* ctx);
/* Screened sweeps of ONE SHARD of a sharded arg-max (select_parameters.py:282-294 spread over several GPUs): with enable = 1
 * the strided seed sample that is scored before the screen starts spans the WHOLE candidate set [0, c_total) instead of the
 * shard [c_begin, c_end), so every shard screens against the same, global, floor -- a shard whose own landscape is flat (far
 * from all measurements: mu -> 0, sigma -> sigma_max) would otherwise have to score everything exactly.  The record the sweep
 * returns may then point at a seed candidate OUTSIDE the shard; the winner over all shards is unchanged.  Default 0.      */
int bogp_set_global_seed(bogp_ctx* ctx, int enable);
/* candidates that went through the screen / that survived it since the last reset (synchronises the stream) */
int bogp_screen_stats(bogp_ctx* ctx, int64_t* h_screened, int64_t* h_survived, int reset);
/* Measurement aid: when enabled, each kernel of the acquisition sweep is bracketed by CUDA
 * events on the launching stream (this serialises the stream; never enable it inside a timed
 * region).  kernel_id: 0 panel, 1 tri-GEMM, 2 finalize, 3 merge.                            */
int bogp_profile(bogp_ctx* ctx, int enable);
int bogp_profile_read(const bogp_ctx* ctx, int kernel_id, double* ms_total, int64_t* launches);

/* Measurement aid (bench.py's roofline denominators; never on the product path): the issue-rate peak of one
 * pipe, measured on this device with CUDA events: burst = best of 5 launches (a few ms each) after a warm-up;
 * sustained (optional) = average over back-to-back launches for sustain_seconds, i.e. under the power cap.
 *   BOGP_PEAK_I8_UMMA   back-to-back tcgen05.mma kind::i8 (M=128,N=256,K=32) on every SM -> int8 TOP/s
 *   BOGP_PEAK_F64_DMMA  independent DMMA.8x8x4 chains, 16 warps per SM                    -> fp64 TFLOP/s   */
#define BOGP_PEAK_I8_UMMA       0
#define BOGP_PEAK_F64_DMMA      1
int bogp_measure_peak(bogp_ctx* ctx, int kind, double sustain_seconds, double* h_burst_tera, double* h_sustained_tera);

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

/* Winner of a sweep as it lives on the DEVICE (24 bytes): what ranks exchange in the multi-GPU mode (one
 * all_gather of these records over NCCL, then bogp_reduce_results) -- no host round trip per rank.            */
typedef struct bogp_result {
    double  score;        /* best acquisition value, -inf if no candidate was scored                          */
    int64_t index;        /* its global flat index (INT64_MAX if none)                                         */
    int32_t nan_flag;     /* 1 if any acquisition value was NaN (reference: IndexError, point_selector.py:207) */
    int32_t reserved;
} bogp_result;

size_t bogp_acquire_workspace_bytes(const bogp_fit* fit, int64_t max_chunk);
int bogp_acquire(bogp_ctx* ctx, const bogp_fit* fit, const bogp_candidates* cand,
                 int64_t c_begin, int64_t c_end, int kind, double explore, double f_best,
                 double prior_diag, double* d_mu_out, double* d_sigma_out, double* d_acq_out,
                 void* d_workspace, size_t workspace_bytes,
                 double* h_best_score, int64_t* h_best_index);

/* The same sweep without any host synchronisation: the winner is left in the caller's device record. */
int bogp_acquire_async(bogp_ctx* ctx, const bogp_fit* fit, const bogp_candidates* cand,
                       int64_t c_begin, int64_t c_end, int kind, double explore, double f_best,
                       double prior_diag, double* d_mu_out, double* d_sigma_out, double* d_acq_out,
                       void* d_workspace, size_t workspace_bytes, bogp_result* d_result);
/* Fold `count` device records (several sweeps, or the all_gather of every rank's record) into one with the
 * reference's tie rule (largest score, then smallest flat index); d_out may be NULL (internal scratch).  With
 * host pointers given it synchronises and returns the winner (BOGP_ERR_NAN_SCORE if any record carries the flag). */
int bogp_reduce_results(bogp_ctx* ctx, const bogp_result* d_results, int count, bogp_result* d_out,
                        double* h_best_score, int64_t* h_best_index);

/* acquisition + arg-max only, on mu/sigma already on the device
 * (lower_confidence_bound(), point_selector.py:197-207).                                */
int bogp_score_argmax(bogp_ctx* ctx, const double* d_mu, const double* d_sigma, int64_t c,
                      int kind, double explore, double f_best, double* d_acq_out,
                      double* h_best_score, int64_t* h_best_index);

/* same, asynchronous, with the flat index of element 0 given (a shard of a larger candidate set) */
int bogp_score_argmax_async(bogp_ctx* ctx, const double* d_mu, const double* d_sigma, int64_t c, int64_t index_offset,
                            int kind, double explore, double f_best, double* d_acq_out, bogp_result* d_result);

/* ---- K3: batched negative log marginal likelihood (+ gradient) -- point_selector.py:111-138
 * R length-scale vectors d_ell[R x dim] against the same (X, y).  nlml_out[R];
 * grad_out[R x dim] or NULL.  Workspace from bogp_nlml_batched_workspace_bytes.          */
size_t bogp_nlml_batched_workspace_bytes(int64_t n, int dim, int64_t r, int want_grad);
int bogp_nlml_batched(bogp_ctx* ctx, const double* d_x, const double* d_y, int64_t n, int dim,
                      const double* d_ell, int64_t r, double jitter, double* d_nlml_out,
                      double* d_grad_out, void* d_workspace, size_t workspace_bytes);

/* ==== Host-buffer "session" API: what the reference-facing PointSelector binds ====================================
 * Every pointer below is a HOST pointer (pageable memory is fine): the arrays the reference's methods work on.  A session
 * owns one context, streams and grow-only device buffers per device; with several devices it shards candidates
 * (contiguous flat-index slices) and restarts over them from the caller's single thread and reduces the winners with
 * the reference's tie rule.  Calls return when the host outputs are complete.                                        */
typedef struct bogp_session bogp_session;
typedef struct bogp_host_candidates {
    const double*  h_points;      /* explicit: c_total x dim row-major (predicted_pts, select_parameters.py:279), or NULL  */
    const double*  h_axes;        /* grid: the dim axes back to back (select_parameters.py:273-274)                        */
    const int32_t* h_axis_len;    /* grid: dim lengths                                                                      */
    int64_t        c_total;
    double         cross_jitter;  /* see bogp_candidates                                                                    */
} bogp_host_candidates;

int  bogp_session_create(const int* devices, int n_devices, bogp_session** out);   /* n_devices == 0: device 0           */
void bogp_session_destroy(bogp_session* s);
int  bogp_session_device_count(const bogp_session* s);
bogp_ctx* bogp_session_ctx(bogp_session* s, int i);                                 /* context of the i-th device        */
int64_t bogp_session_launch_count(const bogp_session* s);
int  bogp_session_set_acquire_path(bogp_session* s, int path);
/* kernel_rbf(x1, x2) with host arrays                                                     point_selector.py:166-195 */
int  bogp_session_kernel_matrix(bogp_session* s, const double* h_a, int64_t na, const double* h_b, int64_t nb, int dim,
                                const double* h_ell, double jitter, double* h_k_out);
/* The table of tune_kernel: nlml (and optionally its gradient) of r length-scale vectors h_ells[r x dim]; restarts are
 * dealt to the devices in contiguous blocks and chunked to the free device memory.        point_selector.py:104-163 */
int  bogp_session_nlml(bogp_session* s, const double* h_x, const double* h_y, int64_t n, int dim, const double* h_ells,
                       int64_t r, double jitter, double* h_nlml_out, double* h_grad_out);
/* update_surrogate: K = k(X,X) + jitter I, Cholesky, posterior mean / sigma of candidates [c_begin, c_end), the
 * acquisition `kind` and its first arg-max (GLOBAL flat index) in the same sweep.  Outputs of length c_end - c_begin,
 * each optional; with none requested the sweep only returns the winner.  h_nlml_out: nlml of this fit.
 * Errors: BOGP_ERR_NOT_POSDEF (reference: LinAlgError), BOGP_ERR_NAN_SCORE (reference: IndexError).  point_selector.py:42-102 */
int  bogp_session_update(bogp_session* s, const double* h_x, const double* h_y, int64_t n, int dim, const double* h_ell,
                         double jitter, const bogp_host_candidates* cand, int64_t c_begin, int64_t c_end, double prior_diag,
                         int kind, double explore, double f_best, double* h_mu_out, double* h_sigma_out, double* h_acq_out,
                         double* h_nlml_out, double* h_best_score, int64_t* h_best_index);
/* lower_confidence_bound(explore) / EI on the mu, sigma the last update left on the device(s) (h_mu == h_sigma == NULL),
 * or on host arrays given here (the caller changed mean_func / cov_func).  c = number of candidates.  point_selector.py:197-207 */
int  bogp_session_score(bogp_session* s, const double* h_mu, const double* h_sigma, int64_t c, int kind, double explore,
                        double f_best, double* h_acq_out, double* h_best_score, int64_t* h_best_index);

#ifdef __cplusplus
}
#endif
#endif /* BOGP_H */
