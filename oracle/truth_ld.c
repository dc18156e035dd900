/*
 * truth_ld.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * Higher-precision truth for the GP posterior and log marginal likelihood of
 * /root/reference/point_selector.py:78-98,111-120: the same formulas, evaluated in x87 extended
 * precision (long double, 64-bit significand, eps = 1.08e-19) with a Cholesky factorisation instead of
 * the reference's explicit inverse.  SURVEY.md 7.3-1: sigma^2 = prior - k^T K^-1 k is a cancellation, so
 * both the reference (np.linalg.inv in fp64) and the B200 path carry an absolute error ~ cond(K) * eps;
 * the parity tests use this file to show that the B200 path is no further from the truth than the
 * reference's own arithmetic is, including the ill-conditioned cases (ell >= 1).
 *
 *   K[i][j] = exp(-0.5 * sum_k (x_ik - x_jk)^2 / ell_k^2) + jitter * (i == j)       point_selector.py:166-195
 *   mu_c    = k_c^T K^-1 y,  var_c = prior - k_c^T K^-1 k_c                          point_selector.py:90-98
 *   nlml    = 0.5 * (y^T K^-1 y + log det K + n log 2 pi)                            point_selector.py:111-120
 *
 * Inputs are the fp64 values the reference would see; every operation after that is long double.
 * Build: gcc -O2 -pthread -shared -fPIC -o _build/libtruth_ld.so truth_ld.c -lm   (oracle/Makefile)
 * Threads: plain pthreads (the image's gcc has no libgomp); rows are dealt round-robin so that the
 * triangular loops balance.  GP_TRUTH_THREADS overrides the thread count (default: online cores, <= 64).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef long double ld;

#define NB 64

typedef void (*row_fn)(long i, void* ctx);
typedef struct { row_fn fn; void* ctx; long lo, hi; int tid, nt; } job_t;

static void* job_main(void* p) {
    job_t* j = (job_t*)p;
    for (long i = j->lo + j->tid; i < j->hi; i += j->nt) j->fn(i, j->ctx);
    return NULL;
}

static int n_threads(void) {
    const char* e = getenv("GP_TRUTH_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}

/* fn(i, ctx) for i in [lo, hi), rows dealt round-robin to the threads */
static void parallel_rows(long lo, long hi, row_fn fn, void* ctx) {
    int nt = n_threads();
    if (hi - lo < 2 * nt) nt = 1;
    pthread_t th[64]; job_t jobs[64];
    for (int t = 0; t < nt; t++) {
        jobs[t] = (job_t){fn, ctx, lo, hi, t, nt};
        if (t > 0) pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    job_main(&jobs[0]);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

typedef struct { ld* a; long n, j0, jb; } chol_ctx;

static void panel_row(long i, void* p) {
    chol_ctx* c = (chol_ctx*)p; ld* a = c->a; const long n = c->n, j0 = c->j0, jb = c->jb;
    for (long j = j0; j < j0 + jb; j++) {
        ld t = a[i * n + j];
        for (long k = j0; k < j; k++) t -= a[i * n + k] * a[j * n + k];
        a[i * n + j] = t / a[j * n + j];
    }
}

static void trailing_row(long i, void* p) {
    chol_ctx* c = (chol_ctx*)p; ld* a = c->a; const long n = c->n, j0 = c->j0, jb = c->jb;
    const ld* li = a + i * n + j0;
    for (long q = j0 + jb; q <= i; q++) {
        const ld* lq = a + q * n + j0;
        ld s = 0.0L;
        for (long k = 0; k < jb; k++) s += li[k] * lq[k];
        a[i * n + q] -= s;
    }
}

/* in-place blocked right-looking Cholesky of the lower triangle, row-major n x n; returns 0 or the
 * 1-based index of the first non-positive pivot */
static int chol_ld(ld* a, long n) {
    for (long j0 = 0; j0 < n; j0 += NB) {
        const long jb = (n - j0 < NB) ? (n - j0) : NB;
        /* diagonal block */
        for (long j = j0; j < j0 + jb; j++) {
            ld s = a[j * n + j];
            for (long k = j0; k < j; k++) s -= a[j * n + k] * a[j * n + k];
            if (!(s > 0.0L)) return (int)(j + 1);
            const ld piv = sqrtl(s);
            a[j * n + j] = piv;
            for (long i = j + 1; i < j0 + jb; i++) {
                ld t = a[i * n + j];
                for (long k = j0; k < j; k++) t -= a[i * n + k] * a[j * n + k];
                a[i * n + j] = t / piv;
            }
        }
        chol_ctx cc = {a, n, j0, jb};
        /* panel below: rows i >= j0 + jb solve against the diagonal block */
        parallel_rows(j0 + jb, n, panel_row, &cc);
        /* trailing update: A[i][q] -= sum_k L[i][k] L[q][k], q <= i */
        parallel_rows(j0 + jb, n, trailing_row, &cc);
    }
    return 0;
}

/* v <- L^-1 v */
static void fwd_ld(const ld* l, long n, ld* v) {
    for (long i = 0; i < n; i++) {
        ld s = v[i];
        const ld* li = l + i * n;
        for (long k = 0; k < i; k++) s -= li[k] * v[k];
        v[i] = s / li[i];
    }
}

typedef struct {
    const double* X; const double* P; long n; int d; double jitter, cross_jitter, prior;
    ld il2[64]; ld* a; ld* z; double* mu; double* var;
} gp_ctx;

static void gram_row(long i, void* p) {
    gp_ctx* g = (gp_ctx*)p; const int d = g->d; const long n = g->n;
    for (long j = 0; j <= i; j++) {
        ld s = 0.0L;
        for (int k = 0; k < d; k++) { const ld df = (ld)g->X[i * d + k] - (ld)g->X[j * d + k]; s += df * df * g->il2[k]; }
        g->a[i * n + j] = expl(-0.5L * s) + (i == j ? (ld)g->jitter : 0.0L);
    }
}

static void cand_row(long q, void* p) {
    gp_ctx* g = (gp_ctx*)p; const int d = g->d; const long n = g->n;
    ld* v = (ld*)malloc(sizeof(ld) * (size_t)n);
    for (long i = 0; i < n; i++) {
        ld s = 0.0L;
        for (int k = 0; k < d; k++) { const ld df = (ld)g->P[q * d + k] - (ld)g->X[i * d + k]; s += df * df * g->il2[k]; }
        v[i] = expl(-0.5L * s) + ((g->cross_jitter != 0.0 && i == q) ? (ld)g->cross_jitter : 0.0L);
    }
    fwd_ld(g->a, n, v);                               /* v = L^-1 k_c */
    ld m = 0.0L, qq = 0.0L;
    for (long i = 0; i < n; i++) { m += v[i] * g->z[i]; qq += v[i] * v[i]; }
    if (g->mu) g->mu[q] = (double)m;
    if (g->var) g->var[q] = (double)((ld)g->prior - qq);
    free(v);
}

/*
 * X (n x d), y (n), ell (d <= 64), P (c x d), all row-major fp64.  jitter: added to the diagonal of K.
 * cross_jitter: added to k(x_i, p_i) (the reference's shape-equality quirk, point_selector.py:173-177), usually 0.
 * Outputs (any may be NULL): mu[c], var[c], nlml[1], logdet[1]; returns 0, or the failing pivot, or -1 (no memory).
 */
int gp_truth_ld(const double* X, const double* y, long n, int d, const double* ell, double jitter,
                const double* P, long c, double prior, double cross_jitter,
                double* mu, double* var, double* nlml, double* logdet) {
    ld* a = (ld*)malloc(sizeof(ld) * (size_t)n * (size_t)n);
    ld* z = (ld*)malloc(sizeof(ld) * (size_t)n);
    if (!a || !z) { free(a); free(z); return -1; }
    gp_ctx g;
    memset(&g, 0, sizeof(g));
    g.X = X; g.P = P; g.n = n; g.d = d; g.jitter = jitter; g.cross_jitter = cross_jitter; g.prior = prior;
    g.a = a; g.z = z; g.mu = mu; g.var = var;
    for (int k = 0; k < d && k < 64; k++) g.il2[k] = 1.0L / ((ld)ell[k] * (ld)ell[k]);
    parallel_rows(0, n, gram_row, &g);
    const int info = chol_ld(a, n);
    if (info) { free(a); free(z); return info; }
    ld ldet = 0.0L;
    for (long i = 0; i < n; i++) ldet += 2.0L * logl(a[i * n + i]);
    for (long i = 0; i < n; i++) z[i] = (ld)y[i];
    fwd_ld(a, n, z);                                  /* z = L^-1 y */
    ld yKy = 0.0L;
    for (long i = 0; i < n; i++) yKy += z[i] * z[i];
    if (logdet) *logdet = (double)ldet;
    if (nlml) *nlml = (double)(0.5L * (yKy + ldet + (ld)n * logl(2.0L * acosl(-1.0L))));
    if (mu || var) parallel_rows(0, c, cand_row, &g);
    free(a); free(z);
    return 0;
}
