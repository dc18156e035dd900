// Host-buffer entry points: what the reference-facing `PointSelector` binds (include/bogp.h, "session").
//
// One call = one step of the reference's hot path with HOST arrays in and out, exactly the data the reference's
// methods see (point_selector.py:42-102 update_surrogate, :104-163 tune_kernel, :197-207 lower_confidence_bound):
// the library owns device memory, streams, host<->device copies and -- with more than one device in the session --
// the sharding of candidates (contiguous flat-index slices) and restarts over the GPUs of the box, all
// driven from the caller's single thread (every kernel launch and copy below is asynchronous on per-device streams,
// so the devices run concurrently).  No torch, no Python objects: a process that only needs this path starts in the
// time it takes to create a CUDA context.
#include "common.cuh"
#include "fit.cuh"

#include <vector>
#include <algorithm>

namespace bogp {

struct Buf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return BOGP_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = (bytes + 255) / 256 * 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; set_error("bogp_session: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return BOGP_ERR_CUDA; }
        cap = want;
        return BOGP_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kMaxPieces = 8;

struct DevState {
    int device = 0;
    bogp_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_in[kMaxPieces] = {}, ev_out[kMaxPieces] = {}, ev_sync = nullptr;
    Buf in, fitws, acqws, cand, out, lml, res;
    // mu / sigma of the last update kept on the device for bogp_session_score
    int64_t out_begin = 0, out_count = 0; bool out_valid = false;
};

}  // namespace bogp

using namespace bogp;

struct bogp_session {
    std::vector<DevState> dev;
    int64_t last_begin = 0, last_end = 0;      // candidate range of the last update with outputs
};

namespace {

struct DeviceGuard {          // every entry point leaves the caller's current device as it found it
    int prev = 0;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

void shard(int64_t begin, int64_t end, int r, int g, int64_t* b, int64_t* e) {      // ceil-sized contiguous slices, SURVEY 8e
    const int64_t total = end - begin, per = (total + g - 1) / g;
    *b = std::min(end, begin + r * per);
    *e = std::min(end, *b + per);
}

#define S_CUDA(expr)                                                                                     \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) { set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); return BOGP_ERR_CUDA; } \
    } while (0)
#define S_TRY(expr) do { const int _rc = (expr); if (_rc) return _rc; } while (0)

// X, y, ell and (grid mode) the axes to the device: one small staging block
struct InputPtrs { double *x, *y, *ell, *axes; };
int upload_inputs(DevState& d, const double* h_x, const double* h_y, int64_t n, int dim, const double* h_ell, int64_t n_ell,
                  const double* h_axes, int64_t n_axes, InputPtrs* out) {
    const size_t total = ((size_t)n * dim + n + n_ell + n_axes + 8) * 8;
    S_TRY(d.in.ensure(total));
    double* base = static_cast<double*>(d.in.p);
    out->x = base; out->y = out->x + n * dim; out->ell = out->y + n; out->axes = out->ell + n_ell;
    S_CUDA(cudaMemcpyAsync(out->x, h_x, (size_t)n * dim * 8, cudaMemcpyHostToDevice, d.stream));
    S_CUDA(cudaMemcpyAsync(out->y, h_y, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
    S_CUDA(cudaMemcpyAsync(out->ell, h_ell, (size_t)n_ell * 8, cudaMemcpyHostToDevice, d.stream));
    if (n_axes) S_CUDA(cudaMemcpyAsync(out->axes, h_axes, (size_t)n_axes * 8, cudaMemcpyHostToDevice, d.stream));
    return BOGP_OK;
}

}  // namespace

extern "C" int bogp_session_create(const int* devices, int n_devices, bogp_session** out) {
    if (!out || n_devices < 0 || n_devices > 64 || (n_devices > 0 && !devices)) { set_error("bogp_session_create: bad argument"); return BOGP_ERR_BAD_ARG; }
    DeviceGuard guard;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("bogp_session_create: no CUDA device available (%s); libbogp has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return BOGP_ERR_CUDA;
    }
    bogp_session* s = new bogp_session();
    const int nd = n_devices == 0 ? 1 : n_devices;
    s->dev.resize(nd);
    for (int i = 0; i < nd; i++) {
        DevState& d = s->dev[i];
        d.device = n_devices == 0 ? 0 : devices[i];
        int rc = bogp_create(d.device, &d.ctx);
        if (rc) { bogp_session_destroy(s); return rc; }
        cudaError_t ce = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking);
        for (int k = 0; k < kMaxPieces && ce == cudaSuccess; k++) {
            ce = cudaEventCreateWithFlags(&d.ev_in[k], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d.ev_out[k], cudaEventDisableTiming);
        }
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d.ev_sync, cudaEventDisableTiming);
        if (ce != cudaSuccess) { set_error("bogp_session_create: %s", cudaGetErrorString(ce)); bogp_session_destroy(s); return BOGP_ERR_CUDA; }
        bogp_set_stream(d.ctx, d.stream);
        if (d.res.ensure((kMaxPieces + 2) * sizeof(bogp_result))) { bogp_session_destroy(s); return BOGP_ERR_CUDA; }
    }
    *out = s;
    return BOGP_OK;
}

extern "C" void bogp_session_destroy(bogp_session* s) {
    if (!s) return;
    DeviceGuard guard;
    for (DevState& d : s->dev) {
        if (!d.ctx) continue;
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.copy_stream) cudaStreamSynchronize(d.copy_stream);
        d.in.release(); d.fitws.release(); d.acqws.release(); d.cand.release(); d.out.release(); d.lml.release(); d.res.release();
        for (int k = 0; k < kMaxPieces; k++) { if (d.ev_in[k]) cudaEventDestroy(d.ev_in[k]); if (d.ev_out[k]) cudaEventDestroy(d.ev_out[k]); }
        if (d.ev_sync) cudaEventDestroy(d.ev_sync);
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
        bogp_destroy(d.ctx);
    }
    delete s;
}

extern "C" int bogp_session_device_count(const bogp_session* s) { return s ? (int)s->dev.size() : 0; }
extern "C" bogp_ctx* bogp_session_ctx(bogp_session* s, int i) { return (s && i >= 0 && i < (int)s->dev.size()) ? s->dev[i].ctx : nullptr; }
extern "C" int64_t bogp_session_launch_count(const bogp_session* s) {
    int64_t t = 0;
    if (s) for (const DevState& d : s->dev) t += bogp_launch_count(d.ctx);
    return t;
}
extern "C" int bogp_session_set_acquire_path(bogp_session* s, int path) {
    if (!s) { set_error("bogp_session_set_acquire_path: null session"); return BOGP_ERR_BAD_ARG; }
    for (DevState& d : s->dev) S_TRY(bogp_set_acquire_path(d.ctx, path));
    return BOGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// kernel_rbf with host arrays                                                          point_selector.py:166-195
// ---------------------------------------------------------------------------------------------------------------
extern "C" int bogp_session_kernel_matrix(bogp_session* s, const double* h_a, int64_t na, const double* h_b, int64_t nb, int dim,
                                          const double* h_ell, double jitter, double* h_k_out) {
    if (!s || !h_a || !h_b || !h_ell || !h_k_out || na <= 0 || nb <= 0 || dim <= 0 || dim > BOGP_MAX_DIM) { set_error("bogp_session_kernel_matrix: bad argument"); return BOGP_ERR_BAD_ARG; }
    DeviceGuard guard;
    DevState& d = s->dev[0];
    S_CUDA(cudaSetDevice(d.device));
    const size_t in_bytes = ((size_t)(na + nb) * dim + dim) * 8;
    S_TRY(d.in.ensure(in_bytes));
    S_TRY(d.out.ensure((size_t)na * nb * 8));
    d.out_valid = false;
    double* da = static_cast<double*>(d.in.p); double* db = da + na * dim; double* dl = db + nb * dim;
    S_CUDA(cudaMemcpyAsync(da, h_a, (size_t)na * dim * 8, cudaMemcpyHostToDevice, d.stream));
    S_CUDA(cudaMemcpyAsync(db, h_b, (size_t)nb * dim * 8, cudaMemcpyHostToDevice, d.stream));
    S_CUDA(cudaMemcpyAsync(dl, h_ell, (size_t)dim * 8, cudaMemcpyHostToDevice, d.stream));
    S_TRY(bogp_kernel_matrix(d.ctx, da, na, db, nb, dim, dl, jitter, static_cast<double*>(d.out.p), nb));
    S_CUDA(cudaMemcpyAsync(h_k_out, d.out.p, (size_t)na * nb * 8, cudaMemcpyDeviceToHost, d.stream));
    S_CUDA(cudaStreamSynchronize(d.stream));
    return BOGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// tune_kernel's table: nlml (and gradient) for R length-scale vectors                 point_selector.py:104-163
// Restarts are dealt to the devices in contiguous blocks; each device works through its block in chunks sized to
// the free device memory (the batched workspace is R * n_pad^2 * 16 B and more).
// ---------------------------------------------------------------------------------------------------------------
extern "C" int bogp_session_nlml(bogp_session* s, const double* h_x, const double* h_y, int64_t n, int dim, const double* h_ells,
                                 int64_t r, double jitter, double* h_nlml_out, double* h_grad_out) {
    if (!s || !h_x || !h_y || !h_ells || !h_nlml_out || n <= 0 || dim <= 0 || dim > BOGP_MAX_DIM || r <= 0) { set_error("bogp_session_nlml: bad argument"); return BOGP_ERR_BAD_ARG; }
    DeviceGuard guard;
    const int g = (int)std::min<int64_t>((int64_t)s->dev.size(), r);
    struct Plan { int64_t b, e, chunk; InputPtrs in; double* nl; double* gr; };
    std::vector<Plan> plan(g);
    const int want_grad = h_grad_out ? 1 : 0;
    for (int i = 0; i < g; i++) {
        DevState& d = s->dev[i]; Plan& p = plan[i];
        shard(0, r, i, g, &p.b, &p.e);
        if (p.e <= p.b) continue;
        S_CUDA(cudaSetDevice(d.device));
        const int64_t mine = p.e - p.b;
        size_t free_b = 0, total_b = 0;
        S_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = std::max<size_t>((size_t)64 << 20, std::min<size_t>((free_b + d.lml.cap) / 2, (size_t)24 << 30));
        int64_t chunk = mine;
        while (chunk > 1 && bogp_nlml_batched_workspace_bytes(n, dim, chunk, want_grad) > budget) chunk = (chunk + 1) / 2;
        p.chunk = chunk;
        // inputs: X, y, this device's length-scale rows; outputs behind them
        const size_t in_d = (size_t)n * dim + n + (size_t)mine * dim + (size_t)mine * (1 + (want_grad ? dim : 0)) + 8;
        S_TRY(d.in.ensure(in_d * 8));
        double* base = static_cast<double*>(d.in.p);
        p.in.x = base; p.in.y = p.in.x + n * dim; p.in.ell = p.in.y + n; p.nl = p.in.ell + mine * dim; p.gr = p.nl + mine;
        S_CUDA(cudaMemcpyAsync(p.in.x, h_x, (size_t)n * dim * 8, cudaMemcpyHostToDevice, d.stream));
        S_CUDA(cudaMemcpyAsync(p.in.y, h_y, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
        S_CUDA(cudaMemcpyAsync(p.in.ell, h_ells + p.b * dim, (size_t)mine * dim * 8, cudaMemcpyHostToDevice, d.stream));
        S_TRY(d.lml.ensure(std::max<size_t>(256, bogp_nlml_batched_workspace_bytes(n, dim, chunk, want_grad))));
        for (int64_t c0 = 0; c0 < mine; c0 += chunk) {
            const int64_t cur = std::min(chunk, mine - c0);
            S_TRY(bogp_nlml_batched(d.ctx, p.in.x, p.in.y, n, dim, p.in.ell + c0 * dim, cur, jitter, p.nl + c0,
                                    want_grad ? p.gr + c0 * dim : nullptr, d.lml.p, d.lml.cap));
        }
    }
    for (int i = 0; i < g; i++) {
        DevState& d = s->dev[i]; Plan& p = plan[i];
        if (p.e <= p.b) continue;
        S_CUDA(cudaSetDevice(d.device));
        S_CUDA(cudaMemcpyAsync(h_nlml_out + p.b, p.nl, (size_t)(p.e - p.b) * 8, cudaMemcpyDeviceToHost, d.stream));
        if (want_grad) S_CUDA(cudaMemcpyAsync(h_grad_out + p.b * dim, p.gr, (size_t)(p.e - p.b) * dim * 8, cudaMemcpyDeviceToHost, d.stream));
        S_CUDA(cudaStreamSynchronize(d.stream));
    }
    return BOGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// update_surrogate: fit + posterior (+ acquisition and arg-max in the same sweep)     point_selector.py:42-102
// ---------------------------------------------------------------------------------------------------------------
extern "C" int bogp_session_update(bogp_session* s, const double* h_x, const double* h_y, int64_t n, int dim, const double* h_ell,
                                   double jitter, const bogp_host_candidates* cand, int64_t c_begin, int64_t c_end, double prior_diag,
                                   int kind, double explore, double f_best, double* h_mu_out, double* h_sigma_out, double* h_acq_out,
                                   double* h_nlml_out, double* h_best_score, int64_t* h_best_index) {
    if (!s || !h_x || !h_y || !h_ell || !cand || n <= 0 || dim <= 0 || dim > BOGP_MAX_DIM || c_begin < 0 || c_end <= c_begin || c_end > cand->c_total ||
        (!cand->h_points && (!cand->h_axes || !cand->h_axis_len))) {
        set_error("bogp_session_update: bad argument"); return BOGP_ERR_BAD_ARG;
    }
    DeviceGuard guard;
    const bool grid = cand->h_points == nullptr;
    int64_t n_axes = 0;
    if (grid) for (int k = 0; k < dim; k++) { if (cand->h_axis_len[k] <= 0) { set_error("bogp_session_update: empty grid axis %d", k); return BOGP_ERR_BAD_ARG; } n_axes += cand->h_axis_len[k]; }
    const bool want_out = h_mu_out || h_sigma_out || h_acq_out;
    const int g = (int)std::min<int64_t>((int64_t)s->dev.size(), c_end - c_begin);
    struct Plan { int64_t b = 0, e = 0; int pieces = 0; int64_t piece = 0; bogp_fit* fit = nullptr; double *mu = nullptr, *sg = nullptr, *aq = nullptr; };
    std::vector<Plan> plan(s->dev.size());
    auto cleanup = [&]() { for (Plan& p : plan) { if (p.fit) bogp_fit_destroy(p.fit); p.fit = nullptr; } };
#define U_TRY(expr) do { const int _rc = (expr); if (_rc) { for (int _i = 0; _i < g; _i++) { cudaSetDevice(s->dev[_i].device); cudaStreamSynchronize(s->dev[_i].stream); cudaStreamSynchronize(s->dev[_i].copy_stream); } cleanup(); return _rc; } } while (0)
#define U_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); U_TRY(BOGP_ERR_CUDA); } } while (0)
    for (DevState& d : s->dev) d.out_valid = false;

    // phase A: inputs + fit on every device (replicated; every accumulation order is fixed, so L is bit-identical)
    std::vector<InputPtrs> in(g);
    for (int i = 0; i < g; i++) {
        DevState& d = s->dev[i]; Plan& p = plan[i];
        shard(c_begin, c_end, i, g, &p.b, &p.e);
        U_CUDA(cudaSetDevice(d.device));
        U_TRY(upload_inputs(d, h_x, h_y, n, dim, h_ell, dim, grid ? cand->h_axes : nullptr, n_axes, &in[i]));
        const size_t fb = bogp_fit_workspace_bytes(n, dim);
        U_TRY(d.fitws.ensure(fb));
        U_TRY(bogp_fit_enqueue(d.ctx, in[i].x, in[i].y, n, dim, in[i].ell, jitter, d.fitws.p, d.fitws.cap, &p.fit));
    }
    // phase B: candidates to the device piece by piece (copy stream), sweep per piece (compute stream)
    for (int i = 0; i < g; i++) {
        DevState& d = s->dev[i]; Plan& p = plan[i];
        const int64_t mine = p.e - p.b;
        if (mine <= 0) continue;
        U_CUDA(cudaSetDevice(d.device));
        const int64_t n_pad = bogp_fit_n_pad(p.fit);
        int64_t chunk = std::min<int64_t>(65536, std::max<int64_t>(8192, ((int64_t)2 << 30) / (n_pad * 8)));     // k_* panel of about 2 GB
        chunk = std::max<int64_t>(64, std::min<int64_t>(chunk, (mine + 63) / 64 * 64));
        const size_t ab = bogp_acquire_workspace_bytes(p.fit, chunk);
        U_TRY(d.acqws.ensure(ab));
        if (want_out) {
            U_TRY(d.out.ensure((size_t)mine * 3 * 8));
            p.mu = static_cast<double*>(d.out.p); p.sg = p.mu + mine; p.aq = p.sg + mine;
        }
        // pieces: explicit candidates are copied in up to kMaxPieces blocks so that the copy of block k+1 runs under the
        // sweep of block k; a grid needs no copy and is swept in one piece unless outputs are streamed back
        p.pieces = (int)std::min<int64_t>(kMaxPieces, std::max<int64_t>(1, mine / (4 * chunk)));
        if (grid && !want_out) p.pieces = 1;
        p.piece = ((mine + p.pieces - 1) / p.pieces + 63) / 64 * 64;
        p.pieces = (int)((mine + p.piece - 1) / p.piece);
        bogp_candidates cd{};
        cd.c_total = cand->c_total; cd.cross_jitter = cand->cross_jitter;
        int32_t lens[BOGP_MAX_DIM] = {};
        if (grid) {
            for (int k = 0; k < dim; k++) lens[k] = cand->h_axis_len[k];
            cd.d_axes = in[i].axes; cd.h_axis_len = lens;
        } else {
            U_TRY(d.cand.ensure((size_t)mine * dim * 8));
            // the sweep indexes the candidate array with GLOBAL flat indices: hand it the address row 0 would have
            cd.d_points = static_cast<const double*>(d.cand.p) - p.b * dim;
        }
        bogp_result* res = static_cast<bogp_result*>(d.res.p);
        // grid sweeps split over devices / pieces: every part screens against the same, global, seed floor
        U_TRY(bogp_set_global_seed(d.ctx, (grid && !want_out && (g > 1 || p.pieces > 1)) ? 1 : 0));
        for (int k = 0; k < p.pieces; k++) {
            const int64_t b = p.b + k * p.piece, e = std::min(p.e, b + p.piece);
            if (!grid) {
                U_CUDA(cudaMemcpyAsync(static_cast<double*>(d.cand.p) + (b - p.b) * dim, cand->h_points + b * dim, (size_t)(e - b) * dim * 8,
                                       cudaMemcpyHostToDevice, d.copy_stream));
                U_CUDA(cudaEventRecord(d.ev_in[k], d.copy_stream));
                U_CUDA(cudaStreamWaitEvent(d.stream, d.ev_in[k], 0));
            }
            const int64_t o = b - p.b;
            U_TRY(bogp_acquire_async(d.ctx, p.fit, &cd, b, e, kind, explore, f_best, prior_diag, want_out ? p.mu + o : nullptr,
                                     want_out ? p.sg + o : nullptr, want_out ? p.aq + o : nullptr, d.acqws.p, d.acqws.cap, res + k));
            U_CUDA(cudaEventRecord(d.ev_out[k], d.stream));
        }
        U_TRY(bogp_set_global_seed(d.ctx, 0));
        U_TRY(bogp_reduce_results(d.ctx, res, p.pieces, res + kMaxPieces, nullptr, nullptr));
    }
    // phase C: outputs back piece by piece (the copy of piece k overlaps the sweep of piece k+1), status, winner
    for (int i = 0; i < g; i++) {
        DevState& d = s->dev[i]; Plan& p = plan[i];
        const int64_t mine = p.e - p.b;
        if (mine <= 0 || !want_out) continue;
        U_CUDA(cudaSetDevice(d.device));
        for (int k = 0; k < p.pieces; k++) {
            const int64_t b = p.b + k * p.piece, e = std::min(p.e, b + p.piece), o = b - p.b, ho = b - c_begin;
            U_CUDA(cudaStreamWaitEvent(d.copy_stream, d.ev_out[k], 0));
            if (h_mu_out) U_CUDA(cudaMemcpyAsync(h_mu_out + ho, p.mu + o, (size_t)(e - b) * 8, cudaMemcpyDeviceToHost, d.copy_stream));
            if (h_sigma_out) U_CUDA(cudaMemcpyAsync(h_sigma_out + ho, p.sg + o, (size_t)(e - b) * 8, cudaMemcpyDeviceToHost, d.copy_stream));
            if (h_acq_out) U_CUDA(cudaMemcpyAsync(h_acq_out + ho, p.aq + o, (size_t)(e - b) * 8, cudaMemcpyDeviceToHost, d.copy_stream));
        }
    }
    double best_s = -INFINITY; int64_t best_i = INT64_MAX; int any_nan = 0; int status = BOGP_OK;
    for (int i = 0; i < g; i++) {
        DevState& d = s->dev[i]; Plan& p = plan[i];
        U_CUDA(cudaSetDevice(d.device));
        double nl = 0.0;
        const int frc = bogp_fit_status(p.fit, &nl);          // synchronises the compute stream
        U_CUDA(cudaStreamSynchronize(d.copy_stream));
        if (frc) { status = frc; continue; }
        if (i == 0 && h_nlml_out) *h_nlml_out = nl;
        if (p.e > p.b) {
            bogp_result h;
            U_CUDA(cudaMemcpy(&h, static_cast<bogp_result*>(d.res.p) + kMaxPieces, sizeof(h), cudaMemcpyDeviceToHost));
            any_nan |= h.nan_flag;
            if (h.score > best_s || (h.score == best_s && h.index < best_i)) { best_s = h.score; best_i = h.index; }
            if (want_out) { d.out_begin = p.b; d.out_count = p.e - p.b; d.out_valid = true; }
        }
    }
    cleanup();
#undef U_TRY
#undef U_CUDA
    if (status) { for (DevState& d : s->dev) d.out_valid = false; return status; }
    s->last_begin = c_begin; s->last_end = c_end;
    if (any_nan) { set_error("bogp_session_update: NaN acquisition value (reference raises IndexError, point_selector.py:207)"); return BOGP_ERR_NAN_SCORE; }
    if (h_best_score) *h_best_score = best_s;
    if (h_best_index) *h_best_index = best_i;
    return BOGP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// lower_confidence_bound / EI on the mu, sigma the last update left on the device(s)   point_selector.py:197-207
// With h_mu / h_sigma given (the caller changed mean_func / cov_func) they are uploaded first (device 0).
// ---------------------------------------------------------------------------------------------------------------
extern "C" int bogp_session_score(bogp_session* s, const double* h_mu, const double* h_sigma, int64_t c, int kind, double explore,
                                  double f_best, double* h_acq_out, double* h_best_score, int64_t* h_best_index) {
    if (!s || c <= 0 || (kind != BOGP_ACQ_LCB && kind != BOGP_ACQ_EI) || ((h_mu == nullptr) != (h_sigma == nullptr))) { set_error("bogp_session_score: bad argument"); return BOGP_ERR_BAD_ARG; }
    DeviceGuard guard;
    if (h_mu) {
        DevState& d = s->dev[0];
        S_CUDA(cudaSetDevice(d.device));
        S_TRY(d.out.ensure((size_t)c * 3 * 8));
        for (DevState& o : s->dev) o.out_valid = false;
        double* mu = static_cast<double*>(d.out.p); double* sg = mu + c;
        S_CUDA(cudaMemcpyAsync(mu, h_mu, (size_t)c * 8, cudaMemcpyHostToDevice, d.stream));
        S_CUDA(cudaMemcpyAsync(sg, h_sigma, (size_t)c * 8, cudaMemcpyHostToDevice, d.stream));
        d.out_begin = 0; d.out_count = c; d.out_valid = true;
        s->last_begin = 0; s->last_end = c;
    }
    if (s->last_end - s->last_begin != c) { set_error("bogp_session_score: %lld candidates asked, the device holds mu / sigma of %lld", (long long)c, (long long)(s->last_end - s->last_begin)); return BOGP_ERR_BAD_ARG; }
    int64_t covered = 0;
    for (DevState& d : s->dev) if (d.out_valid) covered += d.out_count;
    if (covered != c) { set_error("bogp_session_score: no posterior on the device (call bogp_session_update with outputs first)"); return BOGP_ERR_BAD_ARG; }
    for (DevState& d : s->dev) {
        if (!d.out_valid) continue;
        S_CUDA(cudaSetDevice(d.device));
        double* mu = static_cast<double*>(d.out.p); double* sg = mu + d.out_count; double* aq = sg + d.out_count;
        S_TRY(bogp_score_argmax_async(d.ctx, mu, sg, d.out_count, d.out_begin, kind, explore, f_best, h_acq_out ? aq : nullptr,
                                      static_cast<bogp_result*>(d.res.p) + kMaxPieces + 1));
        if (h_acq_out) S_CUDA(cudaMemcpyAsync(h_acq_out + (d.out_begin - s->last_begin), aq, (size_t)d.out_count * 8, cudaMemcpyDeviceToHost, d.stream));
    }
    double best_s = -INFINITY; int64_t best_i = INT64_MAX; int any_nan = 0;
    for (DevState& d : s->dev) {
        if (!d.out_valid) continue;
        S_CUDA(cudaSetDevice(d.device));
        bogp_result h;
        S_CUDA(cudaMemcpyAsync(&h, static_cast<bogp_result*>(d.res.p) + kMaxPieces + 1, sizeof(h), cudaMemcpyDeviceToHost, d.stream));
        S_CUDA(cudaStreamSynchronize(d.stream));
        any_nan |= h.nan_flag;
        if (h.score > best_s || (h.score == best_s && h.index < best_i)) { best_s = h.score; best_i = h.index; }
    }
    if (any_nan) { set_error("bogp_session_score: NaN acquisition value (reference raises IndexError, point_selector.py:207)"); return BOGP_ERR_NAN_SCORE; }
    if (h_best_score) *h_best_score = best_s;
    if (h_best_index) *h_best_index = best_i;
    return BOGP_OK;
}
