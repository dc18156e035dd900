"""Host side of the B200 GP engine: device memory and streams via torch, arithmetic via libbogp.

`GPEngine` is the thin Python layer between the reference-facing `PointSelector` drop-in
(`point_selector.py`) and the C ABI (`include/bogp.h`).  torch is plumbing only: it owns the
HBM buffers and the CUDA stream; every number is produced by the hand-written sm_100a
kernels of `csrc/`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ACQ_EI, ACQ_LCB, BogpError, Candidates

from .session import JITTER_LML, JITTER_POSTERIOR, PRIOR_DIAG  # noqa: F401  (the reference's jitters, defined once)


@dataclass
class CandidateGrid:
    """Row-major Cartesian product of `axes`, axis 0 slowest (select_parameters.py:273-279).
    Never materialised: the kernels turn a flat index into coordinates on the fly."""
    axes: Sequence[np.ndarray]

    @property
    def shape(self):
        return [len(a) for a in self.axes]

    @property
    def size(self):
        return int(np.prod([len(a) for a in self.axes], dtype=np.int64))

    @property
    def dim(self):
        return len(self.axes)


@dataclass
class AcquireResult:
    best_score: Optional[float]
    best_index: Optional[int]
    mu: Optional[torch.Tensor] = None
    sigma: Optional[torch.Tensor] = None
    acq: Optional[torch.Tensor] = None
    record: Optional[torch.Tensor] = None     # sync=False: the 24-byte device record (struct bogp_result) of the winner


class GPFit:
    """Fitted surrogate: Cholesky factor, W = L^-1 (packed), alpha -- all resident in HBM."""

    def __init__(self, engine, handle, workspace, n, dim, nlml, ell, jitter):
        self.engine, self._h, self._ws = engine, handle, workspace
        self.n, self.dim, self.nlml, self.ell, self.jitter = n, dim, nlml, ell, jitter

    @property
    def n_pad(self):
        return int(self.engine.lib.bogp_fit_n_pad(self._h))

    @property
    def logdet(self):
        return float(self.engine.lib.bogp_fit_logdet(self._h))

    def _view(self, ptr, numel):
        base = self._ws.data_ptr()
        off = ptr - base
        return self._ws[off:off + numel * 8].view(torch.float64)

    def chol(self):
        """L as an (n_pad, n_pad) tensor view (lower triangle valid)."""
        npad = self.n_pad
        return self._view(self.engine.lib.bogp_fit_chol(self._h), npad * npad).view(npad, npad)

    def linv(self):
        npad = self.n_pad
        return self._view(self.engine.lib.bogp_fit_linv(self._h), npad * npad).view(npad, npad)

    def alpha(self):
        return self._view(self.engine.lib.bogp_fit_alpha(self._h), self.n_pad)[: self.n]

    def close(self):
        if self._h:
            self.engine.lib.bogp_fit_destroy(self._h)
            self._h = None
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GPEngine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("bayesian_optimisation_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        _lib.check(self.lib.bogp_create(device, C.byref(h)))
        self._ctx = h
        self._acq_ws = None
        self._sync_stream()

    # ------------------------------------------------------------------ plumbing
    def _sync_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.bogp_set_stream(self._ctx, C.c_void_p(s)))

    def to_device(self, a, dtype=torch.float64):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a)).to(device=self.device, dtype=dtype)

    def prefetch_to_device(self, a: np.ndarray):
        """Start a host -> device copy on a side stream (asynchronous for pinned host memory) and return a handle for
        `prefetched()`; work enqueued on the current stream meanwhile overlaps the copy."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        src = torch.from_numpy(a)
        dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
        self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))     # dst may reuse memory still in use there
        with torch.cuda.stream(self._copy_stream):
            dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return dst, done, src

    def prefetched(self, handle) -> torch.Tensor:
        """The device tensor of a `prefetch_to_device` handle, ordered after its copy on the current stream."""
        dst, done, _src = handle
        torch.cuda.current_stream(self.device).wait_event(done)
        dst.record_stream(torch.cuda.current_stream(self.device))
        return dst

    def to_host(self, t: torch.Tensor) -> np.ndarray:
        """Device tensor -> numpy array through pinned host memory (torch's caching host allocator);
        the array owns its buffer, which returns to the cache when the array is collected."""
        host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        host.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    @property
    def launches(self) -> int:
        return int(self.lib.bogp_launch_count(self._ctx))

    @property
    def sm_count(self) -> int:
        return int(self.lib.bogp_sm_count(self._ctx))

    @property
    def acquire_path(self) -> str:
        return "i8" if self.lib.bogp_get_acquire_path(self._ctx) == _lib.PATH_INT8_TCGEN05 else "fp64"

    def set_acquire_path(self, path: str):
        """"fp64": DMMA on the FP64 pipe; "i8": exact digit slices on tcgen05 kind::i8 (include/bogp.h)."""
        code = {"fp64": _lib.PATH_FP64_DMMA, "dmma": _lib.PATH_FP64_DMMA, "i8": _lib.PATH_INT8_TCGEN05, "int8": _lib.PATH_INT8_TCGEN05}[path]
        _lib.check(self.lib.bogp_set_acquire_path(self._ctx, code))

    def set_screening(self, enable: bool):
        """Arg-max-only sweeps: screen by the posterior-mean bound and score only the survivors exactly (default on;
        same winner as the full sweep -- include/bogp.h)."""
        _lib.check(self.lib.bogp_set_screening(self._ctx, 1 if enable else 0))

    def set_global_seed(self, enable: bool):
        """Screened sweeps of one shard of a sharded arg-max: seed the screen with a strided sample of the WHOLE candidate
        set, so that every shard screens against the same floor (include/bogp.h bogp_set_global_seed)."""
        _lib.check(self.lib.bogp_set_global_seed(self._ctx, 1 if enable else 0))

    def set_fused(self, enable: bool, group: int = 0):
        """INT8 path: one persistent fused kernel per sweep, or (default) the per-chunk panel / product / finalize / merge
        kernels; bit-identical outputs (include/bogp.h bogp_set_fused).  `group` = candidate tiles per work group, 0 = automatic."""
        _lib.check(self.lib.bogp_set_fused(self._ctx, 1 if enable else 0, int(group)))

    def screen_stats(self, reset: bool = True):
        """(candidates screened, survivors) since the last reset."""
        a, b = C.c_int64(), C.c_int64()
        self._sync_stream()
        _lib.check(self.lib.bogp_screen_stats(self._ctx, C.byref(a), C.byref(b), 1 if reset else 0))
        return int(a.value), int(b.value)

    def profile(self, enable: bool):
        """Per-kernel CUDA-event timing of the acquisition sweep (measurement aid; serialises the stream)."""
        _lib.check(self.lib.bogp_profile(self._ctx, 1 if enable else 0))

    def profile_read(self):
        out = {}
        for kid, name in enumerate(["panel", "trigemm", "finalize", "merge", "chol_diag", "chol_trsm", "chol_syrk_inner", "chol_syrk_outer"]):
            ms, n = C.c_double(), C.c_int64()
            _lib.check(self.lib.bogp_profile_read(self._ctx, kid, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, int(n.value))
        return out

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.bogp_destroy(self._ctx)
            self._ctx = None

    # ------------------------------------------------------------------ K1
    def kernel_matrix(self, a, b, ell, jitter: float = 0.0) -> torch.Tensor:
        """exp(-0.5 sum_k (a_ik-b_jk)^2/ell_k^2) (+ jitter on i == j)   point_selector.py:166-195"""
        self._sync_stream()
        da, db, dl = self.to_device(a), self.to_device(b), self.to_device(np.asarray(ell, dtype=np.float64).reshape(-1))
        na, dim = da.shape
        nb = db.shape[0]
        out = torch.empty((na, nb), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bogp_kernel_matrix(self._ctx, da.data_ptr(), na, db.data_ptr(), nb, dim, dl.data_ptr(),
                                               float(jitter), out.data_ptr(), nb))
        return out

    # ------------------------------------------------------------------ K2 + fit
    def cholesky(self, a: torch.Tensor):
        """In-place blocked Cholesky of the lower triangle of `a` (n multiple of 64).
        Returns (logdet, info)."""
        self._sync_stream()
        n = a.shape[0]
        linv = torch.zeros_like(a)
        scal = torch.zeros(1, dtype=torch.float64, device=self.device)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.bogp_cholesky(self._ctx, a.data_ptr(), n, a.stride(0), linv.data_ptr(), scal.data_ptr(), info.data_ptr()))
        return float(scal.item()), int(info.item())

    def fit(self, x, y, ell, jitter: float = JITTER_POSTERIOR) -> GPFit:
        """K = k(X,X) + jitter I; Cholesky; alpha; W = L^-1      point_selector.py:79,89-90"""
        self._sync_stream()
        dx, dy = self.to_device(x), self.to_device(np.asarray(y, dtype=np.float64).reshape(-1) if not isinstance(y, torch.Tensor) else y.reshape(-1))
        ell_np = np.asarray(ell.cpu() if isinstance(ell, torch.Tensor) else ell, dtype=np.float64).reshape(-1)
        dl = self.to_device(ell_np)
        n, dim = dx.shape
        if dl.numel() != dim:
            raise ValueError(f"{dl.numel()} length scales for {dim} features")
        if dim > _lib.BOGP_MAX_DIM:
            raise ValueError(f"{dim} features; the kernels support at most {_lib.BOGP_MAX_DIM}")
        nbytes = self.lib.bogp_fit_workspace_bytes(n, dim)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        nlml = C.c_double()
        code = self.lib.bogp_fit_create(self._ctx, dx.data_ptr(), dy.data_ptr(), n, dim, dl.data_ptr(), float(jitter),
                                        ws.data_ptr(), nbytes, C.byref(h), C.byref(nlml))
        if code == _lib.BOGP_ERR_NOT_POSDEF:
            raise np.linalg.LinAlgError(self.lib.bogp_last_error().decode())
        _lib.check(code)
        return GPFit(self, h, ws, n, dim, nlml.value, ell_np, float(jitter))

    # ------------------------------------------------------------------ K4
    def acquire(self, fit: GPFit, candidates, c_begin: int = 0, c_end: Optional[int] = None, kind: int = ACQ_LCB,
                explore: float = 4.0, f_best: float = 0.0, prior_diag: float = PRIOR_DIAG, outputs: bool = False,
                chunk: Optional[int] = None, cross_jitter: float = 0.0, sync: bool = True) -> AcquireResult:
        """Score flat candidate indices [c_begin, c_end) and return the best (score, index).

        `candidates`: CandidateGrid, or an explicit (C, d) array (numpy -> copied to HBM, or a
        CUDA tensor used in place).  sync=False only enqueues the sweep: the winner stays on the device in
        `result.record` (for `sharding.allreduce_maxloc_device` / `reduce_records`)."""
        self._sync_stream()
        keep = []
        cd = Candidates()
        cd.cross_jitter = float(cross_jitter)
        if isinstance(candidates, CandidateGrid):
            if candidates.dim != fit.dim:
                raise ValueError("candidate grid dimension does not match the fit")
            axes = self.to_device(np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in candidates.axes]))
            lens = (C.c_int32 * candidates.dim)(*candidates.shape)
            keep += [axes, lens]
            cd.d_points, cd.d_axes, cd.h_axis_len, cd.c_total = None, axes.data_ptr(), lens, candidates.size
        else:
            pts = self.to_device(candidates)
            if pts.ndim != 2 or pts.shape[1] != fit.dim:
                raise ValueError("candidates must be (C, d)")
            keep.append(pts)
            cd.d_points, cd.d_axes, cd.h_axis_len, cd.c_total = pts.data_ptr(), None, None, pts.shape[0]
        total = int(cd.c_total)
        c_end = total if c_end is None else int(c_end)
        count = c_end - int(c_begin)
        if chunk is None:      # candidates per kernel chunk: a k_* panel of about 2 GB (fewer, longer launches: shorter tails)
            chunk = min(65536, max(8192, (2 << 30) // (fit.n_pad * 8)))
        chunk = max(64, min(int(chunk), (count + 63) // 64 * 64))
        need = self.lib.bogp_acquire_workspace_bytes(fit._h, chunk)
        if self._acq_ws is None or self._acq_ws.numel() < need:
            self._acq_ws = None
            self._acq_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        mu = sigma = acq = None
        if outputs:
            mu = torch.empty(count, dtype=torch.float64, device=self.device)
            sigma = torch.empty_like(mu)
            acq = torch.empty_like(mu)
        if not sync:
            rec = torch.empty(24, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.bogp_acquire_async(self._ctx, fit._h, C.byref(cd), int(c_begin), c_end, int(kind), float(explore),
                                                   float(f_best), float(prior_diag),
                                                   mu.data_ptr() if outputs else None, sigma.data_ptr() if outputs else None,
                                                   acq.data_ptr() if outputs else None,
                                                   self._acq_ws.data_ptr(), need, rec.data_ptr()))
            self._keepalive = keep        # the sweep is still running: its inputs must outlive this call
            return AcquireResult(None, None, mu, sigma, acq, rec)
        bs, bi = C.c_double(), C.c_int64()
        code = self.lib.bogp_acquire(self._ctx, fit._h, C.byref(cd), int(c_begin), c_end, int(kind), float(explore),
                                     float(f_best), float(prior_diag),
                                     mu.data_ptr() if outputs else None, sigma.data_ptr() if outputs else None,
                                     acq.data_ptr() if outputs else None,
                                     self._acq_ws.data_ptr(), need, C.byref(bs), C.byref(bi))
        if code == _lib.BOGP_ERR_NAN_SCORE:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0 (NaN acquisition value)")
        _lib.check(code)
        del keep
        return AcquireResult(bs.value, int(bi.value), mu, sigma, acq)

    def empty_record(self) -> torch.Tensor:
        """The record of a rank that scored nothing: (-inf, no index, no NaN)."""
        import struct
        raw = struct.pack("<dqii", float("-inf"), (1 << 63) - 1, 0, 0)
        return torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(self.device)

    def reduce_records(self, records: torch.Tensor, count: int):
        """Fold `count` gathered 24-byte records on the device and read the winner (ONE host read).  IndexError if any
        record carries the NaN flag (point_selector.py:207)."""
        self._sync_stream()
        bs, bi = C.c_double(), C.c_int64()
        code = self.lib.bogp_reduce_results(self._ctx, records.data_ptr(), int(count), None, C.byref(bs), C.byref(bi))
        if code == _lib.BOGP_ERR_NAN_SCORE:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0 (NaN acquisition value)")
        _lib.check(code)
        return bs.value, int(bi.value)

    def measure_peak(self, kind: str, sustain_seconds: float = 0.0):
        """Issue-rate peak of a pipe on this device (measurement aid, csrc/peaks.cu): "i8" -> int8 TOP/s of back-to-back
        tcgen05.mma kind::i8, "fp64" -> TFLOP/s of DMMA.8x8x4.  Returns (burst, sustained): best of 5 short launches,
        and the average over `sustain_seconds` of back-to-back launches (None if 0)."""
        self._sync_stream()
        burst, sus = C.c_double(), C.c_double()
        _lib.check(self.lib.bogp_measure_peak(self._ctx, {"i8": 0, "fp64": 1}[kind], float(sustain_seconds), C.byref(burst),
                                              C.byref(sus) if sustain_seconds > 0 else None))
        return burst.value, (sus.value if sustain_seconds > 0 else None)

    def score_argmax(self, mu: torch.Tensor, sigma: torch.Tensor, kind: int = ACQ_LCB, explore: float = 4.0,
                     f_best: float = 0.0, want_acq: bool = True) -> AcquireResult:
        """acquisition + first arg-max on device-resident mu/sigma   point_selector.py:197-207"""
        self._sync_stream()
        mu, sigma = self.to_device(mu).reshape(-1), self.to_device(sigma).reshape(-1)
        acq = torch.empty_like(mu) if want_acq else None
        bs, bi = C.c_double(), C.c_int64()
        code = self.lib.bogp_score_argmax(self._ctx, mu.data_ptr(), sigma.data_ptr(), mu.numel(), int(kind), float(explore),
                                          float(f_best), acq.data_ptr() if want_acq else None, C.byref(bs), C.byref(bi))
        if code == _lib.BOGP_ERR_NAN_SCORE:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0 (NaN acquisition value)")
        _lib.check(code)
        return AcquireResult(bs.value, int(bi.value), mu, sigma, acq)

    # ------------------------------------------------------------------ K3
    def nlml_batched(self, x, y, ells, jitter: float = JITTER_LML, want_grad: bool = False):
        """nlml (and d nlml/d ell) for R length-scale vectors at once   point_selector.py:111-138"""
        self._sync_stream()
        dx = self.to_device(x)
        dy = self.to_device(np.asarray(y, dtype=np.float64).reshape(-1) if not isinstance(y, torch.Tensor) else y.reshape(-1))
        de = self.to_device(ells)
        n, dim = dx.shape
        if de.ndim != 2 or de.shape[1] != dim:
            raise ValueError("ells must be (R, d)")
        r = de.shape[0]
        if r == 1 and n >= 2048 and not want_grad:
            # one large system: the pipelined single-matrix factorisation (bogp_fit_create) is faster than the batched
            # driver (3.5 vs 6 ms at n = 4096); same kernels, nlml equal to rounding (the accumulation order of the
            # interleaved inverse differs).  A non-positive-definite matrix gives NaN like the batched path.
            try:
                f = self.fit(dx, dy, de[0] if isinstance(ells, torch.Tensor) else np.asarray(ells, dtype=np.float64)[0], jitter)
            except np.linalg.LinAlgError:
                return torch.full((1,), float("nan"), dtype=torch.float64, device=self.device)
            val = f.nlml
            f.close()
            return torch.tensor([val], dtype=torch.float64, device=self.device)
        out = torch.empty(r, dtype=torch.float64, device=self.device)
        grad = torch.empty((r, dim), dtype=torch.float64, device=self.device) if want_grad else None
        # The batched workspace is r * n_pad^2 * 16 B and more (2500 restarts at n = 1024: 42 GB): work through the
        # restarts in chunks that fit half of the free device memory (at most 24 GB).  Restarts are independent, so
        # the chunking changes no result.
        free_b, _ = torch.cuda.mem_get_info(self.device)
        budget = max(64 << 20, min(free_b // 2, 24 << 30))
        chunk = r
        while chunk > 1 and self.lib.bogp_nlml_batched_workspace_bytes(n, dim, chunk, 1 if want_grad else 0) > budget:
            chunk = (chunk + 1) // 2
        need = self.lib.bogp_nlml_batched_workspace_bytes(n, dim, chunk, 1 if want_grad else 0)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        for r0 in range(0, r, chunk):
            cur = min(chunk, r - r0)
            _lib.check(self.lib.bogp_nlml_batched(self._ctx, dx.data_ptr(), dy.data_ptr(), n, dim, de[r0:].data_ptr(), cur, float(jitter),
                                                  out[r0:].data_ptr(), grad[r0:].data_ptr() if want_grad else None, ws.data_ptr(), ws.numel()))
        torch.cuda.current_stream(self.device).synchronize()
        return (out, grad) if want_grad else out


_default_engine = None


def default_engine() -> GPEngine:
    """Process-wide engine on the current CUDA device (LOCAL_RANK under torchrun)."""
    global _default_engine
    if _default_engine is None:
        import os
        _default_engine = GPEngine(int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.device_count() > 1 else 0)
    return _default_engine
