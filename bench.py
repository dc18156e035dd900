#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: EI candidates scored per second at
N=4096 observations, d=8 (BASELINE.json `metric`, configs[2]), plus the GP fit time.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W [--scaling strong]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic input: a full GP fit
(Gram + Cholesky + log marginal likelihood + W = L^-1) followed by an expected-improvement sweep
with arg-max over a contiguous slice of the 10^8-point grid (10 points per axis, d=8).
  --scaling weak   (default): `--cands` candidates per GPU per step, the slice grows with N;
  --scaling strong: `--total-cands` candidates per step in total, split over the N ranks
                    (north_star: "1e8 candidates, EI sharded across 1/2/4/8").
Ranks score disjoint contiguous slices and exchange ONE 24-byte record per rank per step, device to
device, to pick the winner.  `value` = candidates scored by all ranks per second, timed on the device
with CUDA events, max over ranks.  `e2e` = the same through `PointSelector.update_surrogate()` +
`expected_improvement()` with pageable host numpy arrays (copies inside the timed region).
"""
from __future__ import annotations

import os
import sys

# The CPU arm must use every host core also under torchrun (which exports OMP_NUM_THREADS=1): BLAS reads these
# variables when numpy is first imported, so they are set before that import.
if "reference" in sys.argv[1:] or "--impl=reference" in sys.argv[1:]:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIM, GRID_PTS = 4096, 8, 10
METRIC = "EI candidates scored/sec at N=4096,d=8; GP fit (Cholesky+LML) ms"   # BASELINE.json `metric`; `value` is the first part, `fit_ms` the second
UNIT = "candidates/s"


def synthetic(seed=0):
    """X ~ U[0,1]^{N x d}, y = sin(3 sum x) + 0.1 N(0,1), ell = 0.3."""
    rng = np.random.default_rng(seed)
    X = rng.random((N_OBS, DIM))
    y = np.sin(3.0 * X.sum(axis=1)) + 0.1 * rng.standard_normal(N_OBS)
    return X, y, np.full(DIM, 0.3)


def grid_points_host(axes, start, stop):
    """Rows [start, stop) of the row-major Cartesian grid (axis 0 slowest, select_parameters.py:273-279):
    the synthetic HOST candidate array fed to the end-to-end leg."""
    flat = np.arange(start, stop, dtype=np.int64)
    out = np.empty((len(flat), len(axes)))
    for k in range(len(axes) - 1, -1, -1):
        n = len(axes[k])
        out[:, k] = np.asarray(axes[k], dtype=np.float64)[flat % n]
        flat = flat // n
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def workload_config(scaling, gpus):
    """The workload both arms run -- identical in the two JSON lines; what each arm scored per step is stated beside
    it (`candidates_per_step_per_gpu`, `cpu_baseline.sample`), not inside it."""
    return {"workload": f"synthetic GP N={N_OBS}, d={DIM}, ell=0.3; per step: GP fit (Gram+Cholesky+LML+L^-1) + EI sweep + arg-max over a "
                        f"contiguous candidate slice per GPU of the {GRID_PTS}^{DIM}=1e8-point grid (BASELINE.json configs[2])",
            "n_obs": N_OBS, "dim": DIM, "grid_points_per_axis": GRID_PTS, "acquisition": "EI", "scaling": scaling,
            "sharding": f"contiguous flat-index slices x{gpus}, Cholesky replicated, one 24-byte (score, index, nan) record per rank, "
                        "all_gather + fold on the device (max score, then min flat index)",
            "l2": "per-step working set (k_* panel digits 2 x 0.9 GB + W digits 60 MB, re-streamed per kernel chunk) exceeds the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy path (chunked, diag-only, inv-based)
# ---------------------------------------------------------------------------------------------
def blas_threads_all_cores():
    """Pin the BLAS pool to every host core (also under torchrun) and return the thread count actually in use."""
    want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return want


def cpu_reference_run(steps, warmup, sample):
    """Restatement of point_selector.py:78-98,166-195 + EI on `sample` grid candidates per step (oracle/gp_oracle.py)."""
    from oracle import gp_oracle as o
    X, y, ell = synthetic()
    axes = [np.linspace(0.0, 1.0, GRID_PTS)] * DIM
    t0 = time.perf_counter()
    K = o.kernel_rbf_chunked(X, X, ell)
    K[np.diag_indices_from(K)] += o.JITTER_KERNEL + o.JITTER_EXTRA
    inv = np.linalg.inv(K)                           # point_selector.py:89
    sign, logdet = np.linalg.slogdet(K)              # stands in for np.log(np.linalg.det(.)) (:118), which underflows here
    alpha = inv @ y
    fit_s = time.perf_counter() - t0
    f_best = float(y.min())
    ell2 = ell ** 2

    def sweep(c0):
        P = o.grid_points(axes, c0, c0 + sample)
        best = (-np.inf, -1)
        for s in range(0, sample, 1024):
            Pc = P[s:s + 1024]
            Ks = np.exp(-0.5 * np.sum((Pc[:, None, :] - X[None, :, :]) ** 2 / ell2, axis=2))
            mu = Ks @ alpha
            var = o.PRIOR_DIAG - np.einsum("cm,cm->c", Ks @ inv, Ks)
            ei = o.expected_improvement(mu, np.sqrt(np.abs(var)), f_best)
            i = int(np.flatnonzero(ei == ei.max())[0])
            if ei[i] > best[0]:
                best = (float(ei[i]), c0 + s + i)
        return best

    for w in range(warmup):
        sweep(w * sample)
    t0 = time.perf_counter()
    for k in range(steps):
        sweep((warmup + k) * sample)
    dt = time.perf_counter() - t0
    return sample * steps / dt, dt / steps * 1e3, fit_s * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    cores = blas_threads_all_cores()
    v, ms, fit_ms = cpu_reference_run(args.steps, args.warmup, sample)
    desc = (f"oracle port of point_selector.py:78-98,166-195 + EI (numpy, inv-based, chunked diag-only) on {cores} BLAS threads "
            f"(os.cpu_count() = {os.cpu_count()}): a bounded sample of {sample} grid candidates per step of the same N=4096,d=8 workload "
            f"(the GPU arm scores {args.cands} per GPU per step); the fit (Gram+inv+slogdet, {fit_ms:.0f} ms) is done once, outside the "
            "timed steps, which favours the CPU arm")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.scaling, args.gpus), "candidates_per_step_per_gpu": sample,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "fit_ms": fit_ms, "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
    from bayesian_optimisation_b200.point_selector import PointSelector
    from bayesian_optimisation_b200 import session as sm
    from bayesian_optimisation_b200.sharding import allreduce_maxloc_device, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if "BOGP_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["BOGP_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    eng = GPEngine(local)
    session = sm.Session([local])
    sm.set_default_session(session)
    if args.path:
        eng.set_acquire_path(args.path)
        session.set_acquire_path(args.path)
    X, y, ell = synthetic()
    f_best = float(y.min())
    grid = CandidateGrid([np.linspace(0.0, 1.0, GRID_PTS)] * DIM)
    strong = args.scaling == "strong"
    step_total = args.total_cands if strong else world * args.cands          # candidates per step, all ranks
    cands = -(-step_total // world)                                           # per rank (ceil)
    dX, dy = eng.to_device(X), eng.to_device(y)

    def step_range(k):
        """this rank's slice of step k: `step_total` contiguous candidates, split across the ranks"""
        g0 = (k * step_total) % max(1, grid.size - step_total)
        b, e = shard_range(step_total, rank, world)
        return g0 + b, g0 + e

    def device_step(k, kind=ACQ_EI, **kw):
        fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
        b, e = step_range(k)
        res = eng.acquire(fit, grid, b, e, kind=kind, chunk=args.chunk, sync=False, **kw)
        s, i = allreduce_maxloc_device(eng, res.record)          # world == 1: just the 24-byte read
        fit.close()
        return s, i

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_fits(reps):
        out = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f = eng.fit(dX, dy, ell, JITTER_POSTERIOR); b.record(); torch.cuda.synchronize()
            out.append(a.elapsed_time(b)); f.close()
        return out

    # ---- measured peaks of the two pipes, cold (burst): the roofline denominators (csrc/peaks.cu)
    i8_burst, _ = eng.measure_peak("i8")
    f64_burst, _ = eng.measure_peak("fp64")

    # ---- fit time alone, before the sweeps (CUDA events, 20 back-to-back fits): the fit is latency-bound and follows the
    #      SM clock, which stays power-capped for a while after a sweep -- it is timed again after them, and THAT is the headline
    fits_before = time_fits(20)

    # ---- device-resident throughput ("value"): every candidate of the slice gets its exact posterior variance and score
    #      (screening of arg-max-only sweeps, which drops candidates that provably cannot win, is switched OFF here and
    #      measured separately below)
    eng.set_screening(False)
    for k in range(args.warmup):
        device_step(k)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    last = None
    for k in range(args.steps):
        last = device_step(args.warmup + k)
    e1.record()
    barrier()
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = step_total * args.steps / (ms * 1e-3)
    fits_after = time_fits(15)

    # ---- the reference's own acquisition (explore*sigma - mu, "LCB/UCB", point_selector.py:204) on the same slices:
    #      same sweep, different epilogue; reported beside the EI headline
    device_step(0, kind=0, explore=4.0)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for k in range(2):
        device_step(1 + k, kind=0, explore=4.0)
    a1.record()
    barrier()
    lcb_value = step_total * 2 / (max_over_ranks(a0.elapsed_time(a1)) * 1e-3)

    # ---- the same steps with each sweep as ONE persistent fused kernel (csrc/acquire_fused.cu): bit-identical results, no k_*
    #      traffic to HBM; reported beside the headline, which uses the (faster) two-stream pipeline of separate kernels
    eng.set_fused(True)
    fz_last = device_step(0)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for k in range(2):
        fz_last = device_step(args.warmup + args.steps - 2 + k)
    f1.record()
    barrier()
    fused_value = step_total * 2 / (max_over_ranks(f0.elapsed_time(f1)) * 1e-3)
    eng.set_fused(False)
    fused = {"value": fused_value, "unit": UNIT, "same_winner_as_separate_kernels": bool(fz_last == last),
             "what": "every sweep as one launch of the persistent fused kernel (grid index -> k_* digits -> tcgen05 product -> sigma^2, mu -> EI -> "
                     "max-loc; k_* only in an L2-resident ring); include/bogp.h bogp_set_fused"}

    # ---- arg-max-only sweeps with the screen the library applies by default: the exact (score, index) of the full sweep, the
    #      posterior means of all grid candidates from fp64 GEMMs over per-axis factor tables (csrc/screen_gemm.cu), the N^2
    #      product only for candidates whose bound reaches the running best.  (a) the headline's steps, to check the winner;
    #      (b) steps of 2^26 candidates per GPU (a screened sweep is so short that at 2^20 candidates the fit dominates).
    eng.set_screening(True)
    eng.screen_stats()
    scr_last = device_step(args.warmup + args.steps - 1)
    scr_per_rank = min(1 << 26, grid.size // world)
    scr_total = scr_per_rank * world

    def screened_step(k, kind=ACQ_EI):
        fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
        b0 = (k * scr_total) % max(1, grid.size - scr_total + 1) + rank * scr_per_rank
        res = eng.acquire(fit, grid, b0, b0 + scr_per_rank, kind=kind, explore=4.0, f_best=f_best, chunk=args.chunk, sync=False)
        out = allreduce_maxloc_device(eng, res.record)
        fit.close()
        return out
    screened_step(0)
    eng.screen_stats()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for k in range(args.steps):
        screened_step(k)
    s1.record()
    barrier()
    scr_ms = max_over_ranks(s0.elapsed_time(s1))
    screened_value = scr_total * args.steps / (scr_ms * 1e-3)
    n_scr, n_surv = eng.screen_stats()
    screened = {"value": screened_value, "unit": UNIT, "candidates_per_step": scr_total, "ms_per_step": scr_ms / args.steps,
                "mean_gemm_tflops_lower_bound": 2.0 * N_OBS * scr_per_rank / (scr_ms / args.steps * 1e-3) * 1e-12,      # per GPU; the step also holds the fit and the exact passes
                "survivor_fraction_rank0": (n_surv / n_scr) if n_scr else None,
                "same_winner_as_full_sweep": bool(scr_last == last),
                "what": "arg-max-only EI sweep with the posterior-mean screen (include/bogp.h bogp_set_screening), fit included in every step: exact "
                        "(score, index); the means of all grid candidates come from fp64 GEMMs over per-axis kernel-factor tables "
                        "(csrc/screen_gemm.cu), the N^2 product runs only for candidates whose bound A(mu - eps, sqrt(prior)) reaches the running best"}
    # the reference's own acquisition (explore * sigma - mu, explore = 4): its screen also needs the nearest-measurement variance
    # bound (a max-times product on the CUDA cores, csrc/screen_gemm.cu gs_kmax_kernel), so a step is about twice as long
    screened_step(0, kind=0)
    barrier()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    for k in range(args.steps):
        screened_step(k, kind=0)
    q1.record()
    barrier()
    screened["lcb_value"] = scr_total * args.steps / (max_over_ranks(q0.elapsed_time(q1)) * 1e-3)
    eng.set_screening(False)

    # ---- end to end through the reference-facing class: pageable host numpy in, host numpy out.  Under torchrun every
    #      rank hands the WHOLE candidate array of the step to PointSelector, which scores its slice on its GPU, all-gathers
    #      mu / sigma (every rank ends up with the full arrays, the API of the single-process class) and picks the winner
    #      with the one-record exchange.
    e2e_total = min(step_total, args.e2e_cands * world)
    blocks = [grid_points_host(grid.axes, (k * e2e_total) % (grid.size - e2e_total), (k * e2e_total) % (grid.size - e2e_total) + e2e_total)
              for k in range(2)]                                   # generated BEFORE the timed region; plain (pageable) numpy arrays
    e2e_launch0 = [0]

    def e2e_step(k):
        ps = PointSelector()
        ps.name, ps.iteration = "bench", k
        ps.measured_pts, ps.measured_vals = X, y
        ps.feature_domain = [e2e_total]
        ps.predicted_pts = blocks[k % 2]
        ps.length_scales = np.array([0.3])      # one-point length-scale grid: one LML evaluation (tune_kernel) per step
        ps.update_surrogate()                   # LML fit + posterior fit + sweep; mu / sigma come back as host arrays
        idx = ps.expected_improvement(f_best)   # EI + arg-max on the device copy; acquisition comes back as a host array
        return int(idx[0])
    e2e_step(0); e2e_step(1)
    barrier()
    e2e_launch0[0] = session.launches
    t0 = time.perf_counter()
    for k in range(args.steps):
        e2e_step(k)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_launches = session.launches - e2e_launch0[0]
    e2e_value = e2e_total * args.steps / e2e_s
    b_, e_ = shard_range(e2e_total, rank, world)
    h2d = (N_OBS * DIM + N_OBS) * 8 * 2 + (e_ - b_) * DIM * 8 + 2 * DIM * 8                 # X, y twice (LML fit, posterior fit), this rank's slice, ell
    d2h = 3 * (e_ - b_) * 8 + 8 + 2 * 24                                                    # mu, sigma, EI of the slice, nlml, two winner records
    if world > 1:       # the slices go back up for the NCCL all_gather and the full arrays come down on every rank
        h2d += 3 * (e_ - b_) * 8
        d2h += 3 * e2e_total * 8

    # ---- roofline of the dominant kernel (the acquisition product V = L^-1 k_*), per-launch CUDA-event timing in a separate
    #      pass (the hooks serialise the stream, so never inside the timed steps), and of the fit
    roof = roof_fit = cpu_base = None
    if rank == 0:
        fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
        eng.profile(True)
        prof_cands = min(cands, 16 * args.chunk)
        eng.acquire(fit, grid, 0, prof_cands, kind=ACQ_EI, f_best=f_best, chunk=args.chunk)
        prof = {k: v for k, v in eng.profile_read().items() if k in ("panel", "trigemm", "finalize", "merge")}
        eng.profile(False)
        n_pad = fit.n_pad
        fit.close()
        i8_hot, i8_sus = eng.measure_peak("i8", 0.5)               # right after the serialised sweep: same thermal state
        f64_hot, f64_sus = eng.measure_peak("fp64", 0.3)
        tri_ms, tri_n = prof["trigemm"]
        per_launch_ms = tri_ms / max(1, tri_n)
        launch_c = prof_cands // max(1, tri_n)                       # candidates per tri-GEMM launch (profiling runs unpipelined chunks)
        flops = float(N_OBS) ** 2 * launch_c                         # N^2 fp64 flops per candidate (triangular product)
        fp64_equiv = flops / (per_launch_ms * 1e-3) * 1e-12
        traffic, tnote = None, None
        tj = os.path.join(ROOT, "profiles", "trigemm_traffic.json")
        if os.path.isfile(tj):
            try:
                rec = json.load(open(tj)).get(eng.acquire_path, {})
                if int(rec.get("candidates_per_launch", -1)) == int(launch_c):
                    traffic, tnote = rec.get("dram_bytes_per_launch"), rec.get("source")
                else:
                    tnote = f"no ncu capture of a {launch_c}-candidate launch under profiles/ (the stored one is {rec.get('candidates_per_launch')})"
            except Exception:
                pass
        total_prof = sum(v[0] for v in prof.values())
        share = {k: v[0] / total_prof for k, v in prof.items()} if total_prof > 0 else None
        peaks = {"int8_umma_burst_tops": i8_burst, "int8_umma_after_sweep_tops": i8_hot, "int8_umma_sustained_0.5s_tops": i8_sus,
                 "fp64_dmma_burst_tflops": f64_burst, "fp64_dmma_after_sweep_tflops": f64_hot, "fp64_dmma_sustained_0.3s_tflops": f64_sus,
                 "how": "bogp_measure_peak (csrc/peaks.cu): issue-rate loops on this GPU in this run; burst = best of 5 launches before any sweep"}
        if eng.acquire_path == "i8":
            # executed integer work: 34 digit pairs x n_pad*(n_pad+128)/2 MACs per candidate, 2 ops per MAC
            int8_ops = 34.0 * n_pad * (n_pad + 128) * launch_c
            achieved = int8_ops / (per_launch_ms * 1e-3) * 1e-12
            roof = {"bound": "tensor", "kernel": "trigemm_i8_kernel (tcgen05.mma kind::i8, TMEM accumulators; exact digit-slice fp64 product)",
                    "achieved": achieved, "peak": i8_burst, "unit": "TOP/s", "frac": achieved / i8_burst, "traffic": traffic, "traffic_source": tnote,
                    "frac_of_sustained_peak": achieved / i8_sus, "ms_per_launch": per_launch_ms, "launches_timed": tri_n,
                    "candidates_per_launch": launch_c, "executed_int8_ops_per_launch": int8_ops, "algorithmic_flops_per_launch": flops,
                    "fp64_equivalent_tflops": fp64_equiv, "frac_of_fp64_pipe_peak": fp64_equiv / f64_burst,
                    "peak_source": "of measured: int8 tcgen05.mma issue-rate peak of THIS GPU (burst, max clocks), measured in this run; "
                                   "MEASURED_PEAKS.json has no int8 entry", "peaks": peaks, "share_of_sweep": share,
                    "frac_clock_scaled": (achieved / i8_burst) / (clocks["sm_mhz"] / clocks["sm_max_mhz"]) if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") else None,
                    "frac_clock_scaled_note": "frac divided by (median SM clock under load / max SM clock): the sweep runs at the 1 kW power cap, "
                                              "the burst peak was measured at max clocks; ncu (profiles/r02_trigemm_i8_65536_ncu_raw.csv) reads 89 % of the "
                                              "utcimma int8 peak at its own clock"}
        else:
            roof = {"bound": "tensor", "kernel": "trigemm_kernel (FP64 DMMA)", "achieved": fp64_equiv, "peak": f64_burst, "unit": "TFLOP/s",
                    "frac": fp64_equiv / f64_burst, "traffic": traffic, "traffic_source": tnote, "ms_per_launch": per_launch_ms,
                    "launches_timed": tri_n, "candidates_per_launch": launch_c, "algorithmic_flops_per_launch": flops,
                    "peak_source": "of measured: DMMA.8x8x4 issue-rate peak of THIS GPU, measured in this run", "peaks": peaks, "share_of_sweep": share}
        fit_ms_head = float(np.median(fits_after))
        fit_flops = float(N_OBS) ** 3 / 3.0
        roof_fit = {"bound": "tensor", "kernel": "GP fit: gram + blocked Cholesky (DMMA SYRK) + L^-1 + alpha + nlml", "unit": "TFLOP/s",
                    "achieved": fit_flops / (fit_ms_head * 1e-3) * 1e-12, "peak": f64_burst, "frac": fit_flops / (fit_ms_head * 1e-3) * 1e-12 / f64_burst,
                    "algorithmic_flops": fit_flops, "note": "SURVEY 8d: N^3/3 flops per fit (the N^3/3 of the triangular inverse is extra work the "
                    "design adds and is not counted); latency-bound by the serial chain of diagonal blocks at this N", "ms": fit_ms_head}
        if world == 1 and not args.no_cpu_baseline:
            cores = blas_threads_all_cores()
            v, _, cfit = cpu_reference_run(2, 1, args.ref_sample)
            cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle port (numpy, inv-based, chunked diag-only) of the same N=4096,d=8 EI sweep on {args.ref_sample} grid candidates x 2 steps, "
                                  f"{cores} BLAS threads; its fit (Gram+inv+slogdet) took {cfit:.0f} ms once, outside the timed sample"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": dict(workload_config(args.scaling, world), tensor_path=eng.acquire_path),
                "candidates_per_step_per_gpu": cands, "candidates_per_step": step_total,
                "fit_ms": float(np.median(fits_after)), "fit_ms_best": float(min(fits_before + fits_after)),
                "fit_ms_before_sweeps_median": float(np.median(fits_before)), "fit_ms_after_sweeps_median": float(np.median(fits_after)),
                "lcb_candidates_per_s": lcb_value, "fused_single_kernel": fused, "argmax_only_screened": screened,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "PointSelector.update_surrogate() + expected_improvement(): pageable host numpy arrays in and out through the "
                               "host-buffer C ABI (bogp_session_*)" + ("; one process per GPU, slices all-gathered, one-record max-loc exchange" if world > 1 else ""),
                        "candidates_per_step": e2e_total, "steps": args.steps, "gpu_launches": e2e_launches},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_fit": roof_fit, "cpu_baseline": cpu_base,
                "selected": {"score": last[0], "flat_index": last[1]}}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def capture_stdout():
    """Everything that libraries print to stdout (e.g. NCCL's version banner) goes to stderr; the one
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cands", type=int, default=1 << 20, help="weak scaling: candidates per step per GPU")
    ap.add_argument("--total-cands", type=int, default=1 << 23, help="strong scaling: candidates per step, all GPUs together")
    ap.add_argument("--chunk", type=int, default=65536, help="candidates per kernel chunk (two half-chunks are pipelined)")
    ap.add_argument("--e2e-cands", type=int, default=1 << 20, help="end-to-end leg: candidates per step per GPU (at most)")
    ap.add_argument("--ref-sample", type=int, default=8192, help="candidates per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", default=None, choices=["i8", "fp64"], help="tensor path of the acquisition product (default: library default, i8)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3          # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
