#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: EI candidates scored per second at
N=4096 observations, d=8 (BASELINE.json `metric`, configs[2]), plus the GP fit time.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of synthetic input: a full GP fit
(Gram + Cholesky + log marginal likelihood + W = L^-1) followed by an expected-improvement sweep
with arg-max over a contiguous slice of `--cands` candidates per GPU of the 10^8-point grid
(10 points per axis, d=8).  Ranks score disjoint contiguous slices (weak scaling) and exchange
16 bytes per rank per step to pick the winner.  `value` = candidates scored by all ranks per
second, timed on the device with CUDA events, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy path (chunked, diag-only, inv-based)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample, verbose=False):
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


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    cores = blas_threads()
    v, ms, fit_ms = cpu_reference_run(args.steps, args.warmup, sample)
    desc = (f"oracle port of point_selector.py:78-98,166-195 + EI (numpy, inv-based, chunked diag-only), {sample} grid candidates "
            f"per step of the same N=4096,d=8 problem; the fit (Gram+inv+slogdet, {fit_ms:.0f} ms) is done once outside the timed steps")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.cands, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "fit_ms": fit_ms, "gpu_launches": 0}
    emit(line)


def workload_config(cands, gpus):
    return {"workload": f"synthetic GP N={N_OBS}, d={DIM}, ell=0.3; per step: GP fit (Gram+Cholesky+LML+L^-1) + EI sweep + arg-max over a "
                        f"contiguous {cands}-candidate slice per GPU of the {GRID_PTS}^{DIM}=1e8-point grid (BASELINE.json configs[2])",
            "n_obs": N_OBS, "dim": DIM, "candidates_per_step_per_gpu": cands, "grid_points_per_axis": GRID_PTS,
            "acquisition": "EI", "sharding": f"contiguous flat-index slices x{gpus}, Cholesky replicated, 16-byte all_gather max-loc",
            "l2": "per-step working set (k_* panel 537 MB + W 67 MB, re-streamed per kernel chunk) exceeds the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 6144^3, best of 5, CUDA events: the FP64 roofline denominator (MEASURED_PEAKS.json has no fp64 entry)."""
    n = 6144
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n ** 3 / best * 1e-9


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
    from bayesian_optimisation_b200.point_selector import PointSelector
    from bayesian_optimisation_b200.sharding import allreduce_maxloc, shard_range

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
    if args.path:
        eng.set_acquire_path(args.path)
    X, y, ell = synthetic()
    f_best = float(y.min())
    grid = CandidateGrid([np.linspace(0.0, 1.0, GRID_PTS)] * DIM)
    cands = args.cands
    dX, dy = eng.to_device(X), eng.to_device(y)

    def step_range(k):
        """global slice of step k: world*cands contiguous candidates, split across ranks"""
        g0 = (k * world * cands) % max(1, grid.size - world * cands)
        b, e = shard_range(world * cands, rank, world)
        return g0 + b, g0 + e

    def device_step(k):
        fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
        b, e = step_range(k)
        res = eng.acquire(fit, grid, b, e, kind=ACQ_EI, f_best=f_best, chunk=args.chunk)
        s, i = allreduce_maxloc(res.best_score, res.best_index, device=dev) if world > 1 else (res.best_score, res.best_index)
        fit.close()
        return s, i

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_fits(reps):
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f = eng.fit(dX, dy, ell, JITTER_POSTERIOR); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b)); f.close()
        return best

    # ---- fit time alone, before the sweeps (CUDA events, best of 20 back-to-back fits): the fit is latency-bound and
    #      follows the SM clock, which stays power-capped for a while after a sweep -- it is timed again after them
    fit_ms_before = time_fits(20)

    # ---- device-resident throughput ("value")
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
    ms = e0.elapsed_time(e1)
    launches = eng.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * cands * args.steps / (ms * 1e-3)

    # ---- the reference's own acquisition (explore*sigma - mu, "LCB/UCB", point_selector.py:204) on the same slices:
    #      same sweep, different epilogue; reported beside the EI headline
    def lcb_step(k):
        fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
        b, e = step_range(k)
        res = eng.acquire(fit, grid, b, e, kind=0, explore=4.0, chunk=args.chunk)
        out = allreduce_maxloc(res.best_score, res.best_index, device=dev) if world > 1 else (res.best_score, res.best_index)
        fit.close()
        return out
    lcb_step(0)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for k in range(2):
        lcb_step(1 + k)
    a1.record()
    barrier()
    tl = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    lcb_value = world * cands * 2 / (float(tl.item()) * 1e-3)

    fit_ms_after = time_fits(10)
    fit_ms = min(fit_ms_before, fit_ms_after)

    # ---- end to end through the reference-facing class with HOST buffers
    # PointSelector: host measured points + host candidate array in, host mean/sigma/acquisition out.
    e2e_cands = min(cands, args.e2e_cands)
    # two pinned host candidate blocks, filled BEFORE the timed region (generating synthetic input is not part of the
    # path); every timed step hands one of them to the drop-in, which copies it to the device itself
    pinned_blocks = [torch.empty((e2e_cands, DIM), dtype=torch.float64).pin_memory() for _ in range(2)]
    def e2e_step(k):
        b, _ = step_range(k)
        ps = PointSelector()
        ps._engine = eng                        # same context / tensor path as the device-resident leg
        ps.name, ps.iteration = "bench", k
        ps.measured_pts, ps.measured_vals = X, y
        ps.feature_domain = [e2e_cands]
        ps.predicted_pts = pinned_blocks[k % 2].numpy()
        ps.length_scales = np.array([0.3])      # one-point length-scale grid: one LML evaluation (tune_kernel) per step
        ps.update_surrogate()                   # LML fit + posterior fit + sweep; mu/sigma copied back to host arrays
        idx = ps.expected_improvement(f_best)   # EI + arg-max; acquisition copied back
        return int(idx[0]) + b
    for k in range(2):
        pinned_blocks[k].numpy()[:] = grid_points_host(grid.axes, step_range(k)[0], step_range(k)[0] + e2e_cands)
    e2e_steps = max(1, min(args.steps, 3))
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_cands * e2e_steps / float(t.item())
    h2d = (N_OBS * DIM + N_OBS) * 8 * 2 + e2e_cands * DIM * 8 + 2 * DIM * 8
    d2h = 3 * e2e_cands * 8 + 8 + 16 + 16

    # ---- roofline of the dominant kernel (the acquisition product V = L^-1 k_*), per-launch CUDA-event
    #      timing in a separate pass (the hooks serialise the stream, so never inside the timed steps)
    roof, cpu_base, fp64_peak = None, None, None
    if rank == 0:
        fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
        eng.profile(True)
        eng.acquire(fit, grid, 0, min(cands, 16 * args.chunk), kind=ACQ_EI, f_best=f_best, chunk=args.chunk)
        prof = {k: v for k, v in eng.profile_read().items() if k in ("panel", "trigemm", "finalize", "merge")}
        eng.profile(False)
        n_pad = fit.n_pad
        fit.close()
        fp64_peak = measure_fp64_peak(torch, dev)
        tri_ms, tri_n = prof["trigemm"]
        per_launch_ms = tri_ms / max(1, tri_n)
        chunk_c = min(args.chunk, cands)
        flops = float(N_OBS) ** 2 * chunk_c                      # N^2 fp64 flops per candidate (triangular product)
        fp64_equiv = flops / (per_launch_ms * 1e-3) * 1e-12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        tj = os.path.join(ROOT, "profiles", "trigemm_traffic.json")
        if os.path.isfile(tj):
            try:
                traffic = json.load(open(tj)).get(eng.acquire_path, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        total_prof = sum(v[0] for v in prof.values())
        share = {k: v[0] / total_prof for k, v in prof.items()} if total_prof > 0 else None
        if eng.acquire_path == "i8":
            # executed integer work: 34 digit pairs x n_pad*(n_pad+128)/2 MACs per candidate, 2 ops per MAC
            int8_ops = 34.0 * n_pad * (n_pad + 128) * chunk_c
            achieved = int8_ops / (per_launch_ms * 1e-3) * 1e-12
            bf16 = peaks.get("bf16_tflops")
            peak = 2.0 * bf16 if bf16 else 2.0 * 1590.0
            roof = {"bound": "tensor", "kernel": "trigemm_i8_kernel (tcgen05.mma kind::i8, TMEM accumulators; exact digit-slice fp64 product)",
                    "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak, "traffic": traffic,
                    "ms_per_launch": per_launch_ms, "launches_timed": tri_n, "executed_int8_ops_per_launch": int8_ops,
                    "algorithmic_flops_per_launch": flops, "fp64_equivalent_tflops": fp64_equiv,
                    "fp64_pipe_peak_tflops": max(fp64_peak, 37.0), "frac_of_fp64_pipe_peak": fp64_equiv / max(fp64_peak, 37.0),
                    "frac_of_nominal_int8_peak": achieved / 4500.0,
                    "ncu_utcimma_int8_pct_of_peak": (json.load(open(tj)).get("i8", {}).get("utcimma_int8_ops_pct_of_peak") if os.path.isfile(tj) else None),
                    "peak_source": ("dense int8 tensor peak taken as 2 x the measured cuBLAS bf16 burst figure of MEASURED_PEAKS.json "
                                    f"({bf16} TF/s; int8 runs at twice the bf16 rate, nominal 4500 vs 2250) -- of measured"
                                    if bf16 else "2 x the fallback bf16 figure 1590 TF/s -- of fallback"),
                    "share_of_sweep": share}
        else:
            dmma_peak = 37.0                                      # DMMA issue-rate peak measured with tools/dmma_bench (profiles/)
            peak = max(fp64_peak, dmma_peak)
            roof = {"bound": "tensor", "kernel": "trigemm_kernel (FP64 DMMA)", "achieved": fp64_equiv, "peak": peak, "unit": "TFLOP/s",
                    "frac": fp64_equiv / peak, "traffic": traffic, "ms_per_launch": per_launch_ms, "launches_timed": tri_n,
                    "algorithmic_flops_per_launch": flops,
                    "peak_source": f"fp64 is not in MEASURED_PEAKS.json: max(cuBLAS DGEMM 6144^3 measured live = {fp64_peak:.1f} TF/s, "
                                   f"DMMA.8x8x4 issue-rate microbenchmark tools/dmma_bench = {dmma_peak} TF/s)",
                    "share_of_sweep": share}
        if world == 1 and not args.no_cpu_baseline:
            cores = blas_threads()
            v, _, cfit = cpu_reference_run(2, 1, args.ref_sample)
            cpu_base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle port (numpy, inv-based, chunked diag-only) of the same N=4096,d=8 EI sweep on {args.ref_sample} grid candidates x 2 steps; "
                                  f"its fit (Gram+inv+slogdet) took {cfit:.0f} ms once, outside the timed sample"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": dict(workload_config(cands, world), tensor_path=eng.acquire_path), "fit_ms": fit_ms,
                "fit_ms_after_sweeps": fit_ms_after, "lcb_candidates_per_s": lcb_value,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "PointSelector.update_surrogate() + expected_improvement() with host numpy buffers", "candidates_per_step_per_gpu": e2e_cands,
                        "steps": e2e_steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
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
    ap.add_argument("--cands", type=int, default=1 << 20, help="candidates per step per GPU")
    ap.add_argument("--chunk", type=int, default=65536, help="candidates per kernel chunk (two half-chunks are pipelined)")
    ap.add_argument("--e2e-cands", type=int, default=1 << 20)
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
