"""BASELINE.json configs[2] in full: GP fit at N=4096, d=8 and an expected-improvement sweep over ALL 10^8 grid
candidates, sharded over the ranks of the job (contiguous flat-index ranges, replicated fit, one all_gather of the
24-byte winner records, device to device).  Arg-max-only sweeps are screened by the posterior-mean bound (default;
`--no-screen` scores every candidate exactly); the winner is the same either way.

    python tools/full_sweep.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/full_sweep.py

Prints one JSON line: wall and device time of fit + sweep + reduction, candidates/s, the selected flat index and its
grid coordinates.  The index must be the same for every GPU count (fixed accumulation orders, exact integer product)."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
from bayesian_optimisation_b200.sharding import allreduce_maxloc_device, shard_range

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--n", "--nobs", dest="n", type=int, default=bench.N_OBS, help="measured points (use --nobs under torchrun, whose own parser claims --n)")
ap.add_argument("--dim", type=int, default=bench.DIM)
ap.add_argument("--grid", type=int, default=bench.GRID_PTS, help="grid points per axis")
ap.add_argument("--kind", default="ei", choices=["ei", "lcb"])
ap.add_argument("--per-rank", type=int, default=0, help="score only this many candidates per rank (a slice of every shard); 0 = the whole grid")
ap.add_argument("--no-screen", action="store_true", help="score every candidate exactly (no posterior-mean screen)")
ap.add_argument("--begin", type=int, default=0, help="first flat index of the range to sweep (default: whole grid)")
ap.add_argument("--end", type=int, default=0, help="end of the range (0 = grid size)")
ap.add_argument("--chunk", type=int, default=0, help="candidates per kernel chunk (0 = engine default)")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eng = GPEngine(local)
eng.set_screening(not args.no_screen)
eng.set_global_seed(world > 1)          # every shard screens against the same floor
if (args.n, args.dim) == (bench.N_OBS, bench.DIM):
    X, y, ell = bench.synthetic()
else:
    from oracle import gp_oracle as _o           # synthetic inputs only (a tool, not the product path)
    X, y, ell = _o.synthetic_problem(args.n, args.dim)
grid = CandidateGrid([np.linspace(0.0, 1.0, args.grid)] * args.dim)
KIND = ACQ_EI if args.kind == "ei" else 0
dX, dy = eng.to_device(X), eng.to_device(y)
f_best = float(y.min())
r_begin, r_end = args.begin, (args.end or grid.size)
b, e = shard_range(r_end - r_begin, rank, world)
b, e = b + r_begin, e + r_begin
CH = dict(chunk=args.chunk) if args.chunk else {}
if args.per_rank:
    e = min(e, b + args.per_rank)
scored = (e - b) if not args.per_rank else args.per_rank * world
# warm-up (kernel attributes, workspaces) on a small slice
fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR); eng.acquire(fit, grid, b, min(e, b + 65536), kind=KIND, f_best=f_best, **CH); fit.close()
eng.screen_stats()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
res = eng.acquire(fit, grid, b, e, kind=KIND, f_best=f_best, sync=False, **CH)
score, index = allreduce_maxloc_device(eng, res.record)
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
n_scr, n_surv = eng.screen_stats()
if rank == 0:
    coords = np.unravel_index(index, grid.shape)
    total = (r_end - r_begin) if not args.per_rank else scored
    print(json.dumps({"config": f"N={args.n}, d={args.dim}, {args.kind.upper()} over " + (f"the full {args.grid}^{args.dim}-point grid" if not args.per_rank else f"{args.per_rank} candidates per rank of the {args.grid}^{args.dim}-point grid"),
                      "n_gpus": world, "candidates": total,
                      "device_seconds_max_over_ranks": float(ms.item()) * 1e-3, "wall_seconds_rank0": wall,
                      "candidates_per_s": total / (float(ms.item()) * 1e-3), "fit_ms_included": True, "best_score": score, "best_flat_index": int(index),
                      "best_grid_index": [int(c) for c in coords], "nlml": fit.nlml, "tensor_path": eng.acquire_path,
                      "range": [r_begin, r_end], "screened": not args.no_screen,
                      "survivor_fraction_rank0": (n_surv / n_scr) if n_scr else None}))
fit.close()
if world > 1:
    dist.destroy_process_group()
