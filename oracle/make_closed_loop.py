"""Generate tests/golden/closed_loop.npz: a trace of the UNMODIFIED reference workflow.

    python -m oracle.make_closed_loop          # needs /root/reference (read-only)

TEST INFRASTRUCTURE ONLY.  Runs, in this container and with
the reference's own code:
  * `select_parameters.py` unmodified (via runpy) with the reference `PointSelector`,
  * the DAGMan POST scripts `terminate_opto.py`, `terminate_block.py`, `terminate_algo.py`
    unmodified, chained exactly as `dag_templates/{main,algo,first_pair,second_pair,rise_time,opto}.dag`
    chain them (exit 1 = RETRY the node, exit 0 = next node),
  * a synthetic objective standing in for RAT + time_residuals.py (the write-back of
    time_residuals.py:166-182,204-217 is emulated line for line).
The scripts hard-code `/home/hunt-stokes/bayesian_optimisation`; `builtins.open` is wrapped so that
this prefix lands in a scratch directory, and the process chdir()s there (the scripts mix absolute
and cwd-relative paths, select_parameters.py:28,142,164,250,265).  Nothing outside the scratch
directory is written.

Every call the workflow makes into `PointSelector` is recorded: inputs (measured points/values,
candidate axes, length-scale grids) and outputs (kernel_params, selected index, extrema of the
acquisition).  The GPU tests replay those calls through the B200 drop-in and demand the same
kernel_params and the same selected index for the whole trace -- including the single-measurement
branch (point_selector.py:63-73), block restarts seeded from block_best_params (obj = 1e10 used as a
measured value, select_parameters.py:135,259) and the placeholder objective rows.
"""
from __future__ import annotations

import os

import numpy as np

from . import reference_loader as rl
from .workflow import NAMES, run_workflow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "closed_loop.npz")


def main(max_calls=400, sizes=(1, 1, 20)):
    """sizes: one full algorithm iteration with the deployment's 20-iteration sample loops (run_algo.py:9) for all five
    parameter groups: 100 PointSelector calls, M = 1 .. 21."""
    run = run_workflow(rl.load_reference_class(), sizes=sizes, max_calls=max_calls)
    rec_calls, objectives, events, final = run["calls"], run["objectives"], run["events"], run["final"]

    class rec:
        calls = rec_calls

    # pack: variable-length arrays are concatenated with offsets
    n = len(rec.calls)
    out = dict(n_calls=np.int64(n), final_parameters=np.array([final["parameters"][k] for k in NAMES]),
               objectives=np.array(objectives))
    out["M"] = np.array([len(c["X"]) for c in rec.calls])
    out["d"] = np.array([c["X"].shape[1] for c in rec.calls])
    out["X"] = np.concatenate([c["X"].reshape(-1) for c in rec.calls])
    out["y"] = np.concatenate([c["y"].reshape(-1) for c in rec.calls])
    # candidate axes and length-scale grids repeat: store the distinct ones
    tables, keys = [], {}
    def tid(a):
        k = a.tobytes()
        if k not in keys:
            keys[k] = len(tables)
            tables.append(a)
        return keys[k]
    axes_id, ls_id = [], []
    for c in rec.calls:
        P, fd = c["P"], c["fd"]
        axes = [np.unique(P[:, k]) for k in range(P.shape[1])]
        grid = np.stack([m.reshape(-1) for m in np.meshgrid(*axes, indexing="ij")], axis=1)
        assert np.array_equal(grid, P), "candidate grid is not the row-major product of its axes"
        axes_id.append([tid(a) for a in axes] + [-1] * (2 - len(axes)))
        ls_id.append([tid(a) for a in c["ls"]] + [-1] * (2 - len(c["ls"])))
    out["axes_id"], out["ls_id"] = np.array(axes_id), np.array(ls_id)
    out["table_len"] = np.array([len(t) for t in tables])
    out["tables"] = np.concatenate(tables)
    out["kp"] = np.array([np.pad(c["kp"].reshape(-1), (0, 2 - c["kp"].size)) for c in rec.calls])
    out["kp_ndim"] = np.array([c["kp"].ndim for c in rec.calls])
    out["index"] = np.array([np.pad(c["index"], (0, 2 - len(c["index"])), constant_values=-1) for c in rec.calls])
    for k in ("acq_max", "mu_min", "sig_max"):
        out[k] = np.array([c[k] for c in rec.calls])
    np.savez_compressed(OUT, **out)
    # every file the reference-class run left behind, for the byte-for-byte diff of tests/test_closed_loop_dropin.py
    names = sorted(run["files"])
    blob = b"".join(run["files"][k] for k in names)
    np.savez_compressed(OUT.replace("closed_loop.npz", "closed_loop_files.npz"), names=np.array(names),
                        sizes=np.array([len(run["files"][k]) for k in names]), blob=np.frombuffer(blob, dtype=np.uint8),
                        sizes_arg=np.array(sizes), seed=np.int64(12345))
    print(f"{n} PointSelector calls recorded; M range {out['M'].min()}..{out['M'].max()}; events {events}")
    print("final parameters", dict(zip(NAMES, out["final_parameters"])))


if __name__ == "__main__":
    main()
