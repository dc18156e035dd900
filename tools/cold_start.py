"""Cold-start cost of ONE select-parameters step (the reference runs one OS process per DAG node; VERDICT r1
missing #5).  Spawns fresh interpreters and measures wall time from process start to exit for

  dropin     `from point_selector import PointSelector` through dropin/ (numpy + ctypes + libbogp.so; no torch),
             update_surrogate() + lower_confidence_bound() at the native size (M = 21, C = 2500, 50 x 50 length scales)
  torch      the same after `import torch` (what round 1's host layer paid)
  oracle     the numpy restatement of the reference's arithmetic (oracle.select_next) in a fresh process: the CPU cost
             of the same step on THIS host (the literal class needs /root/reference and is timed in the build container,
             profiles/r02_cold_start.json "reference_literal_build_container")

and prints one JSON object; the child prints its own phase breakdown.   python tools/cold_start.py [reps]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, time, json
t0 = time.perf_counter()
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
MODE = {mode!r}
if MODE == "torch":
    import torch
import numpy as np
t_np = time.perf_counter()
g = dict(np.load({root!r} + "/tests/golden/native2d_t1t2_m21.npz"))
if MODE == "oracle":
    from oracle import gp_oracle as o
    r = o.select_next(g["X"], g["y"], g["P"], list(g["feature_domain"]), np.array([g["ls0"], g["ls1"]]))
    idx = r["index"]; t_imp = t_ctx = t_np; t_upd = time.perf_counter()
else:
    sys.path.insert(0, {root!r} + "/dropin")
    from point_selector import PointSelector
    t_imp = time.perf_counter()
    ps = PointSelector()
    ps.name, ps.iteration = "cold", 1
    ps.measured_pts, ps.measured_vals = g["X"], g["y"]
    ps.feature_domain, ps.predicted_pts = list(g["feature_domain"]), g["P"]
    ps.length_scales = np.array([g["ls0"], g["ls1"]])
    ps._sess()                      # CUDA context + library load
    t_ctx = time.perf_counter()
    ps.update_surrogate()
    t_upd = time.perf_counter()
    idx = ps.lower_confidence_bound()
t1 = time.perf_counter()
assert list(idx) == list(g["index"]), (idx, g["index"])
print(json.dumps(dict(mode=MODE, imports_s=t_np - t0, dropin_import_s=t_imp - t_np, context_s=t_ctx - t_imp, update_surrogate_s=t_upd - t_ctx,
                      acquisition_s=t1 - t_upd, inside_s=t1 - t0, torch_loaded="torch" in sys.modules)))
"""


def run(mode, reps):
    walls, last = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        p = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, mode=mode)], capture_output=True, text=True)
        walls.append(time.perf_counter() - t0)
        if p.returncode != 0:
            return {"error": p.stderr[-800:]}
        last = json.loads(p.stdout.strip().splitlines()[-1])
    return {"wall_s": walls, "wall_s_best": min(walls), "phases_last": last}


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dropin", "torch", "oracle"]
    out = {m: run(m, reps) for m in modes}
    out["host_cores"] = os.cpu_count()
    print(json.dumps(out))
