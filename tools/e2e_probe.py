"""Where does the end-to-end step (PointSelector with host buffers, as in bench.py's e2e leg) spend its time?
Wall-clock per phase with a device synchronise after each (so phases do not overlap here)."""
import sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid
from bayesian_optimisation_b200.point_selector import PointSelector

cands = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
eng = GPEngine(0)
X, y, ell = bench.synthetic()
f_best = float(y.min())
grid = CandidateGrid([np.linspace(0.0, 1.0, bench.GRID_PTS)] * bench.DIM)
pinned = torch.empty((cands, bench.DIM), dtype=torch.float64).pin_memory()
pinned.numpy()[:] = bench.grid_points_host(grid.axes, 0, cands)


def step():
    ps = PointSelector()
    ps._engine = eng
    ps.name, ps.iteration = "probe", 0
    ps.measured_pts, ps.measured_vals = X, y
    ps.feature_domain = [cands]
    ps.predicted_pts = pinned.numpy()
    ps.length_scales = np.array([0.3])
    t0 = time.perf_counter()
    ps.update_surrogate()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    idx = ps.expected_improvement(f_best)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, int(idx[0])


for _ in range(2):
    step()
for _ in range(3):
    a, b, i = step()
    print(f"update_surrogate {a:.2f} ms, expected_improvement {b:.2f} ms, total {a + b:.2f} ms -> {cands / (a + b) * 1e3:.3e} cand/s, idx {i}")
pr = cProfile.Profile()
pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
