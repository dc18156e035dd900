"""Host-side cost of GPEngine.fit(): how long the GPU waits for Python/ctypes before the first kernel.
perf_counter stamps around each step of fit() (no device synchronisation in between)."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200 import _lib
from bayesian_optimisation_b200.engine import GPEngine, JITTER_POSTERIOR
n, d = 4096, 8
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
dX, dy = eng.to_device(X), eng.to_device(y)
for _ in range(3):
    eng.fit(dX, dy, ell, JITTER_POSTERIOR).close()
torch.cuda.synchronize()
T = time.perf_counter
for rep in range(3):
    t = [T()]
    eng._sync_stream(); t.append(T())
    dx, dyy = eng.to_device(dX), eng.to_device(dy.reshape(-1)); t.append(T())
    ell_np = np.asarray(ell, dtype=np.float64).reshape(-1); dl = eng.to_device(ell_np); t.append(T())
    nbytes = eng.lib.bogp_fit_workspace_bytes(n, d); ws = torch.empty(nbytes, dtype=torch.uint8, device=eng.device); t.append(T())
    h = C.c_void_p()
    code = eng.lib.bogp_fit_enqueue(eng._ctx, dx.data_ptr(), dyy.data_ptr(), n, d, dl.data_ptr(), float(JITTER_POSTERIOR), ws.data_ptr(), nbytes, C.byref(h)); t.append(T())
    nl = C.c_double(); eng.lib.bogp_fit_status(h, C.byref(nl)); t.append(T())
    eng.lib.bogp_fit_destroy(h)
    names = ["sync_stream", "to_device x,y", "ell H2D", "workspace", "fit_enqueue (host)", "fit_status (wait)"]
    print("  ".join(f"{nm} {1e6 * (b - a):.0f}us" for nm, a, b in zip(names, t[:-1], t[1:])), f"| total {1e3 * (t[-1] - t[0]):.3f} ms")
