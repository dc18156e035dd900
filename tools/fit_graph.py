"""Experiment: replay the fit as a CUDA graph (is the fit host-launch-bound?)."""
import sys, ctypes as C, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, JITTER_POSTERIOR
from bayesian_optimisation_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
dX, dy, dl = eng.to_device(X), eng.to_device(y), eng.to_device(ell)
nbytes = eng.lib.bogp_fit_workspace_bytes(n, d)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
def enqueue():
    eng._sync_stream()
    h = C.c_void_p()
    _lib.check(eng.lib.bogp_fit_enqueue(eng._ctx, dX.data_ptr(), dy.data_ptr(), n, d, dl.data_ptr(), JITTER_POSTERIOR, ws.data_ptr(), nbytes, C.byref(h)))
    return h
for _ in range(2):
    h = enqueue(); nl = C.c_double(); _lib.check(eng.lib.bogp_fit_status(h, C.byref(nl)))
print("eager nlml", nl.value)
ts = []
for _ in range(5):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); h = enqueue(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print(f"eager enqueue: {min(ts):.3f} ms")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        hg = enqueue()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
eng._sync_stream()
nl2 = C.c_double(); _lib.check(eng.lib.bogp_fit_status(hg, C.byref(nl2)))
print(f"graph replay: {min(ts):.3f} ms  nlml {nl2.value}")
