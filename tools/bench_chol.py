"""Time bogp_cholesky alone (CUDA events, preallocated buffers): python tools/bench_chol.py n [reps]"""
import sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine
from bayesian_optimisation_b200 import _lib
eng = GPEngine(0)
for n in [int(a) for a in sys.argv[1].split(",")]:
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    X, y, ell = o.synthetic_problem(n, 8, seed=1)
    K = torch.from_numpy(o.kernel_rbf(X, X, ell)).cuda()
    a = K.clone(); linv = torch.zeros_like(a)
    X8 = X
    scal = torch.zeros(1, dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    def run():
        _lib.check(eng.lib.bogp_cholesky(eng._ctx, a.data_ptr(), n, n, linv.data_ptr(), scal.data_ptr(), info.data_ptr()))
    ts = []
    for r in range(reps + 3):
        a.copy_(K); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        if r >= 3: ts.append(e0.elapsed_time(e1))
    print(f"cholesky n={n}: min {min(ts)*1e3:.1f} us  median {np.median(ts)*1e3:.1f} us  info={int(info.item())}  ({2*n**3/3/np.median(ts)*1e-9:.2f} TFLOP/s... n^3/3 flops)")

    eng.profile(True)
    a.copy_(K); run(); torch.cuda.synchronize()
    print({k: (round(v[0]*1e3/max(1,v[1]),1), v[1]) for k, v in eng.profile_read().items() if v[1]}, "(avg us, launches)")
    eng.profile(False)

    # full fit with per-kernel timing
    eng.profile(True)
    f = eng.fit(X, y, ell); torch.cuda.synchronize()
    print("fit:", {k: (round(v[0]*1e3/max(1,v[1]),1), v[1], round(v[0],2)) for k, v in eng.profile_read().items() if v[1]}, "(avg us, launches, total ms)")
    eng.profile(False); f.close()
