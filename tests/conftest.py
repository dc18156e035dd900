import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree (/root/reference)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    has_gpu = _has_gpu()
    has_ref = os.path.isfile(os.environ.get("BOGP_REFERENCE_DIR", "/root/reference") + "/point_selector.py")
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(prefix))


def load_golden(name):
    import numpy as np
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def load_closed_loop():
    """Calls recorded from the unmodified reference workflow (oracle/make_closed_loop.py)."""
    import numpy as np
    g = load_golden("closed_loop")
    offs = np.concatenate([[0], np.cumsum(g["table_len"])])
    tables = [g["tables"][offs[i]:offs[i + 1]] for i in range(len(g["table_len"]))]
    calls, xo, yo = [], 0, 0
    for c in range(int(g["n_calls"])):
        M, d = int(g["M"][c]), int(g["d"][c])
        X = g["X"][xo:xo + M * d].reshape(M, d); xo += M * d
        y = g["y"][yo:yo + M]; yo += M
        axes = [tables[i] for i in g["axes_id"][c] if i >= 0]
        ls = [tables[i] for i in g["ls_id"][c] if i >= 0]
        mesh = np.meshgrid(*axes, indexing="ij")
        P = np.stack([m.reshape(-1) for m in mesh], axis=1)
        kp = g["kp"][c][:d]
        kp = kp.reshape(1, 1) if g["kp_ndim"][c] == 2 else kp
        calls.append(dict(X=X, y=y, P=P, axes=axes, feature_domain=[len(a) for a in axes],
                          length_scales=np.array(ls) if len(ls) == 2 else ls[0], kernel_params=kp,
                          index=g["index"][c][:d], acq_max=float(g["acq_max"][c]), mu_min=float(g["mu_min"][c]),
                          sig_max=float(g["sig_max"][c])))
    return calls


# ---------------------------------------------------------------------------------------------
# achieved parity errors: every GPU parity test reports the largest error it saw, the session prints the
# table and writes it to gpurun_out/parity_errors.json (copied to profiles/ by hand once per round)
# ---------------------------------------------------------------------------------------------
_ERRORS = []


def record_error(test, quantity, value, bound=None, note=""):
    _ERRORS.append({"test": test, "quantity": quantity, "value": float(value), "bound": None if bound is None else float(bound), "note": note})
    print(f"[parity] {test}: {quantity} = {float(value):.3e}" + (f" (bound {float(bound):.3e})" if bound is not None else "") + (f" {note}" if note else ""))


def pytest_sessionfinish(session, exitstatus):
    if not _ERRORS:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_errors.json"), "w") as f:
            json.dump(_ERRORS, f, indent=1)
    except OSError:
        pass
