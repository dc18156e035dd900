"""TEST INFRASTRUCTURE ONLY: run the UNMODIFIED reference workflow in-process.

Runs, with the reference's own code from /root/reference (never copied into this repository):
  * `select_parameters.py` unmodified (via runpy) against whatever module is installed as `point_selector`
    (the reference class, or the B200 drop-in `dropin/point_selector.py`),
  * the DAGMan POST scripts `terminate_opto.py`, `terminate_block.py`, `terminate_algo.py`
    unmodified, chained exactly as `dag_templates/{main,algo,first_pair,second_pair,rise_time,opto}.dag`
    chain them (exit 1 = RETRY the node, exit 0 = next node),
  * a synthetic objective standing in for RAT + time_residuals.py (the write-back of
    time_residuals.py:166-182,204-217 is emulated line for line).
The scripts hard-code `/home/hunt-stokes/bayesian_optimisation`; `builtins.open` is wrapped so that
this prefix lands in a scratch directory, and the process chdir()s there (the scripts mix absolute
and cwd-relative paths, select_parameters.py:28,142,164,250,265).  Nothing outside the scratch
directory is written.

`run_workflow` returns every file the workflow left behind (`.npy` measured points, `opto_log.JSON`,
`macros/*.mac`, `submit_files/simulate.submit`), so two runs -- reference class vs drop-in -- can be diffed
byte for byte, plus the recorded `PointSelector` calls (used for tests/golden/closed_loop.npz).
"""
from __future__ import annotations

import builtins
import contextlib
import io
import json
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np

from . import reference_loader as rl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARD = "/home/hunt-stokes/bayesian_optimisation"
NAMES = ["T1", "T2", "T3", "T4", "TR", "A1", "A2", "A3", "A4"]
TRUE = dict(T1=4.9, T2=22.0, T3=110.0, T4=380.0, TR=0.85, A1=0.62, A2=0.28, A3=0.07, A4=0.03)

def emission_hist(p):
    """P(t) = sum_i A_i (exp(-t/t_i) - exp(-t/t_r)) / (t_i - t_r), binned on np.arange(-5, 250, 1)
    (docs/README.md:20; binning of time_residuals.py:130-132)."""
    edges = np.arange(-5, 250, 1.0)
    t = 0.5 * (edges[:-1] + edges[1:])
    pdf = np.zeros_like(t)
    for a, ti in zip([p["A1"], p["A2"], p["A3"], p["A4"]], [p["T1"], p["T2"], p["T3"], p["T4"]]):
        pdf += a * (np.exp(-np.maximum(t, 0) / ti) - np.exp(-np.maximum(t, 0) / p["TR"])) / (ti - p["TR"])
    pdf[t < 0] = 0.0
    return pdf


def synthetic_objective(p, n_events=3.0e5):
    """sum (data - MC)^2 with MC normalised to data (time_residuals.py:138-142); O(1e7..1e9)."""
    data = emission_hist(TRUE) * n_events
    mc = emission_hist(p)
    mc = mc * data.sum() / mc.sum()
    return float(np.sum((data - mc) ** 2))


class Recorder:
    def __init__(self):
        self.calls = []


def make_recording_class(ref_cls, rec):
    class PointSelector(ref_cls):                      # same name: select_parameters.py:1 imports it
        def update_surrogate(self):
            self._rec = dict(X=np.array(self.measured_pts, dtype=np.float64),
                             y=np.array(self.measured_vals, dtype=np.float64),
                             P=np.array(self.predicted_pts, dtype=np.float64),
                             fd=np.array(self.feature_domain),
                             ls=[np.array(a, dtype=np.float64) for a in
                                 (self.length_scales if len(self.length_scales) == 2 else [self.length_scales])])
            super().update_surrogate()

        def lower_confidence_bound(self, explore=4):
            idx = super().lower_confidence_bound(explore)
            r = self._rec
            r.update(kp=np.array(self.kernel_params, dtype=np.float64), index=np.array(idx),
                     acq_max=float(np.amax(self.acq_func_eval)), mu_min=float(np.amin(self.mean_func)),
                     sig_max=float(np.amax(self.cov_func)))
            rec.calls.append(r)
            return idx
    return PointSelector


@contextlib.contextmanager
def sandbox(scratch):
    real_open = builtins.open

    def remap(path):
        if isinstance(path, str) and path.startswith(HARD):
            return scratch + path[len(HARD):]
        return path

    def fake_open(file, *a, **k):
        return real_open(remap(file), *a, **k)

    cwd = os.getcwd()
    builtins.open = fake_open
    os.chdir(scratch)
    try:
        yield
    finally:
        builtins.open = real_open
        os.chdir(cwd)


def run_script(name):
    """Run an unmodified reference script; returns its exit code (0 if it falls off the end)."""
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(rl.REFERENCE_DIR, name), run_name="__main__")
    except SystemExit as e:
        return int(e.code or 0)
    return 0


def write_back_objective(scratch):
    """time_residuals.py:166-182 (block-best update) and :204-217 (objective into the last row)."""
    with open(os.path.join(scratch, "opto_log.JSON")) as f:
        log = json.load(f)
    objective = synthetic_objective(log["parameters"])
    if objective < log["iteration_info"]["current_block"]["block_best_params"]["obj"]:
        log["parameters"]["obj"] = objective
        log["iteration_info"]["current_block"]["block_best_params"] = log["parameters"]
        with open(os.path.join(scratch, "opto_log.JSON"), "w") as f:
            json.dump(log, f, indent=4)
    cur = log["iteration_info"]["current_block"]["param_sampling"]["current_parameters"]
    algo_iter = log["iteration_info"]["full_algo_iter"]
    block_iter = log["iteration_info"]["current_block"]["iteration"]
    if len(cur) == 2 and cur[0] in (0, 2):
        fname = f"measured_points/{NAMES[cur[0]]}_{NAMES[cur[1]]}_ALGO_{algo_iter}_BLOCK_{block_iter}.npy"
        col = 2
    else:
        fname = f"measured_points/{NAMES[cur[0]]}_ALGO_{algo_iter}_BLOCK_{block_iter}.npy"
        col = 1
    path = os.path.join(scratch, fname)
    if os.path.isfile(path):
        vals = np.load(path)
        vals[-1, col] = objective
        np.save(path, vals)
    return objective


def run_workflow(selector_cls, sizes=(1, 1, 7), max_calls=400, seed=12345):
    """Drive the whole workflow with `selector_cls` installed as `point_selector.PointSelector`.

    sizes = (full-algorithm iterations, block iterations, sample-loop iterations): run_algo.py:7-9 deploys (2, 1, 20).
    Returns dict(files={relative path: bytes}, calls=[recorded PointSelector calls], objectives, events, final)."""
    rec = Recorder()
    mod = types.ModuleType("point_selector")
    mod.PointSelector = make_recording_class(selector_cls, rec)
    saved = sys.modules.get("point_selector")
    sys.modules["point_selector"] = mod
    rl.install_plot_stubs()

    scratch = tempfile.mkdtemp(prefix="closed_loop_")
    for d in ("macros", "measured_points", "plots", "submit_files"):
        os.makedirs(os.path.join(scratch, d))
    shutil.copy(os.path.join(rl.REFERENCE_DIR, "bi214_template.mac"), scratch)
    with open(os.path.join(rl.REFERENCE_DIR, "opto_log.JSON")) as f:
        info = json.load(f)
    info["iteration_info"]["max_iter"] = sizes[0]
    info["iteration_info"]["current_block"]["max_iter"] = sizes[1]
    info["iteration_info"]["current_block"]["param_sampling"]["max_iter"] = sizes[2]
    with open(os.path.join(scratch, "opto_log.JSON"), "w") as f:
        json.dump(info, f, indent=4)

    objectives, events = [], []
    np.random.seed(seed)                       # first-ever point is random (select_parameters.py:219)

    def opto_node():
        while len(rec.calls) < max_calls:
            rc = run_script("select_parameters.py")
            assert rc == 0
            objectives.append(write_back_objective(scratch))
            if run_script("terminate_opto.py") == 0:
                return

    def block(two_stage):
        while len(rec.calls) < max_calls:
            opto_node()
            if two_stage:
                opto_node()
            rc = run_script("terminate_block.py")
            events.append(("block", rc))
            if rc == 0:
                return

    try:
        with sandbox(scratch):
            while len(rec.calls) < max_calls:
                block(True)       # FIRST_PAIR : T1,T2 then A1(,A2)
                block(True)       # SECOND_PAIR: T3,T4 then A3(,A4)
                block(False)      # RISE_TIME  : TR
                rc = run_script("terminate_algo.py")
                events.append(("algo", rc))
                if rc == 0:
                    break
            with open(os.path.join(scratch, "opto_log.JSON")) as f:
                final = json.load(f)
        files = {}
        for base, _dirs, names in os.walk(scratch):
            for nm in names:
                p = os.path.join(base, nm)
                with open(p, "rb") as f:
                    files[os.path.relpath(p, scratch)] = f.read()
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
        if saved is not None:
            sys.modules["point_selector"] = saved
        else:
            sys.modules.pop("point_selector", None)
    return dict(files=files, calls=rec.calls, objectives=objectives, events=events, final=final)
