"""Drop-in module: put this directory first on sys.path (or copy this file next to
select_parameters.py, as the deployment of executables/select_parameters.sh:4-5 does) and the
unmodified `from point_selector import PointSelector` of select_parameters.py:1 picks up the
B200 implementation."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from bayesian_optimisation_b200.point_selector import PointSelector  # noqa: E402,F401
