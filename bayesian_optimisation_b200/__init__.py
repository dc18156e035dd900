"""B200-native Gaussian-process surrogate + acquisition hot path (drop-in for the reference's
`point_selector.PointSelector`).  Importing the package does not touch the GPU; the CUDA
library (csrc/libbogp.so) is loaded on first use and there is no CPU fallback."""
from ._lib import ACQ_EI, ACQ_LCB, BogpError  # noqa: F401

__all__ = ["PointSelector", "GPEngine", "CandidateGrid", "ACQ_LCB", "ACQ_EI", "BogpError"]


def __getattr__(name):
    if name == "PointSelector":
        from .point_selector import PointSelector
        return PointSelector
    if name in ("GPEngine", "CandidateGrid", "default_engine"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
