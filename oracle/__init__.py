"""TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy) of the reference's Gaussian-process hot path
(`/root/reference/point_selector.py:13-207`).  Only `tests/`,
`__graft_entry__.smoke()` and the CPU-baseline / reference legs of `bench.py`
may import this package; the product (`bayesian_optimisation_b200`) never does.
"""
