"""stan_b200 — B200-native linear-static hot path of galuszkm/STAN behind a C ABI.

`stan_b200.native` loads libstan_b200.so (CUDA, sm_100a); `stan_b200.solver.Solver` mirrors the
reference's solver interface; `stan_b200.mesh` generates the synthetic workloads of SURVEY.md §8d.
"""
from . import mesh  # noqa: F401

__all__ = ["mesh"]
