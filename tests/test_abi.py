"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/stan_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from stan_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "stan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stan_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stan_b200 import build
    build.build()
    lib = C.CDLL(native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/stan_b200.h but not exported"
    assert sorted(native.SYMBOLS) == declared        # the ctypes table mirrors the header exactly


def test_struct_layouts_match_header():
    assert C.sizeof(native.Options) == 16
    assert C.sizeof(native.CgOptions) == 40
    assert C.sizeof(native.CgReport) == 72
    assert C.sizeof(native.AssemblyStats) == 88
    assert C.sizeof(native.RecoveryStats) == 24


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is only observable without one")
    lib = native.load()
    h = C.c_void_p()
    o = native.Options(-1, 0, 1, 0)
    rc = lib.stan_create(C.byref(o), C.byref(h))
    assert rc == native.E_CUDA and b"cuda" in lib.stan_last_error().lower()
    from stan_b200.solver import Solver
    with pytest.raises(native.StanError):
        Solver()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "stan_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "build.py", f"{f} mentions the oracle"
