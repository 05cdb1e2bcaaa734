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


def _header_structs():
    """typedef struct { ... } name;  ->  {name: [(ctype, field), ...]} from include/stan_b200.h."""
    text = open(os.path.join(ROOT, "include", "stan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef struct \{(.*?)\}\s*(stan_[a-z_]+);", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                t, names = decl.split(None, 1)
                fields += [(t, n.strip()) for n in names.split(",")]
        out[name] = fields
    return out


def test_ctypes_structs_mirror_the_header_field_by_field():
    ct = {"int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double}
    pairs = {"stan_options": native.Options, "stan_cg_options": native.CgOptions, "stan_cg_report": native.CgReport,
             "stan_assembly_stats": native.AssemblyStats, "stan_chol_report": native.CholReport,
             "stan_recovery_stats": native.RecoveryStats}
    hs = _header_structs()
    assert set(hs) == set(pairs)
    for name, cls in pairs.items():
        assert [(n, t) for n, t in cls._fields_] == [(f, ct[t]) for t, f in hs[name]], name
    assert C.sizeof(native.CholReport) == 72


def test_csharp_shim_binds_declared_symbols_with_matching_structs():
    """interop/StanNative.cs is source only (no .NET here): at least keep it consistent with the header."""
    cs = open(os.path.join(ROOT, "interop", "StanNative.cs")).read()
    declared = set(_declared_symbols())
    imported = set(re.findall(r"static extern \w+ (stan_[a-z0-9_]+)\(", cs))
    assert imported and imported <= declared, imported - declared
    assert {"stan_assemble", "stan_solve_cg", "stan_solve_cholesky", "stan_recover", "stan_get_displacements",
            "stan_get_strain_stress"} <= imported
    hs = _header_structs()
    cs_structs = {"Options": "stan_options", "CgOptions": "stan_cg_options", "CgReport": "stan_cg_report",
                  "AssemblyStats": "stan_assembly_stats", "CholReport": "stan_chol_report", "RecoveryStats": "stan_recovery_stats"}
    cs_type = {"int": "int32_t", "long": "int64_t", "double": "double"}
    for cs_name, h_name in cs_structs.items():
        m = re.search(r"internal struct %s\s*\{(.*?)\}" % cs_name, cs, flags=re.S)
        assert m, cs_name
        fields = []
        for t, names in re.findall(r"public (int|long|double) ([^;]+);", m.group(1)):
            fields += [(cs_type[t], n.strip()) for n in names.split(",")]
        assert fields == hs[h_name], cs_name
