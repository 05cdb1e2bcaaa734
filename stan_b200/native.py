"""ctypes binding of libstan_b200.so (include/stan_b200.h).

The library is the product; this module only loads it and declares the ABI.  There is no CPU
fallback: if the CUDA library is missing or no GPU is visible every compute entry raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstan_b200.so")

OK, E_ARG, E_CUDA, E_SINGULAR, E_STATE, E_DOFMAP, E_COMM, E_NOMEM = 0, -1, -2, -3, -4, -6, -7, -8


class StanError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libstan_b200 error {code}: {message}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("flags", C.c_int32)]


class CgOptions(C.Structure):
    _fields_ = [("epsf", C.c_double), ("maxits", C.c_int32), ("its_before_rupdate", C.c_int32),
                ("its_before_restart", C.c_int32), ("merit_check", C.c_int32), ("zero_based_counter", C.c_int32),
                ("time_kernels", C.c_int32), ("reserved", C.c_int32)]


class CgReport(C.Structure):
    _fields_ = [("terminationtype", C.c_int32), ("iterationscount", C.c_int32), ("nmv", C.c_int32),
                ("spmv_launches", C.c_int32), ("r2", C.c_double), ("bnorm", C.c_double), ("solve_ms", C.c_double),
                ("spmv_ms", C.c_double), ("spmv_bytes", C.c_int64), ("iter_bytes", C.c_int64),
                ("kernel_launches", C.c_int64)]


class AssemblyStats(C.Structure):
    _fields_ = [("n_dof", C.c_int64), ("n_fixed", C.c_int64), ("n_rows_local", C.c_int64),
                ("n_blocks_local", C.c_int64), ("nnz_upper", C.c_int64), ("assembly_bytes", C.c_int64),
                ("assembly_flops", C.c_double), ("pattern_ms", C.c_double), ("assembly_ms", C.c_double),
                ("total_ms", C.c_double), ("kernel_launches", C.c_int64)]


class CholReport(C.Structure):
    _fields_ = [("terminationtype", C.c_int32), ("block", C.c_int32), ("n", C.c_int64), ("n_blocks", C.c_int64),
                ("skyline_bytes", C.c_int64), ("flops", C.c_double), ("setup_ms", C.c_double),
                ("factor_ms", C.c_double), ("solve_ms", C.c_double), ("kernel_launches", C.c_int64)]


class RecoveryStats(C.Structure):
    _fields_ = [("recover_ms", C.c_double), ("recover_bytes", C.c_int64), ("kernel_launches", C.c_int64)]


# every symbol include/stan_b200.h declares: name -> (restype, argtypes)
_P, _I64, _I32 = C.c_void_p, C.c_int64, C.c_int32
SYMBOLS = {
    "stan_last_error": (C.c_char_p, []),
    "stan_version": (C.c_int, []),
    "stan_create": (C.c_int, [C.POINTER(Options), C.POINTER(_P)]),
    "stan_destroy": (C.c_int, [_P]),
    "stan_set_mesh": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _P]),
    "stan_set_materials": (C.c_int, [_P, _I32, _P, _P]),
    "stan_set_dof_map": (C.c_int, [_P, _P]),
    "stan_assign_dof": (C.c_int, [_P, _P]),
    "stan_set_spc": (C.c_int, [_P, _I64, _P, _P]),
    "stan_set_loads": (C.c_int, [_P, _I64, _P, _P]),
    "stan_assemble": (C.c_int, [_P, C.POINTER(AssemblyStats)]),
    "stan_solve_cg": (C.c_int, [_P, C.POINTER(CgOptions), C.POINTER(CgReport)]),
    "stan_solve_cholesky": (C.c_int, [_P, C.POINTER(CholReport)]),
    "stan_recover": (C.c_int, [_P, C.POINTER(RecoveryStats)]),
    "stan_get_displacements": (C.c_int, [_P, _P]),
    "stan_get_displacements_local": (C.c_int, [_P, _P]),
    "stan_get_node_displacements": (C.c_int, [_P, _P]),
    "stan_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "stan_host_free": (C.c_int, [_P]),
    "stan_get_strain_stress": (C.c_int, [_P, _P, _P]),
    "stan_get_element_range": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "stan_postprocess": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "stan_get_scalars": (C.c_int, [_P, _P, _P]),
    "stan_get_dof_reduction": (C.c_int, [_P, _P]),
    "stan_get_rhs": (C.c_int, [_P, _P]),
    "stan_get_solution_reduced": (C.c_int, [_P, _P]),
    "stan_get_csr_upper_size": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "stan_get_csr_upper": (C.c_int, [_P, _P, _P, _P]),
    "stan_element_stiffness": (C.c_int, [_P, _I64, _I64, _P]),
    "stan_spmv": (C.c_int, [_P, _P, _P]),
    "stan_time_spmv": (C.c_int, [_P, _I32, C.POINTER(C.c_double), C.POINTER(_I64)]),
    "stan_set_cg_history": (C.c_int, [_P, _I32]),
    "stan_get_cg_history": (C.c_int, [_P, C.POINTER(_I32), _P]),
    "stan_kernel_launches": (_I64, [_P]),
    "stan_event_record": (C.c_int, [_P, _I32]),
    "stan_event_elapsed": (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_double)]),
    "stan_comm_unique_id": (C.c_int, [_P]),
    "stan_comm_init": (C.c_int, [_P, _P]),
    "stan_get_partition": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
}

_lib = None


def load():
    """Loads the in-tree CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m stan_b200.build` (the CUDA library is the "
                              "only implementation; there is no CPU path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != OK:
        raise StanError(rc, load().stan_last_error().decode(errors="replace"))
