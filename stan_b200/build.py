"""Builds stan_b200/lib/libstan_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m stan_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libstan_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["api.cu", "pattern.cu", "assembly.cu", "cg.cu", "cholesky.cu", "recovery.cu", "postprocess.cu", "comm.cu", "dofmap_gpu.cu", "dofmap.cpp"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
         "-fPIC,-fvisibility=hidden,-O2", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _deps():
    hdr = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(os.path.dirname(HERE), "include", "stan_b200.h"))
    return hdr


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    deps = _deps()

    def compile_one(name):
        src = os.path.join(SRC, name)
        obj = os.path.join(OBJ_DIR, name.rsplit(".", 1)[0] + ".o")
        if not force and not _stale(obj, [src] + deps):
            return obj, ""
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {name}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(OUT_DIR, "ptxas.log"), "w") as fh:
            fh.write(log)
        if verbose:
            sys.stderr.write(log)
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


HOST_SRC = os.path.join(HERE, "host")
HOST_BIN = os.path.join(OUT_DIR, "stan_solver")
CXX = "/usr/bin/g++"   # the image's $CXX wrapper lacks some specs; the system g++ is complete


def build_host(force: bool = False) -> str:
    """Native solver host (STdb codec, BDF import, console driver) linked against libstan_b200.so."""
    lib = build(force=False)
    srcs = [os.path.join(HOST_SRC, f) for f in ("stan_solver.cpp", "stdb.cpp", "bdf.cpp", "vtu.cpp", "model_build.cpp")]
    deps = srcs + [os.path.join(HOST_SRC, f) for f in os.listdir(HOST_SRC) if f.endswith(".hpp")] + [lib]
    if force or _stale(HOST_BIN, deps):
        cmd = [CXX, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", HOST_BIN] + srcs + \
              ["-L" + OUT_DIR, "-lstan_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + "/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"host build failed:\n{r.stdout}\n{r.stderr}")
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_host(force="--force" in sys.argv))
