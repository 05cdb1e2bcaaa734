"""include/stan_b200.h must be consumable from plain C (the P/Invoke / cgo / FFI side sees a C ABI)."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_compiles_as_c_and_links(tmp_path):
    from stan_b200 import build, native
    build.build()
    src = tmp_path / "use.c"
    src.write_text(textwrap.dedent('''
        #include <stdio.h>
        #include "stan_b200.h"
        int main(void) {
            stan_options o = {-1, 0, 1, 0};
            stan_cg_options cg = {1e-8, 0, 10, 0, 1, 0, 0, 0};
            stan_handle *h = NULL;
            int rc = stan_create(&o, &h);           /* fails without a GPU: that is the point, no fallback */
            printf("%d %d %d %s\\n", stan_version(), rc, (int)sizeof(cg), rc ? stan_last_error() : "ok");
            if (!rc) stan_destroy(h);
            return 0;
        }
    '''))
    exe = tmp_path / "use"
    libdir = os.path.dirname(native.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                           str(exe), "-L", libdir, "-lstan_b200", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    version, rc, size = out.stdout.split()[:3]
    assert version == "100" and size == "40" and rc in ("0", "-2")


def _build_example(tmp_path):
    from stan_b200 import build, native
    build.build()
    exe = tmp_path / "solve_beam"
    libdir = os.path.dirname(native.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "solve_beam.c"), "-o", str(exe), "-L", libdir, "-lstan_b200",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-lm"])
    return exe


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        return                                                    # the GPU run is test_c_example_solves below
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "stan_create" in out.stderr   # no CPU fallback behind the C ABI


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["cg", "cholesky"])
def test_c_example_solves(tmp_path, solver):
    import json
    exe = _build_example(tmp_path)
    out = subprocess.run([str(exe), "4", "4", "40", solver], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout)
    assert d["elements"] == 640 and d["terminationtype"] in (1, 7)
    assert abs(d["tip_ux"] - d["beam_theory"]) / d["beam_theory"] < 0.06       # cantilever against beam theory
    assert d["max_abs_stress"] > 0
