"""BASELINE.json config 0 analogue: the reference's "Example 1" (14 550 nodes / 11 470 hex8 / 43 650 DOF,
3 parts, CG with tolerance 1e-6 and no iteration limit — images/Solver.PNG, images/Properties.png) is
not in the checkout, so a generated mesh of the same size class runs the same workflow end to end:

    BDF text + pasted BC rows -> `stan_solver --build` -> STdb (materials, part properties, SPC / loads, analysis)
            -> `stan_solver model.STdb` (GPU, reference defaults incl. ALGLIB's energy-functional stop)
            -> STdb with results

and the CPU oracle solves the same model with the same settings on this box's host cores.  Prints
one JSON line next to the numbers of the reference's published screenshot.

    python tools/example1_like.py [workdir]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402  (checker / CPU baseline only)
from stan_b200 import build, mesh, stdb  # noqa: E402

work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="example1_")
host = build.build_host()
m = mesh.beam(15, 15, 51, n_parts=3, tolerance=1e-6, max_iter=0)           # 11 475 elements, 13 312 nodes
bdf, model, solved = (os.path.join(work, f) for f in ("mesh.bdf", "model.STdb", "solved.STdb"))
spc_txt, load_txt = os.path.join(work, "spc.txt"), os.path.join(work, "load.txt")
mesh.write_bdf(m, bdf)
with open(spc_txt, "w") as fh:                                              # the 4-column paste format of README.md:55
    fh.writelines(f"{n + 1}\t{v[0]:g}\t{v[1]:g}\t{v[2]:g}\n" for n, v in zip(m.spc_node, m.spc_val))
with open(load_txt, "w") as fh:
    fh.writelines(f"{n + 1}\t{v[0]!r}\t{v[1]!r}\t{v[2]!r}\n" for n, v in zip(m.load_node, m.load_val.tolist()))
cmd = [host, "--build", bdf, model, "--spc", spc_txt, "--load", load_txt, "--tol", "1e-6"]
for E, nu in zip(m.mat_E, m.mat_nu):
    cmd += ["--material", repr(float(E)), repr(float(nu))]
for pid in sorted(set(m.elem_pid.tolist())):                                # parts alternate between the two materials
    cmd += ["--part-mat", str(pid), str(int(m.elem_mat[m.elem_pid == pid][0]) + 1)]
r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
assert r.returncode == 0, r.stderr
imported = json.loads(r.stdout)
db = stdb.decode(open(model, "rb").read())                                  # what the native builder produced ...
full = stdb.from_model(m)                                                   # ... is the model the Python side describes
assert [n.id for n in db.nodes] == [n.id for n in full.nodes] and [e.nlist for e in db.elems] == [e.nlist for e in full.elems]
assert [e.matid for e in db.elems] == [e.matid for e in full.elems] and [(x.E, x.poisson) for x in db.mats] == [(x.E, x.poisson) for x in full.mats]
assert [[(n, v.M) for n, v in bc.nodal] for _, bc in db.bcs] == [[(n, v.M) for n, v in bc.nodal] for _, bc in full.bcs]

t0 = time.perf_counter()
r = subprocess.run([host, model, "-o", solved], capture_output=True, text=True, timeout=600)
wall = time.perf_counter() - t0
assert r.returncode == 0, r.stdout + r.stderr
out = r.stdout


def grab(pat):
    mm = re.search(pat, out)
    return mm.groups() if mm else None


asm_s = float(grab(r"K Matrix assembly:\s+Done in ([0-9.]+)s")[0])
typ, solve_s = grab(r"\(type (-?\d+)\) in ([0-9.]+)s")
total_s = float(grab(r"Total CPU time: ([0-9.]+) s")[0])
its = int(grab(r"CG iterations: (\d+)")[0])
dev_start_s = float(grab(r"Device start-up[^:]*: ([0-9.]+) s")[0])

res = stdb.decode(open(solved, "rb").read())
ni, disp, strain, stress = stdb.results(res)

t0 = time.perf_counter()
o = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-6, maxits=0, merit_check=1))
cpu_s = time.perf_counter() - t0
ou = o.U_full.reshape(-1, 3)[o.node_index]
line = {
    "workload": "example1_like 15x15x51 G2, 3 parts / 2 materials, CG tol 1e-6, no iteration limit, ALGLIB merit stop on",
    "n_nodes": m.n_nodes, "n_elem": m.n_elem, "n_dof": m.n_dof, "bdf_import": imported,
    "gpu_cli": {"assembly_s": asm_s, "solve_s": float(solve_s), "terminationtype": int(typ), "cg_iterations": its,
                "total_cpu_time_s": total_s, "of_which_cuda_context_s": dev_start_s, "process_wall_s": wall},
    "cpu_oracle": {"threads": oracle.threads(), "assembly_s": o.stats.t_assembly, "solve_s": o.stats.t_solve,
                   "recover_s": o.stats.t_recovery, "terminationtype": o.stats.cg.terminationtype,
                   "cg_iterations": o.stats.cg.iterationscount, "wall_s": cpu_s},
    "agreement": {"dof_numbering_identical": bool(np.array_equal(ni, o.node_index)),
                  "rel_disp_diff": float(np.linalg.norm(disp - ou) / np.linalg.norm(ou)),
                  "rel_stress_diff": float(np.abs(stress - o.stress).max() / np.abs(o.stress).max())},
    "reference_screenshot": {"source": "images/Solver.PNG (author's desktop, CPU unknown)", "n_nodes": 14550, "n_elem": 11470,
                             "n_dof": 43650, "assembly_s": 4.91, "solve_s": 4.12, "terminationtype": 7, "total_cpu_time_s": 9.79},
}
print(json.dumps(line))
