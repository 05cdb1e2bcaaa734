"""Device time of stan_assemble and stan_recover on one GPU: python tools/asm_time.py [workload] [repeats]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "beam_10m_g2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if name in mesh.WORKLOADS:
    m = mesh.workload(name, tolerance=1e-8)
else:
    f = [int(v) for v in name.split(",")]
    m = mesh.beam(f[0], f[1], f[2], tolerance=1e-8)
with Solver() as s:
    s.SetModel(m); s.AssignDOF()
    rows = []
    for _ in range(reps):
        a = s.ParallelAssembly_K()
        rows.append({"pattern_ms": round(a.pattern_ms, 3), "assembly_kernels_ms": round(a.assembly_ms, 3), "total_ms": round(a.total_ms, 3),
                     "elements_per_s": round(m.n_elem / (a.total_ms * 1e-3)), "kernel_GBs_algorithmic": round(a.assembly_bytes / a.assembly_ms / 1e6, 1),
                     "kernel_TFLOPs_algorithmic": round(a.assembly_flops / a.assembly_ms / 1e9, 2)})
    s.LinearSolver_CG(merit_check=0, IterMax=20)
    rec = [s.Recovery_Stress() for _ in range(reps)]
    print(json.dumps({"workload": name, "n_elem": m.n_elem, "assemble": rows,
                      "recover_ms": [round(r.recover_ms, 3) for r in rec],
                      "recover_GBs_algorithmic": round(rec[-1].recover_bytes / rec[-1].recover_ms / 1e6, 1)}))
