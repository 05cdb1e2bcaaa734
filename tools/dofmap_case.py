"""AssignDOF timing: serial host traversal vs level-synchronous device traversal on a named workload.

    python tools/dofmap_case.py [workload]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "beam_10m_g2"
m = mesh.workload(name)
out = {"workload": name, "n_nodes": m.n_nodes, "n_elem": m.n_elem}
res = {}
with Solver() as s:
    s.SetModel(m)
    for mode in ("gpu", "host", "gpu"):
        os.environ["STAN_DOF"] = mode
        l0 = s.kernel_launches()
        t0 = time.perf_counter()
        ni = s.AssignDOF()
        out[f"{mode}_s"] = time.perf_counter() - t0
        if mode == "gpu":
            out["gpu_launches"] = s.kernel_launches() - l0
        res[mode] = ni
out["identical"] = bool(np.array_equal(res["gpu"], res["host"]))
print(json.dumps(out))
