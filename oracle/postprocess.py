"""numpy restatement of Part.Load_Scalar (/root/reference/src/STAN_Database/Part.cs:231-528).

TEST INFRASTRUCTURE ONLY (see oracle/stan_oracle.c).  PARITY UNPINNED: the reference's own numbers
cannot be produced here; eigenvalues come from numpy.linalg.eigvalsh where the reference uses
MathNet's Evd (both return the eigenvalues of a symmetric matrix in ascending order).
"""
import numpy as np


def _tensor(v):                       # Part.cs:325-334: XX YY ZZ XY YZ XZ -> symmetric 3x3
    t = np.zeros(v.shape[:-1] + (3, 3))
    t[..., 0, 0], t[..., 1, 1], t[..., 2, 2] = v[..., 0], v[..., 1], v[..., 2]
    t[..., 0, 1] = t[..., 1, 0] = v[..., 3]
    t[..., 1, 2] = t[..., 2, 1] = v[..., 4]
    t[..., 0, 2] = t[..., 2, 0] = v[..., 5]
    return t


def node_scalars(disp, stress, strain):
    """disp (..., 3), stress/strain (..., 6) -> (..., 24) in the order of Part.cs:268-293."""
    out = np.zeros(disp.shape[:-1] + (24,))
    out[..., 0:3] = disp
    out[..., 3] = np.sqrt((disp ** 2).sum(-1))
    for base, v, scale in ((4, stress, 1.0), (14, strain, 2.0 / 3.0)):
        out[..., base:base + 6] = v
        ev = np.linalg.eigvalsh(_tensor(v))                       # ascending
        p1, p2, p3 = ev[..., 2], ev[..., 1], ev[..., 0]           # Part.cs:335-337
        out[..., base + 6], out[..., base + 7], out[..., base + 8] = p1, p2, p3
        out[..., base + 9] = scale * np.sqrt(((p1 - p2) ** 2 + (p2 - p3) ** 2 + (p3 - p1) ** 2) / 2)
    return out


def load_scalar(model, node_index, U_full, strain, stress):
    """Returns (cell (n_elem,24,3) = max/average/min, point (n_nodes,24)) as float32."""
    disp = U_full.reshape(-1, 3)[node_index]                      # (n_nodes, 3)
    vals = node_scalars(disp[model.conn], stress, strain)         # (n_elem, 8, 24)
    cell = np.stack([vals.max(1), vals.sum(1) / 8, vals.min(1)], axis=-1).astype(np.float32)
    sums = np.zeros((model.n_nodes, 24))
    cnt = np.zeros(model.n_nodes)
    for e in range(model.n_elem):                                 # EList order = ElemLib order (Part.cs:445-447)
        seen = set()
        for i, n in enumerate(model.conn[e]):
            if n in seen:
                continue
            seen.add(n)
            sums[n] += vals[e, i]
            cnt[n] += 1
    point = (sums / cnt[:, None]).astype(np.float32)
    return cell, point
