import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
dim, order, B = 32, 20, 16
for dbg in (6, 7, 8, 9):
    os.environ["Z3D_GEN_DBG"] = str(dbg)
    P = Plan(order, dim)
    m = torch.zeros((B, P.num_moments), dtype=torch.complex64, device="cuda")
    m[:, 0] = 1.0 + 0.5j
    m[:, 2] = 0.25 - 1.0j
    got = P.reconstruct(m, batched=True)[0].cpu().numpy()
    print("dbg", dbg, "|got|", float(np.linalg.norm(got)), "nonzero", int(np.count_nonzero(got)), "sample", got[3, 4, 5], got[0, 0, 0], got[16, 17, 20], "uniq", np.unique(np.round(got, 4))[:8])
