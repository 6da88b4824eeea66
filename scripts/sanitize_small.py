import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
for order, dim, B in [(10, 16, 2), (8, 15, 1), (50, 20, 1), (61, 18, 1)]:
    P = Plan(order, dim)
    v = torch.from_numpy(np.stack([volume(s, dim) for s in range(B)])).cuda()
    m = P.decompose(v)
    m2 = P.decompose(v, pairs=(1, P.half - 1))
    m3 = P.decompose(v, part=(1, 3))
    r = P.reconstruct(m)
    r2 = P.reconstruct(m, pairs=(0, 2))
    inv = P.invariants(m)
    torch.cuda.synchronize()
    print(order, dim, float(m.abs().sum()), float(r.abs().sum()), flush=True)
print("SANITIZE_RUN_DONE")
