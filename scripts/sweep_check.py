"""Batched (tensor-core) path against the per-volume kernel over a sweep of orders and box sizes (edge cases of the planner)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
worst = 0.0
for order in (0, 1, 2, 3, 5, 13, 31, 47, 64, 70, 74):
    for dim, B in ((14, 12), (15, 13), (17, 33), (33, 70), (50, 12)):
        if order > 64 and dim > 33: continue
        P = Plan(order, dim)
        info = P.info()
        vs = torch.from_numpy(np.stack([volume(11 + i, dim) for i in range(B)])).cuda()
        mb = P.decompose(vs).cpu().numpy()
        ms = np.stack([P.decompose(vs[i]).cpu().numpy() for i in (0, B - 1)])
        e = max(rel(mb[0], ms[0]), rel(mb[B - 1], ms[1]))
        worst = max(worst, e)
        tag = "batched" if info["batched_min"] else "per-volume only"
        print(f"order {order:3d} dim {dim:3d} B {B:2d} [{tag}, groups {info['batched_row_groups']}/{info['batched_row_groups_wide']}]: {e:.2e}", flush=True)
        assert np.isfinite(mb.view(np.float32)).all() and e < 4e-6
        del P
print("SWEEP OK, worst", worst)
