"""Minimal single-config run for ncu: python scripts/prof_run.py [dim] [order] [batch] [reps] [what]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
order = int(sys.argv[2]) if len(sys.argv) > 2 else 50
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
what = sys.argv[5] if len(sys.argv) > 5 else "both"
P = Plan(order, dim)
v = torch.from_numpy(volume(1, dim)).cuda().unsqueeze(0).repeat(batch, 1, 1, 1).contiguous()
m = None
for _ in range(reps):
    if what in ("both", "dec"):
        m = P.decompose(v)
    if what in ("both", "rec"):
        if m is None:
            m = torch.randn(batch, P.num_moments, dtype=torch.complex64, device="cuda")
        r = P.reconstruct(m)
torch.cuda.synchronize()
print("ok", float(m.abs().sum()))
