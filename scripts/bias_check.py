"""How much of the batched (tensor-core) path's error is a uniform shrink (truncating accumulation)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for order, dim, B in [(20, 32, 32), (25, 20, 32), (50, 64, 32), (60, 48, 16), (50, 128, 32)]:
    P = Plan(order, dim)
    vs = np.stack([volume(7 + i, dim) for i in range(B)])
    mb = P.decompose(torch.from_numpy(vs).cuda()).cpu().numpy()
    for i in (0, B - 1):
        ex = oracle.decompose(vs[i], order, oracle.EXACT, return_f64=True)
        m = mb[i].astype(np.complex128)
        alpha = (np.vdot(ex, m) / np.vdot(ex, ex)).real - 1.0
        resid = rel(m / (1.0 + alpha), ex)
        inv_b = oracle.invariants(mb[i], order); inv_e = oracle.invariants(ex.astype(np.complex64), order)
        print(f"o{order} d{dim} vol{i}: err {rel(m, ex):.2e}  shrink alpha {alpha:+.2e}  residual after unshrink {resid:.2e}  inv err {rel(inv_b, inv_e):.2e}", flush=True)
