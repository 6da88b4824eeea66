import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/oracle")
import numpy as np, torch, oracle
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
dim, order = int(sys.argv[1]), int(sys.argv[2])
v = volume(1, dim); P = Plan(order, dim); P.profile(True)
m = P.decompose(torch.from_numpy(v).cuda()).cpu().numpy(); ms = P.last_kernel_ms()
t = time.time(); ex = oracle.decompose(v, order, oracle.EXACT); dt = time.time() - t
h = P.half
parts = (P.decompose(torch.from_numpy(v).cuda(), pairs=(0, h // 3)) + P.decompose(torch.from_numpy(v).cuda(), pairs=(h // 3, h))).cpu().numpy()
print(f"{dim}^3 o{order}: kernel {ms:.3f} ms; vs exact {rel(m, ex):.2e}; inv vs exact {rel(oracle.invariants(m, order), oracle.invariants(ex, order)):.2e}; slabs-vs-full {rel(parts, m):.2e}; slabs vs exact {rel(parts, ex):.2e} (oracle {dt:.0f}s, {oracle.max_threads()} thr)")
