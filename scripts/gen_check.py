"""Developer check of the table-free tensor-core kernel (k_decompose_gen) against the per-volume kernel; timing of one tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

cases = [(16, 10, 13), (15, 8, 12), (32, 20, 35), (20, 25, 9), (64, 50, 16), (64, 50, 70)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for dim, order, B in cases:
    P = Plan(order, dim)
    info = P.info()
    vols = np.stack([volume(100 + i, dim) for i in range(B)])
    dv = torch.from_numpy(vols).cuda()
    os.environ["Z3D_GEN"] = "0"
    Pref = Plan(order, dim)
    del os.environ["Z3D_GEN"]
    ref = Pref.decompose(dv).cpu().numpy()
    got = P.decompose(dv)
    torch.cuda.synchronize()
    got2 = P.decompose(dv)
    g = got.cpu().numpy()
    errs = [rel(g[i], ref[i]) for i in range(B)]
    print(f"dim {dim} order {order} batch {B}: gen={info['gen']} sets {info['gen_sets']} x splits {info['gen_splits']} slots {info['gen_slots']} "
          f"maxcols {info['gen_maxcols']}: rel-L2 vs per-volume kernel max {max(errs):.3e} mean {np.mean(errs):.3e}; "
          f"bitwise repeatable {bool(torch.equal(torch.view_as_real(got), torch.view_as_real(got2)))} finite {bool(np.isfinite(g.view(np.float32)).all())}",
          flush=True)
