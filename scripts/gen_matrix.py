"""Developer check: table-free tensor-core kernel vs per-volume kernel over a matrix of (dim, order, batch[, splits])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

for a in sys.argv[1:]:
    f = [int(x) for x in a.split(",")]
    dim, order, B = f[:3]
    if len(f) > 3:
        os.environ["Z3D_GEN_SPLITS"] = str(f[3])
    else:
        os.environ.pop("Z3D_GEN_SPLITS", None)
    if len(f) > 4:
        os.environ["Z3D_GEN_SLOTS"] = str(f[4])
    else:
        os.environ.pop("Z3D_GEN_SLOTS", None)
    P = Plan(order, dim)
    info = P.info()
    vols = np.stack([volume(100 + i, dim) for i in range(B)])
    dv = torch.from_numpy(vols).cuda()
    os.environ["Z3D_GEN"] = "0"
    Pref = Plan(order, dim)
    del os.environ["Z3D_GEN"]
    ref = Pref.decompose(dv).cpu().numpy()
    outs = [P.decompose(dv).cpu().numpy() for _ in range(4)]
    errs = [max(rel(g[i], ref[i]) for i in range(B)) for g in outs]
    same = all(np.array_equal(outs[0].view(np.float32), g.view(np.float32)) for g in outs[1:])
    nch = info["chunks"] / max(1, info["gen_splits"])
    print(f"dim {dim} order {order} B {B}: sets {info['gen_sets']} x {info['gen_splits']} slots {info['gen_slots']} maxcols {info['gen_maxcols']} "
          f"chunks/CTA {nch:.1f}: err {['%.1e' % e for e in errs]} repeatable {same}", flush=True)
