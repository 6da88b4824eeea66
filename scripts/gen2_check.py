"""Developer check: two-tile table-free kernel (more than 64 volumes per call) vs the per-volume kernel, and its time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))

for a in sys.argv[1:]:
    dim, order, B = [int(x) for x in a.split(",")][:3]
    P = Plan(order, dim); P.profile(True)
    info = P.info()
    vols = np.stack([volume(100 + i, dim) for i in range(B)])
    dv = torch.from_numpy(vols).cuda()
    os.environ["Z3D_GEN"] = "0"
    Pref = Plan(order, dim)
    del os.environ["Z3D_GEN"]
    ref = Pref.decompose(dv).cpu().numpy()
    outs = [P.decompose(dv).cpu().numpy() for _ in range(3)]
    errs = [max(rel(g[i], ref[i]) for i in range(B)) for g in outs]
    same = all(np.array_equal(outs[0].view(np.float32), g.view(np.float32)) for g in outs[1:])
    ms = []
    for _ in range(4):
        P.decompose(dv); ms.append(P.last_kernel_ms())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): P.decompose(dv)
    torch.cuda.synchronize(); call = (time.perf_counter() - t0) / 5 * 1e3
    print(f"dim {dim} order {order} B {B}: gen2 {info['gen2']} sets {info['gen2_sets']} x {info['gen2_splits']} slots {info['gen2_slots']} | gen1 {info['gen_sets']} x {info['gen_splits']}: "
          f"err {['%.1e' % e for e in errs]} repeatable {same} kernel(s) {min(ms):.3f} ms call {call:.3f} ms = {call / B * 1e3:.1f} us/vol", flush=True)
