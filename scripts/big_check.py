import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
def rel(a,b): return float(np.linalg.norm(a-b)/np.linalg.norm(b))
for order, dim, B in [(50, 256, 12), (30, 200, 14), (8, 14, 12), (61, 40, 12)]:
    P = Plan(order, dim); print(P.info())
    vs = torch.from_numpy(np.stack([volume(3 + i % 3, dim) for i in range(B)])).cuda()
    mb = P.decompose(vs); torch.cuda.synchronize()
    for i in (0, B - 1):
        one = P.decompose(vs[i])
        print(f"o{order} d{dim} vol{i}: batched vs per-volume {rel(mb[i].cpu().numpy(), one.cpu().numpy()):.2e}", flush=True)
    e0=torch.cuda.Event(True); e1=torch.cuda.Event(True); e0.record(); P.decompose(vs); e1.record(); torch.cuda.synchronize()
    print(f"   batched call {e0.elapsed_time(e1):.2f} ms for {B} volumes; free mem {torch.cuda.mem_get_info()[0]/2**30:.1f} GiB")
    del P, vs, mb
