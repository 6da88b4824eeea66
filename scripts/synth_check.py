"""Developer check: tensor-core batched synthesis (z3d_reconstruct_batched) vs the per-map kernel, and its time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel().astype(np.float64)) / np.linalg.norm(b.ravel().astype(np.float64)))

for a in sys.argv[1:]:
    dim, order, B = [int(x) for x in a.split(",")][:3]
    P = Plan(order, dim); P.profile(True)
    info = P.info()
    vols = torch.from_numpy(np.stack([volume(200 + i, dim) for i in range(min(B, 8))])).cuda()
    mom8 = P.decompose(vols)
    mom = torch.cat([mom8] * ((B + mom8.shape[0] - 1) // mom8.shape[0]))[:B].contiguous()
    ref = P.reconstruct(mom[:mom8.shape[0]], batched=False).cpu().numpy()
    outs = [P.reconstruct(mom, batched=True) for _ in range(2)]
    torch.cuda.synchronize()
    got = outs[0][:mom8.shape[0]].cpu().numpy()
    errs = [rel(got[i], ref[i]) for i in range(got.shape[0])]
    last = outs[0][B - 1].cpu().numpy()
    same = bool(torch.equal(outs[0], outs[1]))
    ms = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); P.reconstruct(mom, batched=True, out=outs[0]); torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    print(f"dim {dim} order {order} B {B}: synth {info['synth']} sets {info['gen2_sets']} x {info['gen2_splits']}: rel err vs per-map kernel max {max(errs):.2e} "
          f"(last map vs its twin {rel(last, ref[(B - 1) % got.shape[0]]):.2e}) finite {bool(np.isfinite(got).all())} repeatable {same} call {min(ms):.3f} ms = {min(ms) / B * 1e3:.1f} us/map", flush=True)
