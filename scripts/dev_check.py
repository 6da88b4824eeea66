"""Developer GPU check: CUDA path vs the CPU oracle + reference goldens, with timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle
from zernike3d_b200.plan import Plan, fma_peak_tflops
from zernike3d_b200.synthetic import volume

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))

print(torch.cuda.get_device_name(0))
for variant in (0, 1):
    print("fma peak variant", variant, [round(fma_peak_tflops(0, variant, 8192), 2) for _ in range(3)], "TFLOP/s")

for name in ["d16_o10", "d15_o8", "d32_o20", "d20_o25", "c1_d64_o50", "c3_d128_o50", "c5_d96_o60"]:
    g = np.load(os.path.join(ROOT, "tests/golden", f"ref_{name}.npz"))
    order, dim = int(g["order"]), int(g["dim"])
    P = Plan(order, dim)
    if name == "d16_o10": print(P.info())
    v = volume(int(g["seeds"][0]), dim)
    dv = torch.from_numpy(v).cuda()
    mom = P.decompose(dv); torch.cuda.synchronize()
    mom = mom.cpu().numpy()
    line = f"{name}: mom vs ref {rel(mom, g['moments'][0]):.2e}"
    if dim <= 64:
        ex = oracle.decompose(v, order, oracle.EXACT)
        line += f" vs exact {rel(mom, ex):.2e}"
    inv = P.invariants(torch.from_numpy(mom).cuda()).cpu().numpy()
    line += f" inv vs ref {rel(inv, g['invariants'][0]):.2e}"
    if "recon" in g.files:
        rec = P.reconstruct(torch.from_numpy(g["moments"][0]).cuda()); torch.cuda.synchronize()
        line += f" rec vs ref {rel(rec.cpu().numpy(), g['recon'][0]):.2e}"
    # slab additivity
    h = P.half
    m2 = P.decompose(dv, pairs=(0, h // 2)) + P.decompose(dv, pairs=(h // 2, h))
    line += f" slabs {rel(m2.cpu().numpy(), mom):.2e}"
    # timing
    B = 4
    dvb = dv.unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
    for _ in range(2): P.decompose(dvb)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record(); 
    for _ in range(3): P.decompose(dvb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3 / B
    W = 4.0 * P.num_moments * dim ** 3
    line += f" | decompose {ms:.3f} ms/vol  W/t={W/ms/1e9:.1f} TFLOP/s"
    dm = torch.from_numpy(g["moments"][0]).cuda().unsqueeze(0).repeat(B, 1).contiguous()
    for _ in range(2): P.reconstruct(dm)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): P.reconstruct(dm)
    e1.record(); torch.cuda.synchronize()
    line += f" reconstruct {e0.elapsed_time(e1)/3/B:.3f} ms/vol"
    print(line, flush=True)
    del P
# host API
P = Plan(50, 64)
vs = np.stack([volume(s, 64) for s in range(3)])
mh = P.decompose_host(vs)
md = P.decompose(torch.from_numpy(vs).cuda()).cpu().numpy()
print("host api decompose == device api:", np.array_equal(mh, md))
rh = P.reconstruct_host(mh)
rd = P.reconstruct(torch.from_numpy(mh).cuda()).cpu().numpy()
print("host api reconstruct == device api:", np.array_equal(rh, rd))
