"""Developer GPU check of the batched analysis kernel against the per-volume kernel, the oracle and the goldens."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

cases = [("d16_o10", 9), ("d15_o8", 12), ("d32_o20", 32), ("d20_o25", 33), ("c1_d64_o50", 32), ("c5_d96_o60", 16), ("c3_d128_o50", 32)]
if len(sys.argv) > 1:  # names, optionally name:count
    want = dict((a.split(":") + [None])[:2] for a in sys.argv[1:])
    cases = [(n, int(want[n]) if want[n] else b) for n, b in cases if n in want]
for name, B in cases:
    g = np.load(os.path.join(ROOT, "tests/golden", f"ref_{name}.npz"))
    order, dim = int(g["order"]), int(g["dim"])
    P = Plan(order, dim)
    info = P.info()
    vs = np.stack([volume(int(g["seeds"][0]) + i, dim) for i in range(B)])
    dv = torch.from_numpy(vs).cuda()
    mb = P.decompose(dv); torch.cuda.synchronize()          # batched route (B >= batched_min)
    ms = torch.stack([P.decompose(dv[i:i + 1])[0] for i in range(B)])   # per-volume route
    torch.cuda.synchronize()
    line = f"{name} B={B} batched_min={info['batched_min']} groups={info['batched_row_groups']}/{info['batched_row_groups_wide']} wide>={info['batched_wide_min']}: "
    line += f"batched vs per-volume {rel(mb.cpu().numpy(), ms.cpu().numpy()):.2e} (worst volume {max(rel(mb[i].cpu().numpy(), ms[i].cpu().numpy()) for i in range(B)):.2e})"
    line += f" vs ref golden {rel(mb[0].cpu().numpy(), g['moments'][0]):.2e}"
    if dim <= 64:
        import oracle
        ex = oracle.decompose(vs[B - 1], order, oracle.EXACT)
        line += f" vs exact(last) {rel(mb[B - 1].cpu().numpy(), ex):.2e} per-vol {rel(ms[B - 1].cpu().numpy(), ex):.2e}"
    tb = timed(lambda: P.decompose(dv))
    t1 = timed(lambda: P.decompose(dv[:4]))
    line += f" | batched {tb:.3f} ms ({tb / B * 1e3:.1f} us/vol), per-volume kernel {t1 / 4 * 1e3:.1f} us/vol"
    print(line, flush=True)
    del P
