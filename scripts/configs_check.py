"""Times the BASELINE.json configurations on one GPU (device-resident inputs, CUDA events): C1/C2 64^3 order 50 decomposition and
reconstruction, C3 128^3 single volume, C4 3DVA bundle (6 x 128^3 in one call), C5a share of one GPU (128 x 96^3 at order 60),
C5b 256^3 single volume."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def vols(n, dim, base=0):
    v = np.stack([volume(base + i, dim) for i in range(min(n, 8))])
    return torch.from_numpy(np.concatenate([v] * (-(-n // len(v))))[:n].copy()).cuda()

P = Plan(50, 64); v = vols(1, 64); m = P.decompose(v)
print(f"C1  64^3 order 50, one volume: decompose {timed(lambda: P.decompose(v)):.3f} ms; C2 reconstruct {timed(lambda: P.reconstruct(m)):.3f} ms", flush=True)
P = Plan(50, 128); v = vols(1, 128)
print(f"C3  128^3 order 50, one volume: decompose {timed(lambda: P.decompose(v)):.3f} ms", flush=True)
v = vols(6, 128)
print(f"C4  3DVA bundle, 6 x 128^3 order 50 in one call: {timed(lambda: P.decompose(v)):.3f} ms", flush=True)
v = vols(128, 128)
t = timed(lambda: P.decompose(v), 3)
print(f"    bench step, 128 x 128^3 order 50: {t:.3f} ms = {128 / t * 1e3:.0f} volumes/s", flush=True)
del P, v
P = Plan(60, 96); v = vols(128, 96)
t = timed(lambda: P.decompose(v), 3)
print(f"C5a 128 x 96^3 order 60 (one GPU's share of 1024): {t:.3f} ms = {128 / t * 1e3:.0f} volumes/s; plan {P.info()['stream_sets']} sets x {P.info()['stream_splits']} splits, T table {P.info()['stream_table_mib']} MiB", flush=True)
del P, v
P = Plan(50, 256); v = vols(1, 256)
print(f"C5b 256^3 order 50, one volume: decompose {timed(lambda: P.decompose(v), 3):.3f} ms", flush=True)
