"""Developer aid: timing ablations of k_synth_gen (library built with -DZ3D_SYNTH_DEBUG): bits 1 no MMAs, 2 no T work, 4 no epilogue work, 8 no seed work."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from zernike3d_b200.plan import Plan
dim, order, B = 128, 50, 64
for dbg in (0, 1, 2, 4, 8, 3, 7, 15):
    os.environ["Z3D_GEN_DBG"] = str(dbg)
    P = Plan(order, dim)
    m = torch.randn((B, P.num_moments), dtype=torch.complex64, device="cuda")
    out = P.reconstruct(m, batched=True)
    ms = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); P.reconstruct(m, batched=True, out=out); torch.cuda.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    print(f"dbg {dbg:2d}: {min(ms):.3f} ms per 64 maps (k_unfold_tile ~1.9 ms of it)", flush=True)
    del P, out
