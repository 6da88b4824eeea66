"""Small run of the tensor-core kernels for compute-sanitizer (memcheck): table-free batched analysis and batched synthesis."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
for order, dim, B in [(10, 16, 9), (12, 20, 13)]:
    P = Plan(order, dim)
    v = torch.from_numpy(np.stack([volume(s, dim) for s in range(B)])).cuda()
    m = P.decompose(v)                       # k_fold_tile_gen + k_decompose_gen + k_reduce_stream
    m12 = torch.cat([m, m])[:max(B, 12)].contiguous()
    r = P.reconstruct(m12, batched=True)     # k_prepare_synth + k_synth_gen (clusters) + k_unfold_tile
    torch.cuda.synchronize()
    print(order, dim, B, float(m.abs().sum()), float(r.abs().sum()), flush=True)
print("SANITIZE_TC_DONE")
