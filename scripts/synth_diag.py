import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.indexing import nlm_table
dim, order, B = 32, 20, 16
P = Plan(order, dim)
nlm = nlm_table(order)
for mu in (0, 1, 2, 5, 40, 300):
    m = torch.zeros((B, P.num_moments), dtype=torch.complex64, device="cuda")
    m[:, mu] = 1.0 + 0.5j
    ref = P.reconstruct(m[:1], batched=False)[0].cpu().numpy()
    got = P.reconstruct(m, batched=True)[0].cpu().numpy()
    num = float((got.astype(np.float64) * ref).sum()); den = float((ref.astype(np.float64) ** 2).sum())
    print("mu", mu, "nlm", nlm[mu], "|ref|", np.linalg.norm(ref), "|got|", np.linalg.norm(got), "got.ref/ref.ref", num / den if den else None,
          "ref[3,4,5]", ref[3, 4, 5], "got[3,4,5]", got[3, 4, 5], "nonzero got", int(np.count_nonzero(got)))
