import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
from zernike3d_b200.indexing import nlm_table
dim, order, B = 64, 50, 16
P = Plan(order, dim)
vols = np.stack([volume(100 + i, dim) for i in range(B)]); dv = torch.from_numpy(vols).cuda()
os.environ["Z3D_GEN"] = "0"; Pref = Plan(order, dim); del os.environ["Z3D_GEN"]
ref = Pref.decompose(dv).cpu().numpy()
nlm = nlm_table(order)
for rep in range(3):
    g = P.decompose(dv).cpu().numpy()
    err = np.abs(g - ref).max(axis=0) / np.abs(ref).max()
    bad = err > 1e-5
    print(f"rep {rep}: {bad.sum()} of {bad.size} moments off; per volume max err {[float('%.1e' % (np.abs(g[b]-ref[b]).max()/np.abs(ref[b]).max())) for b in range(0, B, 5)]}")
    if bad.any():
        ms = sorted(set(nlm[bad, 2])); ls = sorted(set(nlm[bad, 1]))
        print("  m values:", ms[:60]); print("  l values:", ls[:60])
        for m in ms[:6]:
            sel = bad & (nlm[:, 2] == m)
            print(f"   m={m}: l in {sorted(set(nlm[sel,1]))[:30]} n in [{nlm[sel,0].min()},{nlm[sel,0].max()}] count {sel.sum()} of {(nlm[:,2]==m).sum()}")
