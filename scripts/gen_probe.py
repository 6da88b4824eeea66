"""Developer aid (-DZ3D_GEN_DEBUG builds): run the table-free kernel under one hazard probe (Z3D_GEN_DBG) and report error rates."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from zernike3d_b200 import _lib
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume

dim, order, B = (int(x) for x in sys.argv[1].split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = _lib.lib()
armed = hasattr(L, "z3d_debug_trap_arm") and L.z3d_debug_trap_arm() == 0
vols = np.stack([volume(100 + i, dim) for i in range(B)])
dv = torch.from_numpy(vols).cuda()
os.environ["Z3D_GEN"] = "0"
Pref = Plan(order, dim)
del os.environ["Z3D_GEN"]
ref = Pref.decompose(dv).cpu().numpy()
P = Plan(order, dim)
scale = np.abs(ref).max()
try:
    res = []
    for r in range(reps):
        g = P.decompose(dv).cpu().numpy()
        bad = (np.abs(g - ref).max(axis=0) / scale) > 1e-5
        res.append(int(bad.sum()))
    print(f"DBG={os.environ.get('Z3D_GEN_DBG', '0'):>4s} SPLITS={os.environ.get('Z3D_GEN_SPLITS', '-')} dim {dim} order {order} B {B}: moments off per rep {res}", flush=True)
except Exception as e:
    print(f"DBG={os.environ.get('Z3D_GEN_DBG', '0')} FAILED: {str(e).splitlines()[0]}", flush=True)
    if armed:
        buf = (ctypes.c_int * 512)()
        L.z3d_debug_trap_read(buf, 512)
        n = min(buf[0], 60)
        print(f"  {buf[0]} warps timed out; first records (wait id, block, warp, parity, chunk):")
        for i in range(min(n, 40)):
            print("   ", list(buf[8 + 8 * i: 8 + 8 * i + 5]))
    os._exit(0)
