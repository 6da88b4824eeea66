import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["Z3D_LIB_OVERRIDE"] = os.path.join(ROOT, "scripts/_abl/lib_timing.so")
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200 import _lib
from zernike3d_b200.synthetic import volume
dim, order = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
P = Plan(order, dim); P.profile(True)
v = torch.from_numpy(volume(1, dim)).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
L = _lib.lib()
buf = (ctypes.c_longlong * 8)()
for _ in range(2): P.decompose(v)
L.z3d_debug_dump(buf, 1)
P.decompose(v); ms = P.last_kernel_ms()
L.z3d_debug_dump(buf, 1)
names = ["wait barrier B", "geo + wait MMA", "flush", "T and a panels", "wait barrier A", "issue MMA + stage", "fold", "recurrences + trig"]
info = P.info()
nw = info["batched_row_groups"] * info["batched_splits"] * 32
tot = sum(buf[i] for i in range(8))
print(f"kernel {ms:.3f} ms = {ms*1.965e6:.0f} clk @1965MHz; warps {nw}")
for i, n in enumerate(names):
    print(f"  {n:16s} {buf[i]/nw:10.0f} clk/warp  {buf[i]/tot*100:5.1f}%")
