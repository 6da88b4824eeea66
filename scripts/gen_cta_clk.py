"""Developer instrumentation run: library built with -DZ3D_GEN_CTACLK; duration of every CTA (column set x split) of one k_decompose_gen
launch, next to the planner's description of the sets (Z3D_GEN_DUMP on stderr)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["Z3D_GEN_DUMP"] = "1"
import torch
from zernike3d_b200 import _lib
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
dim, order = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
P = Plan(order, dim); P.profile(True)
L = _lib.lib()
out = (ctypes.c_int * 11)()
L.z3d_gen_layout(order, 148, out, 11)          # prints the sets (stderr)
v = torch.from_numpy(volume(1, dim)).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
for _ in range(3): P.decompose(v)
ms = P.last_kernel_ms()
i = P.info()
n = i["gen_sets"] * i["gen_splits"]
buf = (ctypes.c_longlong * (5 * 1024))()
L.z3d_debug_gen_cta(buf, 5 * 1024)
nch = i["chunks"] // i["gen_splits"]
print(f"kernel {ms:.3f} ms; {i['gen_sets']} sets x {i['gen_splits']} splits; {nch} chunks per CTA; est period {out[10]}")
print("per chunk of split 0: CTA duration | issuing warp: wait flushed, wait tfull, wait afull, elected MMA + commit block | rest")
for s in range(i["gen_sets"]):
    c = [buf[s * i["gen_splits"] + k] for k in range(i["gen_splits"])]
    b = s * i["gen_splits"]
    w = [buf[k * 1024 + b] / nch for k in range(5)]
    print(f"set {s:3d}: " + " ".join(f"{x / nch:7.1f}" for x in c) + f"  clk per chunk | {w[1]:6.1f} {w[2]:6.1f} {w[3]:6.1f} {w[4]:6.1f} | {w[0] - sum(w[1:]):6.1f}")
