"""Developer instrumentation run: library built with -DZ3D_GEN_TIMING [-DGT_BLOCK=n]; clocks per role and wait reason of one CTA."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from zernike3d_b200 import _lib
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
dim, order = int(sys.argv[1]), int(sys.argv[2])
P = Plan(order, dim); P.profile(True)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
v = torch.from_numpy(volume(1, dim)).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
for _ in range(2): P.decompose(v)
L = _lib.lib()
buf = (ctypes.c_longlong * 32)()
L.z3d_debug_gen(buf, 1)
P.decompose(v); ms = P.last_kernel_ms()
L.z3d_debug_gen(buf, 1)
i = P.info()
nch = i["chunks"] // (i["gen2_splits"] if B > 64 and i["gen2"] else i["gen_splits"])
names = ["issuer wait flushed", "issuer wait tfull", "issuer wait afull", "issuer issue", "seed0 wait gfree", "seed0 work (per its chunks)",
         "A0 wait gfull", "A0 wait afree", "A0 work", "A0 flush", "T0 wait gfull", "T0 wait tfree", "T0 work", "T0 flush",
         "A0 fold read + panels + tcgen05.st issue", "A0 tcgen05.wait::st", "A0 fence + arrive",
         "issuer tcgen05.fence::after", "issuer MMA issue (elected lane)", "issuer commits (elected lane)", "issuer elect block + syncwarp"]
print(f"kernel {ms:.3f} ms, {nch} chunks per CTA, {ms * 1e-3 * 1.965e9 / nch:.0f} clk per chunk at 1965 MHz")
for n, c in zip(names, list(buf)):
    print(f"  {n:32s} {c:12d} clk = {c / nch:8.1f} per chunk of the CTA")
