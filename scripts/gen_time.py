import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from zernike3d_b200.plan import Plan
from zernike3d_b200.synthetic import volume
dim, order = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
P = Plan(order, dim); P.profile(True)
v = torch.from_numpy(volume(1, dim)).cuda().unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
for _ in range(3): P.decompose(v)
ms = []
for _ in range(5):
    P.decompose(v); ms.append(P.last_kernel_ms())
i = P.info()
print(f"dim {dim} order {order} B {B}: gen2 {i["gen2_sets"]} x {i["gen2_splits"]} gen sets {i['gen_sets']} x {i['gen_splits']} slots {i['gen_slots']}: contraction kernel {min(ms):.3f} ms per {B} volumes = {min(ms) / B * 1e3:.1f} us/vol")
