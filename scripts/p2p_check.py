"""Multi-GPU check (torchrun, one rank per GPU): fused reduce + peer-memory all-reduce vs NCCL / single-GPU results."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from zernike3d_b200 import dist as zdist
from zernike3d_b200.plan import Plan, PeerComm
from zernike3d_b200.synthetic import volume

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
order, dim = int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(sys.argv[2]) if len(sys.argv) > 2 else 64
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
P = Plan(order, dim, local)
vols = torch.from_numpy(np.stack([volume(3, dim), volume(4, dim)])).cuda()
full = P.decompose(vols)
comm = PeerComm(P, batch_max=2)
ok = True
for it in range(12):                      # back-to-back calls exercise both buffer slots / epochs
    v = vols if it % 3 == 0 else vols[:1]
    fused = P.decompose_allreduce(v, comm)
    want = zdist.decompose_work_sharded(P.decompose, v, deterministic=True)   # all-gather + rank-ordered sum
    nccl = zdist.decompose_work_sharded(P.decompose, v)                        # plain NCCL all-reduce
    torch.cuda.synchronize()
    ok &= bool(torch.equal(torch.view_as_real(fused), torch.view_as_real(want)))
    ok &= rel(fused.cpu().numpy(), full[: v.shape[0]].cpu().numpy()) < 5e-7
    ok &= rel(nccl.cpu().numpy(), full[: v.shape[0]].cpu().numpy()) < 5e-7
ok &= comm.error() == 0
# slab-sharded analysis and reconstruction through the plane-pair interface
mom = zdist.decompose_slab_sharded(P.decompose, vols[0], P.half)
ok &= rel(mom.cpu().numpy(), full[0].cpu().numpy()) < 5e-7
rec = zdist.reconstruct_slab_sharded(P.reconstruct, full[0], P.half, gather=True)
ok &= rel(rec.cpu().numpy(), P.reconstruct(full[0]).cpu().numpy()) < 1e-6
# batches through the tensor-core tiles: work-sharded (every rank a part of every volume's chunk list, one all-reduce) and
# volume-sharded (contiguous shares of independent volumes, gathered at the end)
nb = 16
bvols = torch.from_numpy(np.stack([volume(20 + i, dim) for i in range(nb)])).cuda()
bfull = P.decompose(bvols)
bws = zdist.decompose_work_sharded(P.decompose, bvols, deterministic=True)
ok &= rel(bws.cpu().numpy(), bfull.cpu().numpy()) < 1e-6
count = 2 * 14 * world
allv = [volume(100 + i, dim) for i in range(count)]
got, (lo, hi) = zdist.decompose_batch_sharded(P.decompose, bvols, count,
                                              lambda a, b: torch.from_numpy(np.stack(allv[a:b])).cuda())
ok &= tuple(got.shape) == (count, P.num_moments)
mine = P.decompose(torch.from_numpy(np.stack(allv[lo:hi])).cuda())
ok &= bool(torch.equal(torch.view_as_real(got[lo:hi]), torch.view_as_real(mine)))
other = (lo + count // 2) % count                                   # a volume another rank decomposed
ok &= rel(got[other].cpu().numpy(), P.decompose(torch.from_numpy(allv[other]).cuda()).cpu().numpy()) < 2e-6
# latency of one volume spread over the ranks: fused push exchange vs compute + NCCL all-reduce (device time, max over ranks)
def _ms(fn, reps=50):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
one = vols[:1].contiguous()
t_fused, t_nccl = _ms(lambda: P.decompose_allreduce(one, comm)), _ms(lambda: zdist.decompose_work_sharded(P.decompose, one))
t_part = _ms(lambda: P.decompose(one, part=(rank, world)))
if rank == 0:
    print(f"P2P_TIME world={world} order={order} dim={dim}: fused {t_fused:.4f} ms, compute + NCCL {t_nccl:.4f} ms, compute share alone {t_part:.4f} ms")
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("P2P_CHECK", "OK" if flag.item() == 1 else "FAILED", f"world={world} order={order} dim={dim}")
comm.close()
dist.destroy_process_group()
