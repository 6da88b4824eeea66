"""Multi-GPU sharding of the spharm path: one process per GPU, ``torch.distributed`` for the plumbing.

The reference has no multi-GPU code.  The path shards two ways:

* **voxel slabs of one volume** - the moment vector is a sum over voxels, so each rank sums the
  mirrored plane pairs it owns (``pair_range``) with the global normalisation and grid coordinates and the
  ~12k-element partial vectors are combined by ONE all-reduce (96 KB at order 50, latency bound).
  Reconstruction shards the same way with no collective at all: each rank writes its own planes.
* **batches of volumes** (3DVA groups, ensembles) - volumes are independent: contiguous chunks per
  rank, no data-path communication; results are gathered at the end if the caller wants them.

The functions take the local compute as a callable so the host logic (ranges, collectives, gather
order) is testable on CPU with the ``gloo`` backend; on GPUs the callable is ``Plan.decompose`` /
``Plan.reconstruct`` and the backend is NCCL over NVLink.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def pair_range(half: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced share of the mirrored plane pairs [0, half) for ``rank``."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return (half * rank // world, half * (rank + 1) // world)


def planes_of_pairs(dim: int, lo: int, hi: int) -> list[int]:
    """Array-axis-0 planes a rank holding pairs [lo, hi) reads / writes (centre plane once for odd dim)."""
    planes = list(range(lo, hi))
    planes += [dim - 1 - p for p in range(hi - 1, lo - 1, -1) if dim - 1 - p != p]
    return planes


def batch_range(count: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced share of ``count`` independent volumes for ``rank``."""
    return (count * rank // world, count * (rank + 1) // world)


def parse_cpulist(text: str) -> set[int]:
    """``"0-15,32-47"`` (the kernel's cpulist format) -> set of CPU numbers."""
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(device: int, sysfs: str = "/sys") -> int:
    """Pin this process to the CPUs of the NUMA node ``device`` hangs off, BEFORE it allocates pinned host buffers (first touch
    then places them on that node), so that a rank's H2D / D2H traffic does not cross the socket interconnect.  Returns the node,
    or -1 when the platform does not say (virtualised boxes report ``numa_node = -1``) - then nothing is changed."""
    import os

    try:
        p = torch.cuda.get_device_properties(device)
        pci = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"{sysfs}/bus/pci/devices/{pci}/numa_node").read())
        if node < 0:
            return -1
        cpus = parse_cpulist(open(f"{sysfs}/devices/system/node/node{node}/cpulist").read()) & os.sched_getaffinity(0)
        if not cpus:
            return -1
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return -1


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def allreduce_moments(partial: torch.Tensor, group=None, deterministic: bool = False) -> torch.Tensor:
    """Sum complex64 partial moment vectors ``[..., num_moments]`` over the group, in place.

    ``deterministic=True`` all-gathers the vectors and adds them in rank order, so the result does not
    depend on the collective algorithm NCCL picks (bit-reproducible for a fixed world size)."""
    rank, world = _world(group)
    if world == 1:
        return partial
    flat = torch.view_as_real(partial)  # NCCL / gloo reduce float32, not complex
    if deterministic:
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat.contiguous(), group=group)
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        flat.copy_(acc)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return partial


def decompose_slab_sharded(local_decompose: Callable[..., torch.Tensor], vol: torch.Tensor, half: int,
                           group=None, deterministic: bool = False) -> torch.Tensor:
    """One (batch of) large volume(s), voxel-sharded: every rank ends with the full moments.

    ``local_decompose(vol, pairs=(lo, hi))`` must return this rank's partial, already-scaled moments
    (``Plan.decompose``); only the planes of the rank's pairs need to be valid in ``vol``."""
    rank, world = _world(group)
    lo, hi = pair_range(half, rank, world)
    partial = local_decompose(vol, pairs=(lo, hi))
    return allreduce_moments(partial, group=group, deterministic=deterministic)


def decompose_work_sharded(local_decompose: Callable[..., torch.Tensor], vol: torch.Tensor, group=None,
                           deterministic: bool = False) -> torch.Tensor:
    """One (batch of) volume(s) replicated on every rank, sharded by WORK: rank r sums the r-th of ``world`` equal
    ranges of the kernel's work list (``Plan.decompose(vol, part=(r, world))``), then one all-reduce.  Unlike plane
    slabs this keeps the transpose pairing (16-fold mirror folding) intact on every rank, so the ranks stay balanced."""
    rank, world = _world(group)
    partial = local_decompose(vol, part=(rank, world))
    return allreduce_moments(partial, group=group, deterministic=deterministic)


def reconstruct_slab_sharded(local_reconstruct: Callable[..., torch.Tensor], moments: torch.Tensor, half: int,
                             group=None, gather: bool = False) -> torch.Tensor:
    """Each rank synthesises only its planes (no collective).  ``gather=True`` additionally sums the
    disjoint slabs so every rank holds the whole map (one all-reduce of the volume)."""
    rank, world = _world(group)
    lo, hi = pair_range(half, rank, world)
    vol = local_reconstruct(moments, pairs=(lo, hi))
    if gather and world > 1:
        dist.all_reduce(vol, op=dist.ReduceOp.SUM, group=group)  # slabs are disjoint, the rest is zero
    return vol


def decompose_batch_sharded(local_decompose: Callable[..., torch.Tensor], vols: Sequence, count: int,
                            load: Callable[[int, int], torch.Tensor], group=None, gather: bool = True):
    """Independent volumes: rank r decomposes volumes ``batch_range(count, r, world)`` obtained from
    ``load(lo, hi)``; no data-path collective.  With ``gather=True`` the per-rank results are
    all-gathered into ``[count, num_moments]`` on every rank (ranks may hold unequal counts)."""
    rank, world = _world(group)
    lo, hi = batch_range(count, rank, world)
    local = local_decompose(load(lo, hi)) if hi > lo else None
    if not gather or world == 1:
        return local, (lo, hi)
    sizes = [batch_range(count, r, world) for r in range(world)]
    width = max(h - l for l, h in sizes)
    ref = local if local is not None else vols  # device / dtype template
    nm = local.shape[-1] if local is not None else None
    nm_t = torch.tensor([nm or 0], device=ref.device if isinstance(ref, torch.Tensor) else "cpu")
    dist.all_reduce(nm_t, op=dist.ReduceOp.MAX, group=group)
    nm = int(nm_t.item())
    dev = ref.device if isinstance(ref, torch.Tensor) else "cpu"
    pad = torch.zeros((width, nm, 2), dtype=torch.float32, device=dev)
    if local is not None:
        pad[: hi - lo] = torch.view_as_real(local)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.cat([p[: h - l] for p, (l, h) in zip(parts, sizes)], dim=0)
    return torch.view_as_complex(out.contiguous()), (lo, hi)
