"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes, the oracle standing in for the kernels
(tests may use the oracle as the checker *and* as the stand-in local compute; the product never does)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_l2
from zernike3d_b200 import dist as zdist
from zernike3d_b200.synthetic import volume

ORDER, DIM = 8, 12


def test_pair_ranges_partition_the_volume():
    for dim in (12, 15, 128):
        half = (dim + 1) // 2
        for world in (1, 2, 3, 4, 8):
            rs = [zdist.pair_range(half, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == half and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            planes = sorted(p for lo, hi in rs for p in zdist.planes_of_pairs(dim, lo, hi))
            assert planes == list(range(dim))  # every plane exactly once, centre plane of odd boxes included
    with pytest.raises(ValueError):
        zdist.pair_range(8, 2, 2)
    assert [zdist.batch_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import sys
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle
    oracle.set_threads(1)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        half = (DIM + 1) // 2
        v = volume(5, DIM)

        def local_decompose(vol, pairs):
            lo, hi = pairs
            acc = np.zeros(oracle.num_moments(ORDER), np.complex128)
            for p in zdist.planes_of_pairs(DIM, lo, hi):
                acc += oracle.decompose(vol.numpy(), ORDER, oracle.EXACT, rows=(p, p + 1), return_f64=True)
            return torch.from_numpy(acc.astype(np.complex64))

        # planes this rank does not own are poisoned: the sharded path must not read them
        mine = zdist.planes_of_pairs(DIM, *zdist.pair_range(half, rank, world))
        local = np.full_like(v, np.nan); local[mine] = v[mine]
        for det in (False, True):
            mom = zdist.decompose_slab_sharded(local_decompose, torch.from_numpy(local), half, deterministic=det)
            np.save(os.path.join(out_dir, f"slab_{int(det)}_{rank}.npy"), mom.numpy())

        def local_part(vol, part):  # stand-in for Plan.decompose(vol, part=(i, n)): any disjoint split of the voxel sum
            i, n = part
            lo, hi = DIM * i // n, DIM * (i + 1) // n
            return torch.from_numpy(oracle.decompose(vol.numpy(), ORDER, oracle.EXACT, rows=(lo, hi)))

        mom = zdist.decompose_work_sharded(local_part, torch.from_numpy(v), deterministic=True)
        np.save(os.path.join(out_dir, f"work_{rank}.npy"), mom.numpy())

        def local_reconstruct(moments, pairs):
            out = torch.zeros(DIM, DIM, DIM)
            for p in zdist.planes_of_pairs(DIM, *pairs):
                out[p] = torch.from_numpy(oracle.reconstruct(moments.numpy(), DIM, ORDER, oracle.EXACT, rows=(p, p + 1))[0])
            return out

        full = torch.from_numpy(oracle.decompose(v, ORDER, oracle.EXACT))
        rec = zdist.reconstruct_slab_sharded(local_reconstruct, full, half, gather=True)
        np.save(os.path.join(out_dir, f"rec_{rank}.npy"), rec.numpy())

        vols = [volume(20 + i, DIM) for i in range(3)]  # 3 volumes over 2 ranks: unequal shares
        res, (lo, hi) = zdist.decompose_batch_sharded(
            lambda batch: torch.stack([torch.from_numpy(oracle.decompose(b.numpy(), ORDER, oracle.EXACT)) for b in batch]),
            torch.zeros(1), 3, lambda lo, hi: torch.from_numpy(np.stack(vols[lo:hi])), gather=True)
        np.save(os.path.join(out_dir, f"batch_{rank}.npy"), res.numpy())
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path, oracle_mod):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    v = volume(5, DIM)
    want = oracle_mod.decompose(v, ORDER, oracle_mod.EXACT)
    for det in (0, 1):
        got = [np.load(tmp_path / f"slab_{det}_{r}.npy") for r in range(world)]
        assert np.array_equal(got[0], got[1])           # every rank ends with the same full vector
        assert rel_l2(got[0], want) < 1e-6
    work = [np.load(tmp_path / f"work_{r}.npy") for r in range(world)]
    assert np.array_equal(work[0], work[1]) and rel_l2(work[0], want) < 1e-6
    rec_want = oracle_mod.reconstruct(want, DIM, ORDER, oracle_mod.EXACT)
    for r in range(world):
        assert rel_l2(np.load(tmp_path / f"rec_{r}.npy"), rec_want) < 1e-6
    bw = np.stack([oracle_mod.decompose(volume(20 + i, DIM), ORDER, oracle_mod.EXACT) for i in range(3)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"batch_{r}.npy"), bw)


def test_parse_cpulist_and_numa_binding_is_a_noop_without_topology(tmp_path):
    from zernike3d_b200 import dist as zdist

    assert zdist.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert zdist.parse_cpulist("") == set()
    assert zdist.bind_to_gpu_numa(0, sysfs=str(tmp_path)) == -1   # no CUDA device / no sysfs entry: nothing changes
