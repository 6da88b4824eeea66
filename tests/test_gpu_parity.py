"""Parity of the CUDA path (through the C ABI) with the reference goldens and the CPU oracle.

Tolerances are the north-star's, norm-wise (relative L2 over the whole vector):
  moments <= 1e-5, reconstructed map <= 1e-5, invariant vector <= 1e-6.
The reference itself is 2e-6 (64^3) .. 5e-6 (128^3) away from an exact evaluation because it sums float32
products in BLAS (zm:2210); against the float64 oracle the kernels are held to a tighter 1.5e-6.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from zernike3d_b200.synthetic import bundle_3dva, volume

pytestmark = pytest.mark.gpu

TOL_MOM, TOL_MAP, TOL_INV = 1e-5, 1e-5, 1e-6
TOL_EXACT = 1.5e-6
# batches take the tensor-core path (3xTF32 products, TMEM accumulators): T = R*Q is a float32 recurrence before the
# contraction, which costs a little against the float64 oracle (1.0e-6 at order 50, 1.4e-6 at order 60)
TOL_EXACT_BATCHED = 2.0e-6


@pytest.fixture(scope="module")
def plans():
    from zernike3d_b200.plan import Plan

    cache = {}

    def get(order, dim):
        if (order, dim) not in cache:
            cache[(order, dim)] = Plan(order, dim)
        return cache[(order, dim)]

    yield get
    cache.clear()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", ["d16_o10", "d15_o8", "d32_o20", "d20_o25", "c1_d64_o50", "c3_d128_o50", "c5_d96_o60"])
def test_decompose_vs_reference_golden(plans, oracle_mod, name):
    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    P = plans(order, dim)
    for si, seed in enumerate(g["seeds"]):
        v = volume(int(seed), dim)
        mom = P.decompose(dev(v)).cpu().numpy()
        assert mom.dtype == np.complex64 and mom.shape == (P.num_moments,)
        assert rel_l2(mom, g["moments"][si]) < TOL_MOM
        inv = P.invariants(dev(mom)).cpu().numpy()
        # the 1e-6 bar on invariants holds against the exact oracle everywhere and against the reference's own float32-BLAS
        # numbers up to 64^3.  At 96^3 / 128^3 the reference does not reproduce ITSELF to 1e-6: tests/golden/ref_selfnoise.npz
        # (oracle/make_selfnoise.py) holds the difference between two runs of the unmodified reference that only differ in the
        # order of its float32 summation (zm:2210) - invariants 1.5e-6 / 3.4e-6, moments 5.2e-6 / 7.6e-6.  The bar against the
        # reference there is that measured self-noise (x 1.5), not a number picked to make the test pass.
        assert rel_l2(inv, oracle_mod.invariants(mom, order)) < 2e-7
        tol_inv_ref = TOL_INV
        if dim > 64:
            noise = golden("selfnoise")
            tol_inv_ref = max(TOL_INV, 1.5 * float(noise[f"{name}/invariants_fwd_vs_rev"]))
            assert float(noise[f"{name}/invariants_fwd_vs_exact"]) > 0.9 * TOL_INV      # the reference itself is >= 1e-6 from exact
        assert rel_l2(inv, g["invariants"][si]) < tol_inv_ref
        if dim <= 64:
            exact = oracle_mod.decompose(v, order, oracle_mod.EXACT)
            assert rel_l2(mom, exact) < TOL_EXACT
            assert rel_l2(inv, oracle_mod.invariants(exact, order)) < TOL_INV


@pytest.mark.parametrize("name", ["d16_o10", "d15_o8", "d32_o20", "d20_o25", "c1_d64_o50"])
def test_reconstruct_vs_reference_golden(plans, name):
    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    P = plans(order, dim)
    for si in range(len(g["seeds"])):
        rec = P.reconstruct(dev(g["moments"][si])).cpu().numpy()
        assert rec.dtype == np.float32 and rec.shape == (dim, dim, dim)
        assert np.isfinite(rec).all()
        assert rel_l2(rec, g["recon"][si]) < TOL_MAP


def test_128_order50_against_exact_oracle_rows(plans, oracle_mod):
    """Headline configuration against the float64 oracle on a bounded sample: slab partials of 4 plane pairs."""
    order, dim = 50, 128
    P = plans(order, dim)
    v = volume(1, dim)
    dv = dev(v)
    exact = oracle_mod.decompose(v, order, oracle_mod.EXACT)     # ~2 s on the GPU box's host cores
    full = P.decompose(dv).cpu().numpy()
    assert rel_l2(full, exact) < TOL_EXACT
    assert rel_l2(P.invariants(dev(full)).cpu().numpy(), oracle_mod.invariants(exact, order)) < TOL_INV
    for lo, hi in [(0, 1), (37, 38), (63, 64)]:
        got = P.decompose(dv, pairs=(lo, hi)).cpu().numpy()
        want = (oracle_mod.decompose(v, order, oracle_mod.EXACT, rows=(lo, hi), return_f64=True)
                + oracle_mod.decompose(v, order, oracle_mod.EXACT, rows=(dim - hi, dim - lo), return_f64=True))
        assert rel_l2(got, want) < TOL_EXACT
    mom = golden("c3_d128_o50")["moments"][0]
    rec = P.reconstruct(dev(mom)).cpu().numpy()
    scale = np.sqrt(np.mean(rec.astype(np.float64) ** 2))        # whole-map RMS: the bar is norm-wise over the map
    for p in [0, 40, 64, 127]:
        want = oracle_mod.reconstruct(mom, dim, order, oracle_mod.EXACT, rows=(p, p + 1))[0]
        assert rel_l2(rec[p], want) < TOL_MAP                    # even per plane (edge planes are ~100x fainter)
        assert np.sqrt(np.mean((rec[p] - want) ** 2)) / scale < TOL_EXACT


def test_3dva_bundle_batched(plans):
    g = golden("3dva_d24_o12")
    order, dim = int(g["order"]), int(g["dim"])
    _, cons, comps = bundle_3dva(dim, components=5, particles=64, seed=2)
    vols = np.stack([cons, *comps])
    P = plans(order, dim)
    moms = P.decompose(dev(vols)).cpu().numpy()       # one batched call for base map + 5 components
    assert moms.shape == (6, P.num_moments)
    for i in range(6):
        assert rel_l2(moms[i], g["moments"][i]) < TOL_MOM
    one_by_one = np.stack([P.decompose(dev(v)).cpu().numpy() for v in vols])
    assert np.array_equal(moms, one_by_one)           # batching does not change a single bit


BATCHED = [("d16_o10", 13), ("d15_o8", 12), ("d32_o20", 35), ("d20_o25", 45), ("d16_o10", 139), ("c1_d64_o50", 32),
           ("c1_d64_o50", 50), ("c5_d96_o60", 16)]


@pytest.mark.parametrize("name,count", BATCHED)
def test_batched_tensor_core_path(plans, oracle_mod, name, count):
    """Batches of >= info['gen_min'] volumes take the TABLE-FREE tensor-core kernel: k_fold_tile_gen + k_decompose_gen (tcgen05,
    T = R Q generated in the kernel by the recurrences of zm:1991-2000 / zm:2021-2026, nothing materialised) + k_reduce_stream."""
    from zernike3d_b200.plan import launch_count

    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    P = plans(order, dim)
    info = P.info()
    assert info["gen"] == 1 and info["tables_level"] == 0 and 0 < info["gen_min"] <= count
    seed0 = int(g["seeds"][0])
    vols = np.stack([volume(seed0 + i, dim) for i in range(count)])
    dv = dev(vols)
    P.decompose(dv)                                           # (allocates the plan's torch-owned scratch)
    free0 = torch.cuda.mem_get_info()[0]
    l0 = launch_count()
    mb = P.decompose(dv)
    assert P.info()["stream_state"] != 2 and P.info()["tables_level"] == 0       # no basis table appeared behind the caller's back
    assert free0 - torch.cuda.mem_get_info()[0] < (8 << 20)
    tiles, left, covered = 0, count, 0                        # the C side's tiling: 64-volume tiles while >= gen_min volumes are
    while left >= info["gen_min"]:                            # left; the rest goes through the per-volume kernel
        take = min(64, left)
        tiles, left, covered = tiles + 1, left - take, covered + take
    tail = left
    assert launch_count() - l0 >= 3 * tiles + (2 if tail else 0)  # fold + contraction + reduction per tile, then the tail
    assert mb.shape == (count, P.num_moments) and mb.dtype == torch.complex64
    assert torch.equal(torch.view_as_real(mb), torch.view_as_real(P.decompose(dv)))     # run-to-run bitwise
    mb = mb.cpu().numpy()
    assert np.isfinite(mb.view(np.float32)).all()
    # volume 0 is the golden's volume: the reference's own numbers
    assert rel_l2(mb[0], g["moments"][0]) < TOL_MOM
    # every volume against the per-volume kernel (different summation order and arithmetic, same tables of coefficients)
    for i in range(count):
        one = P.decompose(dv[i]).cpu().numpy()
        if i >= covered:                                      # tail volumes took the per-volume kernel inside the batch
            assert np.array_equal(mb[i], one)
        else:
            assert rel_l2(mb[i], one) < 2 * TOL_EXACT_BATCHED
    # first and last volume of the first tile against the float64 oracle, moments and invariants
    for i in ([0, covered - 1] if dim <= 64 else []):
        exact = oracle_mod.decompose(vols[i], order, oracle_mod.EXACT)
        assert rel_l2(mb[i], exact) < TOL_EXACT_BATCHED
        inv = P.invariants(dev(mb[i])).cpu().numpy()
        assert rel_l2(inv, oracle_mod.invariants(exact, order)) < TOL_INV


@pytest.mark.parametrize("name,count", [("d32_o20", 35), ("c1_d64_o50", 70)])
def test_opt_in_table_fed_kernels(oracle_mod, name, count):
    """z3d_plan_prepare_tables: level 1 = R / Q factor tables (k_decompose_batched, 32- and 64-volume tiles), level 2 = + the
    streamed product table (k_decompose_stream).  Same results as the table-free default to the batched tolerance; a T table
    over the byte cap leaves the plan at level 1; level 0 frees everything and returns to the default kernel bit for bit."""
    from zernike3d_b200.plan import Plan

    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    vols = np.stack([volume(int(g["seeds"][0]) + i, dim) for i in range(count)])
    dv = dev(vols)
    P = Plan(order, dim)
    m_gen = P.decompose(dv).cpu().numpy()
    assert P.prepare_tables(2, max_bytes=1) == 1 and P.info()["stream_state"] != 2   # capped: factor tables only
    m_rq = P.decompose(dv).cpu().numpy()
    assert P.prepare_tables(2) == 2 and P.info()["stream_state"] == 2
    m_t = P.decompose(dv).cpu().numpy()
    for m in (m_gen, m_rq, m_t):
        assert rel_l2(m[0], g["moments"][0]) < TOL_MOM
        for i in (0, count - 1):
            assert rel_l2(m[i], oracle_mod.decompose(vols[i], order, oracle_mod.EXACT)) < TOL_EXACT_BATCHED
    assert rel_l2(m_rq, m_t) < 2 * TOL_EXACT_BATCHED and rel_l2(m_gen, m_t) < 2 * TOL_EXACT_BATCHED
    assert not np.array_equal(m_rq, m_gen) and not np.array_equal(m_t, m_gen)        # three different kernels really ran
    assert P.prepare_tables(1) == 1
    assert np.array_equal(P.decompose(dv).cpu().numpy(), m_rq)
    assert P.prepare_tables(0) == 0 and P.info()["tables_level"] == 0
    assert np.array_equal(P.decompose(dv).cpu().numpy(), m_gen)


def test_batched_128_order50_headline(plans, oracle_mod):
    """The bench configuration: 32 volumes of 128^3 at order 50 in one tile."""
    order, dim = 50, 128
    P = plans(order, dim)
    g = golden("c3_d128_o50")
    vols = np.stack([volume(int(g["seeds"][0]) + i, dim) for i in range(32)])
    dv = dev(vols)
    mb = P.decompose(dv).cpu().numpy()
    assert rel_l2(mb[0], g["moments"][0]) < TOL_MOM
    for i in (0, 31):
        exact = oracle_mod.decompose(vols[i], order, oracle_mod.EXACT)         # ~2 s each on the host cores
        assert rel_l2(mb[i], exact) < TOL_EXACT_BATCHED
        inv = P.invariants(dev(mb[i])).cpu().numpy()
        assert rel_l2(inv, oracle_mod.invariants(exact, order)) < TOL_INV
    # the same volumes in one 64-volume tile (N = 128 MMAs, 4-tile row groups, 3 rounds)
    wide = P.decompose(torch.cat([dv, dv])).cpu().numpy()
    assert rel_l2(wide[:32], mb) < 1.5e-6 and np.array_equal(wide[:32], wide[32:])
    for i in (0, 31):
        assert rel_l2(wide[i], oracle_mod.decompose(vols[i], order, oracle_mod.EXACT)) < TOL_EXACT_BATCHED
    # work parts of a batch are additive (multi-GPU work sharding of replicated volumes), linearity in the volumes
    parts = sum(P.decompose(dv, part=(r, 3)) for r in range(3)).cpu().numpy()
    assert rel_l2(parts, mb) < 1e-6
    mix = (0.5 * dv[:16] - 2.0 * dv[16:]).contiguous()
    lin = P.decompose(torch.cat([mix, mix])).cpu().numpy()[:16]
    assert rel_l2(lin, 0.5 * mb[:16] - 2.0 * mb[16:]) < 2e-6


def test_latent_scaling_on_device_matches_apply_scaling(plans, tmp_path):
    """z3d_scale_moments (+ ZernikeSpherical.ReconstructLatents) against Moments.apply_scaling (zm:78-95)."""
    import argparse

    from zernike3d_b200.moments import Moments
    from zernike3d_b200.spherical import ZernikeSpherical

    g = golden("3dva_d24_o12")
    order, dim = int(g["order"]), int(g["dim"])
    P = plans(order, dim)
    dims, lat = np.asarray(g["moments"], np.complex64), np.asarray(g["latents"], np.float32)
    got = P.scale_moments(dev(dims), dev(lat[:40])).cpu().numpy()
    want = dims[0][None] + lat[:40] @ dims[1:]
    assert got.shape == (40, P.num_moments) and rel_l2(got, want) < 1e-6
    M = Moments(quiet=True)
    for v in dims:
        M.from_array(v)
        M.add_dimension()
    M.latents = lat
    M.stats(quiet=True)
    for z in (0, 7, 39):
        assert rel_l2(got[z], M.apply_scaling(z_ind=z)) < 1e-6
    with pytest.raises(ValueError):
        P.scale_moments(dev(dims[:, :5]), dev(lat))
    # maps at several latent coordinates, scaling and synthesis on the device
    args = argparse.Namespace(order=order, output=str(tmp_path / "traj.mrc"), apix=None, gpu_id=0, z_ind=None, scales=None)
    Z = ZernikeSpherical(args)
    Z.dim, Z.moments = dim, M
    jobs = Z.CalculatePolynomials()
    maps = Z.ReconstructLatents(z_indices=[0, 7, 39], batch=2, write=False)
    for i, z in enumerate((0, 7, 39)):
        one = P.reconstruct(dev(M.apply_scaling(z_ind=z))).cpu().numpy()
        assert rel_l2(maps[i], one) < 1e-6
    # a trajectory of 40 latent coordinates in one launch: routed to the tensor-core synthesis kernel (>= synth_min maps)
    traj = Z.ReconstructLatents(z_indices=list(range(40)), write=False)
    assert traj.shape[0] == 40 and 40 >= P.info()["synth_min"] and P.info()["synth"] == 1
    for z in (0, 7, 39):
        one = P.reconstruct(dev(M.apply_scaling(z_ind=z)), batched=False).cpu().numpy()
        assert rel_l2(traj[z], one) < TOL_EXACT
    names = Z.ReconstructLatents(latents=lat[:2])
    assert [n.rsplit("_", 1)[1] for n in names] == ["p0.mrc", "p1.mrc"] and all(__import__("os").path.exists(n) for n in names)
    assert len(jobs) == P.num_moments


def test_zcc_on_device_matches_the_host_curve(plans):
    """z3d_zcc (Plan.zcc) against Moments.zcc, the restatement of Z_3DVA.calc_ZCC (moment_analysis_3dva.py:402-422)."""
    from zernike3d_b200.moments import Moments

    order, dim = 20, 32
    P = plans(order, dim)
    vols = np.stack([volume(40 + i, dim) for i in range(4)])
    mom = P.decompose(dev(vols))
    curves = P.zcc(mom[:2], mom[2:]).cpu().numpy()            # pairs (0, 2) and (1, 3)
    assert curves.shape == (2, order + 1)
    M = Moments()
    for i in range(4):
        M.from_array(mom[i].cpu().numpy(), order)
        M.add_dimension()
    for p, (i, j) in enumerate(((0, 2), (1, 3))):
        assert np.allclose(curves[p], M.zcc(M, i, j), rtol=2e-6, atol=1e-7)
    same = P.zcc(mom[0], mom[0]).cpu().numpy()
    assert same.shape == (order + 1,) and np.allclose(same, 1.0, atol=1e-6)


def test_zcc_and_latent_scaling_on_device_against_reference_run_fixture(plans):
    """z3d_zcc and z3d_scale_moments against tests/golden/ref_zcc_scaling.npz: curves from the reference's own calc_ZCC statements
    (moment_analysis_3dva.py:402-421) and moments scaled by the reference's own Moments.apply_scaling (zm:78-95)."""
    g = golden("zcc_scaling")
    P = plans(int(g["order"]), 64)
    sets = dev(np.asarray(g["moment_sets"], np.complex64))
    ia = [int(i) for i, _ in g["pairs"]]
    ib = [int(j) for _, j in g["pairs"]]
    curves = P.zcc(sets[ia].contiguous(), sets[ib].contiguous()).cpu().numpy()
    assert curves.shape == (len(ia), int(g["order"]) + 1)
    assert np.allclose(curves, g["zcc"], rtol=3e-6, atol=1e-7)
    b = golden("3dva_d24_o12")
    Pb = plans(int(b["order"]), int(b["dim"]))
    zs = [int(z) for z in g["scaling_z"]]
    lat = np.asarray(b["latents"], np.float32)[zs]
    got = Pb.scale_moments(dev(np.asarray(b["moments"], np.complex64)), dev(lat)).cpu().numpy()
    for k in range(len(zs)):
        assert rel_l2(got[k], g["scaled"][k]) < 1e-6


@pytest.mark.parametrize("n,crop,bin_,center", [(32, None, 2, None), (32, 20, None, (14, 16, 17)), (40, 24, 2, (20, 19, 21)),
                                                  (33, None, 3, None), (64, 48, 1.5, (30, 33, 32)), (31, 17, 2.5, (15, 15, 16))])
def test_crop_bin_on_device_is_bit_identical_to_the_host_path(n, crop, bin_, center):
    """z3d_crop_bin (plan.crop_bin) against ZernikeSpherical.bin_crop_array_center / scipy.ndimage.zoom(order=0), the host path
    that mirrors zm:2781-2815 and zm:2867-2876 (index work: bit-exact, including scipy's out-of-range zero at e.g. 32 -> 16)."""
    import scipy.ndimage

    from zernike3d_b200.plan import crop_bin
    from zernike3d_b200.spherical import ZernikeSpherical

    rng = np.random.default_rng(n)
    vols = (rng.normal(size=(3, n, n, n)) + 2.0).astype(np.float32)
    got = crop_bin(dev(vols), crop_to=crop, bin=bin_, center=center).cpu().numpy()
    if crop is not None:
        want = np.stack(ZernikeSpherical.bin_crop_array_center(list(vols), crop, bin=bin_, center=center))
    else:
        want = np.stack([scipy.ndimage.zoom(v, 1 / bin_, order=0) for v in vols])
    assert got.shape == want.shape and np.array_equal(got, want)


def test_slab_partials_are_additive_and_only_read_their_planes(plans):
    order, dim = 20, 32
    P = plans(order, dim)
    v = volume(9, dim)
    full = P.decompose(dev(v)).cpu().numpy()
    h = P.half
    acc = np.zeros_like(full)
    for lo, hi in [(0, 3), (3, 11), (11, h)]:
        poisoned = np.full_like(v, np.nan)
        planes = list(range(lo, hi)) + [dim - 1 - p for p in range(lo, hi)]
        poisoned[planes] = v[planes]
        part = P.decompose(dev(poisoned), pairs=(lo, hi)).cpu().numpy()
        assert np.isfinite(part).all()
        acc += part
    assert rel_l2(acc, full) < 5e-7
    # reconstruction writes only its planes
    out = torch.full((1, dim, dim, dim), 7.0, device="cuda")
    P.reconstruct(dev(full).unsqueeze(0), pairs=(2, 5), out=out)
    out = out[0].cpu().numpy()
    ref = P.reconstruct(dev(full)).cpu().numpy()
    touched = [2, 3, 4, dim - 3, dim - 4, dim - 5]
    for p in range(dim):
        assert np.array_equal(out[p], ref[p]) if p in touched else np.all(out[p] == 7.0)


def test_work_parts_are_additive(plans):
    """z3d_decompose_part: n equal ranges of the work list (multi-GPU sharding of one replicated volume) sum to the whole."""
    order, dim = 50, 64
    P = plans(order, dim)
    dv = dev(volume(12, dim))
    full = P.decompose(dv).cpu().numpy()
    for n in (2, 3, 8):
        acc = sum(P.decompose(dv, part=(i, n)).cpu().numpy() for i in range(n))
        assert rel_l2(acc, full) < 5e-7
    with pytest.raises(Exception):
        P.decompose(dv, part=(3, 3))
    with pytest.raises(ValueError):
        P.decompose(dv, part=(0, 2), pairs=(0, 4))


def test_deterministic_and_linear(plans):
    order, dim = 50, 64
    P = plans(order, dim)
    a, b = dev(volume(1, dim)), dev(volume(2, dim))
    m1 = P.decompose(a)
    m2 = P.decompose(a)
    assert torch.equal(torch.view_as_real(m1), torch.view_as_real(m2))     # fixed-order partial combination
    mab = P.decompose(2.0 * a - 0.5 * b).cpu().numpy()
    lin = (2.0 * m1 - 0.5 * P.decompose(b)).cpu().numpy()
    assert rel_l2(mab, lin) < 2e-6


def test_invariants_under_axis_permutation_and_flips(plans):
    """F_nl is rotation invariant; the 90-degree rotations / mirror flips that map the symmetric grid onto itself
    are exact symmetries of the discrete sum, so the invariants agree to float32 round-off."""
    order, dim = 30, 48
    P = plans(order, dim)
    v = volume(4, dim)
    base = P.invariants(P.decompose(dev(v))).cpu().numpy()
    for w in (np.rot90(v, 1, axes=(0, 1)), np.rot90(v, 1, axes=(1, 2)), v[::-1, ::-1, :], np.transpose(v, (2, 0, 1))):
        inv = P.invariants(P.decompose(dev(np.ascontiguousarray(w)))).cpu().numpy()
        assert rel_l2(inv, base) < TOL_INV


def test_round_trip_full_size_properties(plans):
    """BASELINE sizes without a CPU oracle run: decompose -> reconstruct gain ~ 1/(4 pi) ((N-1)/N)^3 (zm:2198)."""
    for order, dim in [(50, 128), (60, 96)]:
        P = plans(order, dim)
        v = volume(3, dim)
        rec = P.reconstruct(P.decompose(dev(v))).cpu().numpy().astype(np.float64)
        gain = float(np.vdot(rec.ravel(), v.ravel()) / np.vdot(v.ravel(), v.ravel()))
        want = (1 / (4 * np.pi)) * ((dim - 1) / dim) ** 3
        assert abs(gain / want - 1) < 0.1
        smooth = rec / gain
        cc = np.corrcoef(smooth.ravel(), v.ravel())[0, 1]
        assert cc > 0.9


def test_256_stress_case(plans, oracle_mod):
    """C5b: 256^3 order 50 (the reference would need a 223 GB basis).  Plane-pair partials and reconstructed planes against the
    float64 oracle (bounded CPU work: `rows=`), one whole volume against the oracle, a batch of 8 through the table-free
    tensor-core kernel (the same kernel as at 128^3: no table, so no size cliff), slab additivity, real m = 0 moments, determinism."""
    order, dim = 50, 256
    P = plans(order, dim)
    v = volume(8, dim)
    dv = dev(v)
    full = P.decompose(dv)
    parts = P.decompose(dv, pairs=(0, 50)) + P.decompose(dv, pairs=(50, 128))
    assert rel_l2(parts.cpu().numpy(), full.cpu().numpy()) < 1e-6   # different float32 summation grouping only
    assert torch.equal(torch.view_as_real(full), torch.view_as_real(P.decompose(dv)))
    from zernike3d_b200.indexing import nlm_table

    m0 = nlm_table(order)[:, 2] == 0
    f = full.cpu().numpy()
    assert np.all(f.imag[m0] == 0) and np.isfinite(f.view(np.float32)).all()
    # plane pairs (edge, off-centre, centre) against oracle rows: zm:2194-2217 restricted to the planes of the pair
    for lo in (0, 77, 127):
        got = P.decompose(dv, pairs=(lo, lo + 1)).cpu().numpy()
        want = (oracle_mod.decompose(v, order, oracle_mod.EXACT, rows=(lo, lo + 1), return_f64=True)
                + oracle_mod.decompose(v, order, oracle_mod.EXACT, rows=(dim - 1 - lo, dim - lo), return_f64=True))
        assert rel_l2(got, want) < TOL_EXACT
    exact = oracle_mod.decompose(v, order, oracle_mod.EXACT)          # the whole volume once (~15 s of host threads)
    assert rel_l2(f, exact) < TOL_EXACT
    assert rel_l2(P.invariants(full).cpu().numpy(), oracle_mod.invariants(exact, order)) < TOL_INV
    # reconstructed planes against oracle rows (zm:2596-2614)
    rec = P.reconstruct(full).cpu().numpy()
    scale = np.sqrt(np.mean(rec.astype(np.float64) ** 2))
    for p_ in (1, 100, 128, 255):
        want = oracle_mod.reconstruct(f, dim, order, oracle_mod.EXACT, rows=(p_, p_ + 1))[0]
        assert np.sqrt(np.mean((rec[p_] - want) ** 2)) / scale < TOL_EXACT
    del rec
    # a batch of 8 (>= gen_min): the tensor-core kernel at 256^3, volume 0 = v
    batch = torch.stack([dv] + [dev(volume(80 + i, dim)) for i in range(7)])
    info = P.info()
    assert info["gen"] == 1 and batch.shape[0] >= info["gen_min"]
    mb = P.decompose(batch).cpu().numpy()
    assert rel_l2(mb[0], exact) < TOL_EXACT_BATCHED
    assert rel_l2(P.invariants(dev(mb[0])).cpu().numpy(), oracle_mod.invariants(exact, order)) < TOL_INV
    for i in (3, 7):
        assert rel_l2(mb[i], P.decompose(batch[i]).cpu().numpy()) < 2 * TOL_EXACT_BATCHED


def test_3dva_bundle_at_size(plans, oracle_mod):
    """C4 at the BASELINE size: base map + 5 components at 128^3, order 50, ONE call (zm:2886-2889).  The base map is the golden
    volume of C3, so its moments are checked against the reference's own run; every member against the float64 oracle."""
    order, dim = 50, 128
    P = plans(order, dim)
    g = golden("c3_d128_o50")
    base = volume(int(g["seeds"][0]), dim)
    comps = []
    for i in range(5):
        c = 0.1 * volume(300 + i, dim)
        comps.append((c - c.mean()).astype(np.float32))
    vols = np.stack([base, *comps])
    moms = P.decompose(dev(vols)).cpu().numpy()
    assert rel_l2(moms[0], g["moments"][0]) < TOL_MOM
    for i in range(6):
        exact = oracle_mod.decompose(vols[i], order, oracle_mod.EXACT)
        assert rel_l2(moms[i], exact) < TOL_EXACT
        assert rel_l2(P.invariants(dev(moms[i])).cpu().numpy(), oracle_mod.invariants(exact, order)) < TOL_INV
    assert np.array_equal(moms, np.stack([P.decompose(dev(x)).cpu().numpy() for x in vols]))   # bundle == one by one, bit for bit


def test_reconstruction_at_full_size_against_oracle_planes(plans, oracle_mod):
    """Synthesis at the BASELINE sizes (96^3 / order 60, 128^3 / order 50) against float64 oracle planes of zm:2596-2614 spread over
    the box, instead of a property check: per plane and against the whole-map RMS."""
    for name, planes in (("c5_d96_o60", (0, 17, 47, 48, 70, 95)), ("c3_d128_o50", (0, 9, 63, 64, 101, 127))):
        g = golden(name)
        order, dim = int(g["order"]), int(g["dim"])
        P = plans(order, dim)
        mom = np.asarray(g["moments"][0], np.complex64)
        rec = P.reconstruct(dev(mom)).cpu().numpy()
        scale = np.sqrt(np.mean(rec.astype(np.float64) ** 2))
        for p_ in planes:
            want = oracle_mod.reconstruct(mom, dim, order, oracle_mod.EXACT, rows=(p_, p_ + 1))[0]
            assert np.sqrt(np.mean((rec[p_] - want) ** 2)) / scale < TOL_EXACT      # the bar is norm-wise over the map
            if np.sqrt(np.mean(want.astype(np.float64) ** 2)) > 0.05 * scale:       # and per plane wherever the plane is not
                assert rel_l2(rec[p_], want) < TOL_MAP                              # ~100x fainter than the map (edge planes)
        # and the round trip: decompose(reconstruct(mom)) is a fixed linear map of mom; the oracle's planes above pin the synthesis,
        # here the analysis of that map must reproduce the oracle's analysis of the same map on a plane pair
        lo = dim // 3
        got = P.decompose(dev(rec), pairs=(lo, lo + 1)).cpu().numpy()
        want = (oracle_mod.decompose(rec, order, oracle_mod.EXACT, rows=(lo, lo + 1), return_f64=True)
                + oracle_mod.decompose(rec, order, oracle_mod.EXACT, rows=(dim - 1 - lo, dim - lo), return_f64=True))
        assert rel_l2(got, want) < TOL_EXACT


@pytest.mark.parametrize("name,count", [("d16_o10", 13), ("d15_o8", 70), ("d32_o20", 16), ("d20_o25", 12), ("c1_d64_o50", 14)])
def test_batched_synthesis_vs_reference_golden(plans, name, count):
    """Tensor-core batched synthesis (z3d_reconstruct_batched: tiles of 64 moment sets, zm:2642-2715 for an ensemble): the golden
    moment sets of the reference, scaled copies and sums of them as one batch; every map against the reference's own reconstruction
    (linearity of zm:2596-2614) and against the per-map kernel."""
    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    P = plans(order, dim)
    assert P.info()["synth"] == 1
    base = np.stack([np.asarray(m, np.complex64) for m in g["moments"]])
    rng = np.random.default_rng(5)
    w = rng.normal(size=(count, base.shape[0])).astype(np.float32)
    w[: base.shape[0]] = np.eye(base.shape[0], dtype=np.float32)
    moms = (w.astype(np.complex64) @ base).astype(np.complex64)
    rec = P.reconstruct(dev(moms), batched=True).cpu().numpy()
    assert rec.shape == (count, dim, dim, dim) and np.isfinite(rec).all()
    want = np.tensordot(w, np.stack(g["recon"]).astype(np.float64), axes=1)
    for i in range(count):
        assert rel_l2(rec[i], want[i]) < TOL_MAP
    per_map = P.reconstruct(dev(moms[:3]), batched=False).cpu().numpy()
    for i in range(3):
        assert rel_l2(rec[i], per_map[i]) < TOL_EXACT
    again = P.reconstruct(dev(moms), batched=True).cpu().numpy()
    assert np.array_equal(rec, again)                                   # deterministic
    auto = P.reconstruct(dev(moms)).cpu().numpy()                       # count >= synth_min: routed to the tensor cores by default
    assert np.array_equal(auto, rec) == (count >= P.info()["synth_min"])


def test_batched_synthesis_at_full_size_against_oracle_planes(plans, oracle_mod):
    """128^3 / order 50, 70 moment sets (a full tile + a ragged one) through the tensor-core synthesis kernel: planes of maps of
    both tiles against the float64 oracle's zm:2596-2614, and the 96^3 / order 60 layout (143 column sets)."""
    for name, count, planes in (("c3_d128_o50", 70, (0, 40, 64, 127)), ("c5_d96_o60", 20, (1, 47, 80))):
        g = golden(name)
        order, dim = int(g["order"]), int(g["dim"])
        P = plans(order, dim)
        m0 = np.asarray(g["moments"][0], np.complex64)
        rng = np.random.default_rng(11)
        moms = np.stack([m0 * np.float32(s) for s in rng.uniform(0.5, 1.5, size=count)]).astype(np.complex64)
        moms[1::2] = np.conj(moms[1::2])                                # conjugated moments = the map mirrored in y: a different map
        rec = P.reconstruct(dev(moms), batched=True)
        for b in (0, 1, count - 1):
            r = rec[b].cpu().numpy()
            scale = np.sqrt(np.mean(r.astype(np.float64) ** 2))
            for p_ in planes:
                want = oracle_mod.reconstruct(moms[b], dim, order, oracle_mod.EXACT, rows=(p_, p_ + 1))[0]
                assert np.sqrt(np.mean((r[p_] - want) ** 2)) / scale < TOL_EXACT_BATCHED
        del rec


def test_batched_synthesis_error_paths(plans):
    from zernike3d_b200 import _lib

    P = plans(10, 16)
    m = torch.zeros((16, P.num_moments), dtype=torch.complex64, device="cuda")
    with pytest.raises(ValueError):
        P.reconstruct(m, pairs=(0, 4), batched=True)                    # whole maps only
    L = _lib.lib()
    need = L.z3d_reconstruct_batched_workspace_bytes(P._h, 16)
    assert need > 0
    ws = torch.empty(need - 1, dtype=torch.uint8, device="cuda")
    out = torch.empty((16, 16, 16, 16), dtype=torch.float32, device="cuda")
    rc = L.z3d_reconstruct_batched(P._h, m.data_ptr(), 16, out.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == _lib.ERR_WORKSPACE
    assert L.z3d_reconstruct_batched(P._h, None, 16, out.data_ptr(), ws.data_ptr(), ws.numel(), None) == _lib.ERR_INVALID
    # guard bands around the output and the workspace stay untouched (compute-sanitizer is closed on the GPU pool: an own bounds check)
    for order, dim, B in ((10, 16, 16), (12, 24, 70), (20, 32, 13)):
        P = plans(order, dim)
        m = torch.randn((B, P.num_moments), dtype=torch.complex64, device="cuda")
        need = L.z3d_reconstruct_batched_workspace_bytes(P._h, B)
        G = 1 << 16
        ws = torch.full((need + 2 * G,), 0x5A, dtype=torch.uint8, device="cuda")
        n_out = B * dim ** 3
        out = torch.full((n_out + 2 * G,), -7.5, dtype=torch.float32, device="cuda")
        rc = L.z3d_reconstruct_batched(P._h, m.data_ptr(), B, out.data_ptr() + 4 * G, ws.data_ptr() + G, need, None)
        torch.cuda.synchronize()
        assert rc == 0
        assert bool((ws[:G] == 0x5A).all()) and bool((ws[G + need:] == 0x5A).all())
        assert bool((out[:G] == -7.5).all()) and bool((out[G + n_out:] == -7.5).all())
        body = out[G:G + n_out].view(B, dim, dim, dim)
        assert bool(torch.isfinite(body).all()) and not bool((body == -7.5).any())          # every voxel written
        assert rel_l2(body[B - 1].cpu().numpy(), P.reconstruct(m[B - 1], batched=False).cpu().numpy()) < TOL_EXACT_BATCHED


def test_host_buffer_api_matches_device_api(plans):
    order, dim = 50, 64
    P = plans(order, dim)
    vols = np.stack([volume(s, dim) for s in range(5)])
    pinned = torch.from_numpy(vols).pin_memory()
    mh = P.decompose_host(pinned)
    md = P.decompose(dev(vols)).cpu().numpy()
    assert np.array_equal(mh, md)
    rh = P.reconstruct_host(mh)
    rd = P.reconstruct(dev(mh)).cpu().numpy()
    assert np.array_equal(rh, rd)
    # a batch large enough for the tensor-core tiles: the host call stages 32 volumes at a time
    P = plans(20, 32)
    vols = np.stack([volume(s, 32) for s in range(40)])
    mh = P.decompose_host(torch.from_numpy(vols).pin_memory())
    md = P.decompose(dev(vols)).cpu().numpy()
    assert np.array_equal(mh, md)
    # streaming form: two batches submitted back to back, one wait; identical to the blocking call
    pin = torch.from_numpy(vols).pin_memory()
    o1 = torch.empty((40, P.num_moments), dtype=torch.complex64).pin_memory()
    o2 = torch.empty((25, P.num_moments), dtype=torch.complex64).pin_memory()
    pin2 = torch.from_numpy(np.ascontiguousarray(vols[:25][::-1])).pin_memory()
    P.decompose_host_submit(pin, o1)
    P.decompose_host_submit(pin2, o2)
    P.host_wait()
    assert np.array_equal(o1.numpy(), md) and np.array_equal(o2.numpy(), md[:25][::-1])
    with pytest.raises(ValueError):
        P.decompose_host_submit(pin, o2)
    # ... and the host synthesis call takes the tensor-core kernel for groups of >= synth_min maps (z3d_reconstruct_batched)
    rh = P.reconstruct_host(mh)
    rd = P.reconstruct(dev(mh), batched=True).cpu().numpy()
    assert rh.shape == (40, 32, 32, 32) and np.array_equal(rh[:32], rd[:32])      # the first staged group of 32 = the same tile
    for i in (33, 39):
        assert rel_l2(rh[i], rd[i]) < TOL_EXACT                                     # (the remaining 8 go through the per-map kernel)


def test_guard_bands_around_analysis_buffers(plans):
    """compute-sanitizer is closed on the GPU pool: an own bounds check of the global writes of the analysis kernels (per-volume and
    table-free tensor-core path) - guard bands around the workspace and the moments stay untouched."""
    from zernike3d_b200 import _lib

    L = _lib.lib()
    for order, dim, B in ((10, 16, 3), (10, 16, 9), (12, 24, 70), (20, 32, 13), (25, 20, 8)):
        P = plans(order, dim)
        vols = dev(np.stack([volume(40 + i, dim) for i in range(B)]))
        need = L.z3d_workspace_bytes(P._h, B)
        G = 1 << 16
        ws = torch.full((need + 2 * G,), 0x5A, dtype=torch.uint8, device="cuda")
        n_out = B * P.num_moments * 2
        out = torch.full((n_out + 2 * G,), -7.5, dtype=torch.float32, device="cuda")
        rc = L.z3d_decompose(P._h, vols.data_ptr(), B, 0, P.half, out.data_ptr() + 4 * G, ws.data_ptr() + G, need, None)
        torch.cuda.synchronize()
        assert rc == 0
        assert bool((ws[:G] == 0x5A).all()) and bool((ws[G + need:] == 0x5A).all())
        assert bool((out[:G] == -7.5).all()) and bool((out[G + n_out:] == -7.5).all())
        got = torch.view_as_complex(out[G:G + n_out].view(B, P.num_moments, 2))
        assert torch.equal(torch.view_as_real(got), torch.view_as_real(P.decompose(vols)))


def test_error_paths(plans):
    from zernike3d_b200 import _lib

    P = plans(10, 16)
    with pytest.raises(ValueError):
        P.decompose(torch.zeros(16, 16, 8, device="cuda"))
    with pytest.raises(ValueError):
        P.decompose(torch.zeros(16, 16, 16))
    with pytest.raises(_lib.Z3DError):
        P.decompose(torch.zeros(16, 16, 16, device="cuda"), pairs=(0, 99))
    with pytest.raises(ValueError):
        P.reconstruct(torch.zeros(5, dtype=torch.complex64, device="cuda"))


def test_cli_end_to_end(tmp_path, oracle_mod, monkeypatch):
    """`zernike_master.py decomposition` -> .npz -> `reconstruct` -> .mrc, compared with the oracle."""
    import zernike_master as cli
    from zernike3d_b200 import mrc
    from zernike3d_b200.moments import Moments

    order, dim = 12, 24
    v = volume(6, dim)
    src = str(tmp_path / "in.mrc")
    mrc.write(src, v, 1.4)
    out = str(tmp_path / "mom")
    assert cli.main(f"decomposition --mode spharm --input {src} --order {order} --compute_mode 3 --output {out} --verbose".split()) == 0
    with np.load(out + ".npz") as f:
        assert sorted(f.files) == ["apix", "arr_0", "dim"] and int(f["dim"]) == dim and abs(float(f["apix"]) - 1.4) < 1e-6
        mom = f["arr_0"]
    assert rel_l2(mom, oracle_mod.decompose(v, order, oracle_mod.EXACT)) < TOL_EXACT
    rec = str(tmp_path / "rec.mrc")
    assert cli.main(f"reconstruct --mode spharm --moments {out}.npz --order {order} --dimensions {dim} --output {rec}".split()) == 0
    back = mrc.read(str(tmp_path / "rec_0.mrc")).data          # `<out>_<rec>.mrc` (zm:2714)
    assert rel_l2(back, oracle_mod.reconstruct(mom, dim, order, oracle_mod.EXACT)) < TOL_EXACT
    # --verbose: the running sum after every fifth order in the working directory (zm:2703-2705)
    from zernike3d_b200.indexing import num_moments

    monkeypatch.chdir(tmp_path)
    assert cli.main(f"reconstruct --mode spharm --moments {out}.npz --order {order} --dimensions {dim} --output {rec} --verbose".split()) == 0
    for n in (0, 5, 10):
        part = mrc.read(str(tmp_path / f"intermediate_vol_component_0n_{n:03d}.mrc")).data
        cut = mom.copy()
        cut[num_moments(n):] = 0
        assert rel_l2(part, oracle_mod.reconstruct(cut, dim, order, oracle_mod.EXACT)) < 2 * TOL_EXACT

    # 3DVA bundle in, latents + arr_0..arr_K out, then a latent-scaled reconstruction
    part, cons, comps = bundle_3dva(dim, components=2, particles=16, seed=2)
    bundle = str(tmp_path / "bundle.npz")
    np.savez(bundle, particles=part, consensus_map=cons, components=comps)
    out2 = str(tmp_path / "mom3dva.npz")
    assert cli.main(f"decomposition --mode spharm --input {bundle} --order {order} --output {out2}".split()) == 0
    M = Moments(quiet=True)
    assert M.load(out2) == 3 and M.latents.shape == (16, 2)
    assert cli.main(f"reconstruct --mode spharm --moments {out2} --order {order} --dimensions {dim} --z_ind 3 --output {rec}".split()) == 0
    scaled = mrc.read(str(tmp_path / "rec_scaled_3.mrc")).data
    w = np.concatenate([[1.0], M.latents[3]]).astype(np.float32)
    want = oracle_mod.reconstruct((w[:, None] * M.as_matrix()).sum(0).astype(np.complex64), dim, order, oracle_mod.EXACT)
    assert rel_l2(scaled, want) < 5e-6
