"""Pins the CPU oracle (oracle/zernike_oracle.c) to outputs of the UNMODIFIED reference, generated in the
authoring container by oracle/make_golden.py (tests/golden/ref_*.npz).  Tolerances: the reference sums its
float32 products inside BLAS in an unspecified order (zm:2210), so agreement is norm-wise."""
import numpy as np
import pytest

from conftest import golden, rel_l2
from zernike3d_b200.synthetic import bundle_3dva, volume

SMALL = ["d16_o10", "d15_o8", "d32_o20", "d20_o25"]


@pytest.mark.parametrize("name", SMALL + ["c1_d64_o50"])
def test_decompose_matches_reference(oracle_mod, name):
    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    for si, seed in enumerate(g["seeds"]):
        v = volume(int(seed), dim)
        ref = g["moments"][si]
        tol = 5e-7 if dim <= 32 else 4e-6  # 64^3: the reference's own BLAS float32 summation noise is 2.1e-6
        for mode in (oracle_mod.FAITHFUL, oracle_mod.EXACT):
            mom = oracle_mod.decompose(v, order, mode)
            assert mom.dtype == np.complex64 and mom.shape == ref.shape
            assert rel_l2(mom, ref) < tol
        inv = oracle_mod.invariants(oracle_mod.decompose(v, order), order)
        assert rel_l2(inv, g["invariants"][si]) < 1e-6
        # m = 0 moments are purely real (README.md:91-97)
        mu = 0
        for n in range(order + 1):
            for l in range(n % 2, n + 1, 2):
                assert abs(mom[mu].imag) <= 1e-7 * (abs(mom[mu].real) + 1e-30) or mom[mu].imag == 0
                mu += l + 1


@pytest.mark.parametrize("name", SMALL + ["c1_d64_o50"])
def test_reconstruct_matches_reference(oracle_mod, name):
    g = golden(name)
    order, dim = int(g["order"]), int(g["dim"])
    for si in range(len(g["seeds"])):
        rec = oracle_mod.reconstruct(g["moments"][si], dim, order, oracle_mod.FAITHFUL)
        assert rec.shape == (dim, dim, dim) and rec.dtype == np.float32
        assert rel_l2(rec, g["recon"][si]) < 5e-7
        assert rel_l2(oracle_mod.reconstruct(g["moments"][si], dim, order, oracle_mod.EXACT), g["recon"][si]) < 1e-6


def test_3dva_bundle_matches_reference(oracle_mod):
    g = golden("3dva_d24_o12")
    order, dim = int(g["order"]), int(g["dim"])
    part, cons, comps = bundle_3dva(dim, components=5, particles=64, seed=2)
    for i, v in enumerate([cons, *comps]):
        assert rel_l2(oracle_mod.decompose(v, order), g["moments"][i]) < 5e-7
    from zernike3d_b200.moments import Moments

    M = Moments(quiet=True)
    M.set_latents(particles=part, components=5)
    assert np.array_equal(M.latents, g["latents"])


def test_slab_partials_add_up(oracle_mod):
    v = volume(7, 16)
    full = oracle_mod.decompose(v, 10, oracle_mod.EXACT, return_f64=True)
    parts = sum(oracle_mod.decompose(v, 10, oracle_mod.EXACT, rows=r, return_f64=True) for r in [(0, 5), (5, 11), (11, 16)])
    assert rel_l2(parts, full) < 1e-12


def test_round_trip_gain(oracle_mod):
    # zm:2198 carries an extra 1/(4 pi): decompose -> reconstruct returns ~ f / (4 pi) ((N-1)/N)^3
    dim, order = 24, 30
    v = volume(11, dim)
    rec = oracle_mod.reconstruct(oracle_mod.decompose(v, order, oracle_mod.EXACT), dim, order, oracle_mod.EXACT)
    gain = float(np.vdot(rec.ravel(), v.ravel()) / np.vdot(v.ravel(), v.ravel()))
    assert abs(gain / ((1 / (4 * np.pi)) * ((dim - 1) / dim) ** 3) - 1) < 0.15
