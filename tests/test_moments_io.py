"""Wire formats either side of the path: the `.npz` moment layout (zm:334-360, 268-296; README.md:61-98) and MRC."""
import os

import numpy as np
import pytest

from zernike3d_b200 import indexing as ix
from zernike3d_b200 import mrc
from zernike3d_b200.moments import Moments
from zernike3d_b200.synthetic import bundle_3dva, volume


def _vec(order, seed):
    rng = np.random.default_rng(seed)
    n = ix.num_moments(order)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def test_single_volume_npz_layout(tmp_path):
    M = Moments(quiet=True)
    v = _vec(12, 0)
    M.from_array(v)
    out = str(tmp_path / "single.npz")
    M.save(output=out, compress=False, dim=32, apix=1.4)
    with np.load(out) as f:
        assert sorted(f.files) == ["apix", "arr_0", "dim"]  # the 3-key layout `load` sniffs (zm:271)
        assert f["arr_0"].dtype == np.complex64 and np.array_equal(f["arr_0"], v)
        assert int(f["dim"]) == 32 and float(f["apix"]) == 1.4
    L = Moments(quiet=True)
    assert L.load(out) == 1
    assert L.order == 12 and L.dimensions == 1
    assert list(L.ndmoments[0].keys())[:4] == [(0, 0, 0), (1, 1, 0), (1, 1, 1), (2, 0, 0)]
    assert L.get_moment(5, 3, 2) == v[ix.nlm_to_mu((5, 3, 2))]
    assert np.array_equal(L.as_array(0), v)


def test_3dva_npz_layout_and_scaling(tmp_path):
    part, _, _ = bundle_3dva(8, components=3, particles=20, seed=2)
    M = Moments(quiet=True)
    vecs = [_vec(8, s) for s in range(4)]
    for v in vecs:
        M.from_array(v)
        M.add_dimension()
    M.set_latents(particles=part, components=3)
    assert M.latents.shape == (20, 3) and M.latents.dtype == np.float32
    out = str(tmp_path / "bundle.npz")
    M.save(output=out, dim=8, apix=1)
    with np.load(out) as f:
        assert set(f.files) == {"latents", "dim", "apix", "arr_0", "arr_1", "arr_2", "arr_3"}
    L = Moments(quiet=True)
    assert L.load(out) == 4
    assert np.array_equal(L.as_matrix(), np.stack(vecs))
    # apply_scaling: Omega = [1, latents[z]] . Omega_dims (zm:78-95)
    scaled = L.apply_scaling(z_ind=7)
    want = (np.concatenate([[1.0], part_lat(part, 7, 3)]).astype(np.float32)[:, None] * np.stack(vecs)).sum(0)
    assert np.allclose(scaled, want, rtol=1e-6, atol=1e-6)
    assert L.scaled_moments[(2, 2, 1)] == scaled[ix.nlm_to_mu((2, 2, 1))]
    scaled2 = L.apply_scaling(scales=["-34", "0", "2.5"])  # strings, as argparse delivers them
    want2 = vecs[0] - 34 * vecs[1] + 2.5 * vecs[3]
    assert np.allclose(scaled2, want2, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        L.apply_scaling(scales=["1", "2"])


def part_lat(part, z, k):
    return np.array([part[f"components_mode_{i}/value"][z] for i in range(k)], dtype=np.float32)


def test_invariants_definition():
    M = Moments(quiet=True)
    v = _vec(6, 3)
    M.from_array(v)
    M.add_dimension()
    M.calculate_invariants()
    t = ix.nlm_table(6)
    for (n, l), val in M.ndinvariants[0].items():
        sel = v[(t[:, 0] == n) & (t[:, 1] == l)]
        want = np.sqrt(abs(sel[0]) ** 2 + 2 * np.sum(np.abs(sel[1:]) ** 2))  # zm:65-74
        assert abs(val - want) < 1e-5 * want
    assert len(M.ndinvariants[0]) == ix.num_radial(6)


def test_8bit_roundtrip_is_lossy_but_loadable(tmp_path):
    M = Moments(quiet=True)
    M.from_array(_vec(5, 1))
    M.save(output=str(tmp_path / "m.npz"), compress=True, dim=16, apix=1)
    L = Moments(quiet=True)
    assert L.load(str(tmp_path / "m.8bit.npz")) == 1 and L.order == 5


def test_incomplete_vector_rejected():
    with pytest.raises(ValueError):
        Moments(quiet=True).from_array(np.zeros(100, np.complex64))


def test_mrc_roundtrip(tmp_path):
    v = volume(3, 12)
    p = str(tmp_path / "v.mrc")
    mrc.write(p, v, voxel_size=1.4)
    assert os.path.getsize(p) == 1024 + v.nbytes
    m = mrc.read(p)
    assert m.data.shape == (12, 12, 12) and m.data.dtype == np.float32 and np.array_equal(m.data, v)
    assert abs(m.voxel_size["x"] - 1.4) < 1e-6 and abs(m.voxel_size.z - 1.4) < 1e-6
    raw = open(p, "rb").read(1024)
    assert raw[208:212] == b"MAP " and np.frombuffer(raw[:16], "<i4").tolist() == [12, 12, 12, 2]
    with mrc.open(p, mode="w+") as out:  # the reference's write idiom (zm:2718-2721)
        out.set_data(np.array(v * 2, np.float32))
        out.voxel_size = 2
    with mrc.open(p) as back:
        assert np.array_equal(back.data, v * 2) and back.voxel_size["x"] == 2.0
    # int16 big-endian input
    hdr = np.zeros(256, ">i4")
    hdr[0:3] = (2, 3, 4); hdr[3] = 1; hdr[7:10] = (2, 3, 4)
    hdr.view(">f4")[10:13] = (2.0, 3.0, 4.0)
    b = bytearray(hdr.tobytes()); b[208:212] = b"MAP "; b[212:216] = b"\x11\x11\x00\x00"
    data = np.arange(24, dtype=">i2")
    q = str(tmp_path / "be.mrc")
    open(q, "wb").write(bytes(b) + data.tobytes())
    m = mrc.read(q)
    assert m.data.shape == (4, 3, 2) and m.data.ravel().tolist() == list(range(24)) and m.voxel_size.x == 1.0
    open(q, "wb").write(b"short")
    with pytest.raises(ValueError):
        mrc.read(q)


def test_zcc_curve_matches_the_reference_formula():
    """Moments.zcc against a literal restatement of Z_3DVA.calc_ZCC (moment_analysis_3dva.py:402-422)."""
    from zernike3d_b200.moments import Moments

    order = 9
    rng = np.random.default_rng(4)
    A, B = Moments(), Moments()
    for M in (A, B):
        for n in range(order + 1):
            for l in range(n % 2, n + 1, 2):
                for m in range(l + 1):
                    M.set_moment(n, l, m, np.complex64(rng.normal() + 1j * rng.normal() * (m > 0)))
        M.add_dimension()
    got = A.zcc(B)
    want = []
    a, b = A.ndmoments[0], B.ndmoments[0]
    for n in range(order + 1):          # calc_ZCC, line by line
        _sum, xs, ys = [], [], []
        for l in range(n + 1):
            if (n - l) % 2 == 0:
                for m in range(l + 1):
                    momA = a[(n, l, m)]
                    momB = np.conjugate(b[(n, l, m)])
                    _sum.append(np.abs(momA * momB))
                    xs.append(momA)
                    ys.append(momB)
        want.append(np.sum(_sum) / np.sqrt(np.sum(np.abs(xs) ** 2) * np.sum(np.abs(ys) ** 2)))
    assert got.shape == (order + 1,) and np.allclose(got, want, rtol=1e-6)
    assert np.allclose(A.zcc(A), 1.0, atol=1e-6)


def test_zcc_and_latent_scaling_against_reference_run_fixture():
    """Moments.zcc / Moments.apply_scaling against tests/golden/ref_zcc_scaling.npz: curves produced by the reference's OWN calc_ZCC statements
    (moment_analysis_3dva.py:402-421, executed unchanged by oracle/make_zcc_golden.py) and maps of moments scaled by the reference's own
    Moments.apply_scaling (zm:78-95)."""
    from conftest import golden
    from zernike3d_b200.moments import Moments

    g = golden("zcc_scaling")
    sets = np.asarray(g["moment_sets"], np.complex64)
    M = Moments(quiet=True)
    for v in sets:
        M.from_array(v)
        M.add_dimension()
    for (i, j), want in zip(g["pairs"], g["zcc"]):
        got = M.zcc(M, int(i), int(j))
        assert got.shape == (int(g["order"]) + 1,) and np.allclose(got, want, rtol=2e-6, atol=1e-7)
    b = golden("3dva_d24_o12")
    B = Moments(quiet=True)
    for v in np.asarray(b["moments"], np.complex64):
        B.from_array(v)
        B.add_dimension()
    B.latents = np.asarray(b["latents"], np.float32)
    B.stats(quiet=True)
    for z, want in zip(g["scaling_z"], g["scaled"]):
        got = np.asarray(B.apply_scaling(z_ind=int(z)))
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-6
