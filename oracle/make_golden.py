"""TEST INFRASTRUCTURE ONLY - generates ``tests/golden/ref_*.npz`` by running the UNMODIFIED reference
(mode '3', zernike_master.py:2281-2323 / 2194-2217 / 2596-2614, via ``ref_shim``) on seeded synthetic
volumes.  Runs only in the authoring container (needs ``/root/reference``); the fixtures it writes are
committed so the GPU box never needs the reference.

    python oracle/make_golden.py small          # seconds-minutes
    python oracle/make_golden.py c1             # 64^3 order 50 (+ reconstruction), ~2 min, 3.5 GB
    python oracle/make_golden.py c3             # 128^3 order 50 moments, ~28 GB RAM
    python oracle/make_golden.py c5             # 96^3 order 60 moments, ~17 GB RAM
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_shim import ReferenceRunner, load_reference  # noqa: E402
from zernike3d_b200.synthetic import bundle_3dva, volume  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def invariants_of(zm, moments: np.ndarray) -> np.ndarray:
    """F_nl via the reference's own ``Moments.calculate_invariants`` (zernike_master.py:62-76),
    flattened in sorted (n, l) order."""
    M = zm.Moments()
    for mu, mom in enumerate(moments):
        n, l, m = zm.mu_to_nlm(mu)
        M.set_moment(n, l, m, mom)
    M.add_dimension()
    M.calculate_invariants()
    inv = M.ndinvariants[0]
    return np.array([inv[k] for k in sorted(inv)], dtype=np.float64)


def case(name, order, dim, seeds, recon=True):
    zm = load_reference()
    t0 = time.time()
    R = ReferenceRunner(order, dim)
    t_basis = time.time() - t0
    moms, invs, recs, t_dec, t_rec = [], [], [], [], []
    for s in seeds:
        v = volume(s, dim)
        t0 = time.time()
        m = R.decompose(v)
        t_dec.append(time.time() - t0)
        moms.append(m)
        invs.append(invariants_of(zm, m))
        if recon:
            t0 = time.time()
            recs.append(R.reconstruct(m))
            t_rec.append(time.time() - t0)
    out = dict(order=order, dim=dim, seeds=np.array(seeds), moments=np.stack(moms),
               invariants=np.stack(invs),
               timing=np.array([t_basis, float(np.mean(t_dec)), float(np.mean(t_rec)) if t_rec else 0.0]))
    if recon:
        out["recon"] = np.stack(recs).astype(np.float32)
    np.savez_compressed(os.path.join(GOLDEN, f"ref_{name}.npz"), **out)
    print(f"[golden] {name}: order {order} dim {dim} basis {t_basis:.1f}s decompose {np.mean(t_dec):.1f}s"
          + (f" reconstruct {np.mean(t_rec):.1f}s" if t_rec else ""), flush=True)


def case_3dva(name, order, dim):
    """3DVA bundle through the reference's own multi-volume loop (zernike_master.py:2886-2890)."""
    zm = load_reference()
    part, cons, comps = bundle_3dva(dim, components=5, particles=64, seed=2)
    R = ReferenceRunner(order, dim)
    moms = [R.decompose(v) for v in [cons, *comps]]
    M = zm.Moments()
    M.set_latents(particles=part, components=5)
    np.savez_compressed(os.path.join(GOLDEN, f"ref_{name}.npz"), order=order, dim=dim,
                        moments=np.stack(moms), latents=M.latents)
    print(f"[golden] {name}: 3DVA bundle order {order} dim {dim}", flush=True)


def main(which):
    os.makedirs(GOLDEN, exist_ok=True)
    if which in ("small", "all"):
        case("d16_o10", 10, 16, [0, 5])
        case("d15_o8", 8, 15, [3])            # odd box: r = 0 voxel, nan_to_num path (SURVEY §7 hard part 5)
        case("d32_o20", 20, 32, [1])
        case("d20_o25", 25, 20, [4])          # order > dim
        case_3dva("3dva_d24_o12", 12, 24)
    if which in ("c1", "all"):
        case("c1_d64_o50", 50, 64, [0])
    if which in ("c3", "all"):
        case("c3_d128_o50", 50, 128, [1], recon=False)
    if which in ("c5", "all"):
        case("c5_d96_o60", 60, 96, [100], recon=False)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small")
