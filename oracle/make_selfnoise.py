"""TEST INFRASTRUCTURE ONLY - measures the float32 SELF-NOISE of the unmodified reference at the headline sizes and writes it
to ``tests/golden/ref_selfnoise.npz`` (runs only in the authoring container, needs ``/root/reference`` and ~30 GB of RAM).

The reference accumulates every moment as a float32 / complex64 BLAS dot product over all voxels (``cp.tensordot(volume, Z,
axes=3)``, zernike_master.py:2210).  Its result therefore depends on the summation order.  This script runs the reference's own
``Decompose`` twice on the same volume and the same basis arrays: once as shipped, once with volume and basis presented in
REVERSED voxel order (flipped views of the reference's own ``Y_lm`` / ``R_nl`` arrays: the same products, summed back to front).
The two runs are the same mathematical sum; their difference is the reference's own summation noise, which bounds how closely
ANY implementation can be expected to agree with it.  ``tests/test_gpu_parity.py`` uses the recorded invariant noise to justify
the invariant tolerance against the reference at 96^3 / 128^3 (against the float64 oracle the 1e-6 bar is held everywhere).

    python oracle/make_selfnoise.py            # 96^3 / order 60 and 128^3 / order 50, ~15 min
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

from make_golden import GOLDEN, invariants_of  # noqa: E402
from ref_shim import ReferenceRunner, load_reference  # noqa: E402
from zernike3d_b200.synthetic import volume  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def measure(order, dim, seed):
    import oracle

    zm = load_reference()
    t0 = time.time()
    R = ReferenceRunner(order, dim)
    v = volume(seed, dim)
    fwd = R.decompose(v)
    # the same arrays, reversed voxel order (views: no copy of the basis)
    for mem in R.Z.Y_lm:
        for key in list(R.Z.Y_lm[mem]):
            R.Z.Y_lm[mem][key] = R.Z.Y_lm[mem][key][::-1, ::-1, ::-1]
    for mem in R.Z.R_nl:
        for key in list(R.Z.R_nl[mem]):
            R.Z.R_nl[mem][key] = R.Z.R_nl[mem][key][::-1, ::-1, ::-1]
    rev = R.decompose(np.ascontiguousarray(v[::-1, ::-1, ::-1]))
    exact = oracle.decompose(v, order, oracle.EXACT)
    inv_f, inv_r, inv_e = invariants_of(zm, fwd), invariants_of(zm, rev), oracle.invariants(exact, order)
    out = dict(moments_fwd_vs_rev=rel(fwd, rev), invariants_fwd_vs_rev=rel(inv_f, inv_r),
               moments_fwd_vs_exact=rel(fwd, exact), moments_rev_vs_exact=rel(rev, exact),
               invariants_fwd_vs_exact=rel(inv_f, inv_e), invariants_rev_vs_exact=rel(inv_r, inv_e))
    print(f"[selfnoise] order {order} dim {dim} seed {seed} ({time.time() - t0:.0f} s): " +
          ", ".join(f"{k} {x:.3e}" for k, x in out.items()), flush=True)
    return out


def main():
    oracle_dir = HERE
    sys.path.insert(0, oracle_dir)
    import oracle

    oracle.build()
    res = {}
    for name, order, dim, seed in (("c5_d96_o60", 60, 96, 100), ("c3_d128_o50", 50, 128, 1)):
        for k, x in measure(order, dim, seed).items():
            res[f"{name}/{k}"] = x
    np.savez(os.path.join(GOLDEN, "ref_selfnoise.npz"), **res)


if __name__ == "__main__":
    main()
