"""TEST INFRASTRUCTURE ONLY - reference-run fixtures for the two 3DVA helpers whose device versions were so far only checked against this
repository's own host restatements (runs only in the authoring container, needs ``/root/reference``):

* ``Z_3DVA.calc_ZCC`` (moment_analysis_3dva.py:402-422).  The module cannot be imported here (cupy, joypy, mrcfile, matplotlib back-ends), and
  the method plots; its ARITHMETIC lines (402-419, up to ``c.append(nom/dom)``) do not touch ``self``, so they are cut out of the reference
  file as text and executed unchanged - every statement that produces the curve is the reference's own line.  (It hardcodes order 50.)
* ``Moments.apply_scaling`` (zernike_master.py:78-95) through the imported, unmodified ``Moments`` class (``oracle/ref_shim.py``).

    python oracle/make_zcc_golden.py        # -> tests/golden/ref_zcc_scaling.npz
"""
from __future__ import annotations

import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from ref_shim import REFERENCE_DIR, load_reference  # noqa: E402
from zernike3d_b200.indexing import nlm_table  # noqa: E402


def reference_calc_zcc():
    """The reference's own calc_ZCC statements up to the curve, as a plain function ``f(mapA, mapB) -> list``."""
    lines = open(os.path.join(REFERENCE_DIR, "moment_analysis_3dva.py")).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.strip().startswith("def calc_ZCC("))
    stop = next(i for i in range(start, len(lines)) if "c.append(nom/dom)" in lines[i])
    body = textwrap.dedent("\n".join(lines[start + 1:stop + 1]))
    src = "def calc_zcc_reference(mapA, mapB):\n" + textwrap.indent(body, "    ") + "\n    return c\n"
    ns = {"np": np}
    exec(compile(src, "moment_analysis_3dva.py:calc_ZCC", "exec"), ns)
    return ns["calc_zcc_reference"], (start + 1, stop + 1)


def main():
    order = 50
    nlm = [tuple(int(x) for x in r) for r in nlm_table(order)]
    g1 = np.load(os.path.join(ROOT, "tests", "golden", "ref_c1_d64_o50.npz"), allow_pickle=True)
    g3 = np.load(os.path.join(ROOT, "tests", "golden", "ref_c3_d128_o50.npz"), allow_pickle=True)
    a, b = np.asarray(g1["moments"][0], np.complex64), np.asarray(g3["moments"][0], np.complex64)
    rng = np.random.default_rng(17)
    noisy = (a + (0.3 * np.abs(a).mean()) * (rng.normal(size=a.shape) + 1j * rng.normal(size=a.shape))).astype(np.complex64)
    sets = np.stack([a, b, noisy, np.conj(a)])
    f, (l0, l1) = reference_calc_zcc()
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (0, 0)]
    curves = []
    for i, j in pairs:
        A = {k: sets[i][n] for n, k in enumerate(nlm)}
        B = {k: sets[j][n] for n, k in enumerate(nlm)}
        curves.append(np.asarray(f(A, B), dtype=np.float64))
    # apply_scaling through the reference's own Moments class: the golden 3DVA bundle (reference-made latents)
    ref = load_reference()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_3dva_d24_o12.npz"), allow_pickle=True)
    dims, lat = np.asarray(g["moments"], np.complex64), np.asarray(g["latents"], np.float32)
    o3 = int(g["order"])
    nlm3 = [tuple(int(x) for x in r) for r in nlm_table(o3)]
    M = ref.Moments()
    for d in range(dims.shape[0]):
        for n, k in enumerate(nlm3):
            M.set_moment(k[0], k[1], k[2], dims[d][n])
        M.add_dimension()
    M.latents = lat
    M.order, M.dimensions = o3, dims.shape[0]      # what Moments.stats() / the .npz loader set (zm:97-113)
    zs = [0, 3, 11, 39]
    scaled = []
    for z in zs:
        M.apply_scaling(z_ind=z)                   # -> M.scaled_moments {(n, l, m): complex}
        scaled.append(np.asarray([M.scaled_moments[k] for k in nlm3], np.complex64))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_zcc_scaling.npz"), order=order, moment_sets=sets, pairs=np.asarray(pairs),
                        zcc=np.stack(curves), calc_zcc_lines=np.asarray([l0, l1]), scaling_order=o3, scaling_z=np.asarray(zs),
                        scaled=np.stack(scaled))
    print("calc_ZCC lines", l0, l1, "curves", np.stack(curves)[:, :4], "scaled", np.stack(scaled).shape)


if __name__ == "__main__":
    main()
