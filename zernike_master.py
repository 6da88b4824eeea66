#!/usr/bin/env python
"""Drop-in ``zernike_master.py`` for the spharm path of charbj/zernike3d, running on the B200-native
kernels of ``zernike3d_b200``.

    python zernike_master.py decomposition --mode spharm --input map.mrc --order 50 --output moments.npz
    python zernike_master.py reconstruct   --mode spharm --moments moments.npz --order 50 --dimensions 64 --output rec.mrc

Sub-commands, flag names, defaults and file formats are the reference's (zernike_master.py:2917-3216,
README).  Flags that only steer the reference's materialised-basis memory tiers are accepted and ignored
(see ``zernike3d_b200.spherical``).  ``--mode monomial`` is refused: in the reference it cannot run either
(``ZernikeMonomial.__init__`` takes no ``args``, zm:3209 vs zm:551).
"""
from __future__ import annotations

import argparse
import os
import sys

from zernike3d_b200.indexing import *  # noqa: F401,F403  (the reference re-exports indexing the same way, zm:22)
from zernike3d_b200.moments import Moments  # noqa: F401
from zernike3d_b200.spherical import ZernikeSpherical  # noqa: F401


def _ext_check(allowed, msg):
    def check(param):
        if os.path.splitext(param)[1].lower() not in allowed:
            raise argparse.ArgumentTypeError(msg)
        return param

    return check


def _npz_output(param):
    return param if os.path.splitext(param)[1].lower() == ".npz" else param + ".npz"


def build_parser() -> argparse.ArgumentParser:
    top = argparse.ArgumentParser(
        prog="zernike.py",
        description="ZERNIKE - 3D Zernike decomposition and reconstruction of cryo-EM volumes (B200-native spharm path)",
        formatter_class=argparse.RawTextHelpFormatter)
    subs = top.add_subparsers(title="Operating modes", dest="operating_mode", required=True,
                              description="Select the operating mode: `decomposition` or `reconstruct`.")

    valid_input = _ext_check((".npz", ".mrc"), "File must have a .npz or .mrc extension")
    valid_mrc = _ext_check((".mrc",), "File must have a .mrc extension. See README.")
    valid_com = _ext_check((".com", ".cmm"), "File must have a .com or .cmm extension. See README.")

    def common(p):
        p.add_argument("--order", type=int, required=True, help="Order `n` of Zernike components to calculate.")
        p.add_argument("--mode", choices=["spharm", "monomial"], type=str, required=True,
                       help="Algorithm; only `spharm` is implemented (stable to n > 100).")
        p.add_argument("--apix", type=float, required=False, help="Override the pixel size (Angstrom per pixel).")
        p.add_argument("--compute_mode", type=str, help="Accepted for compatibility; ignored (no materialised basis).")
        p.add_argument("--gpu_id", type=int, help="Force GPU device ID. [Default: current device]")
        p.add_argument("--cpu_tasks", type=int, help="Accepted for compatibility; ignored.")
        p.add_argument("--bit8", action="store_true", help="Accepted for compatibility; ignored.")
        p.add_argument("--parity", action="store_true", help="Accepted for compatibility; ignored "
                                                             "(mirror symmetry is always exploited inside the kernels).")
        p.add_argument("--mixed", action="store_true", help="Accepted for compatibility; ignored.")
        p.add_argument("--save", action="store_true", help="Accepted for compatibility; ignored.")
        p.add_argument("--verbose", action="store_true", help="Make verbose.")
        p.add_argument("--preallocate_memory_off", action="store_true", help="Accepted for compatibility; ignored.")

    dec = subs.add_parser("decomposition", help="Decompose an input map to Zernike moments")
    dec.add_argument("--input", type=valid_input, required=True,
                     help="Input map (.mrc) or 3DVA bundle (WIGGLE .npz) to be decomposed.")
    dec.add_argument("--output", type=_npz_output, required=True, help="Output moment file (.npz appended if missing).")
    dec.add_argument("--com", type=valid_com, required=False, help="ChimeraX marker file giving the crop centre.")
    dec.add_argument("--crop_to", type=int, required=False, help="Crop box to this size (pixels).")
    dec.add_argument("--bin", type=float, default=None, required=False, help="Bin the input volume by this factor.")
    dec.add_argument("--save_bit8", action="store_true", help="Also save the moments in 8bit encoding.")
    dec.add_argument("--onthefly", action="store_true", help="Accepted for compatibility; the basis is always generated on the fly.")
    common(dec)

    rec = subs.add_parser("reconstruct", help="Reconstruct the object from Zernike moments")
    rec.add_argument("--moments", type=valid_input, required=True, help="Complex coefficients ('moments.npz')")
    rec.add_argument("--output", type=valid_mrc, required=True, help="Name of the output volume.")
    rec.add_argument("--dimensions", type=int, default=None, help="Reconstruction box size.")
    rec.add_argument("--z_ind", type=int, default=None, help="Index of the particle whose latent coordinates scale the moments.")
    rec.add_argument("--scales", nargs="+", default=None, help="Scaling factors e.g. 4.5 78 2 ... (one per component).")
    common(rec)
    return top


def main(argv=None) -> int:
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.mode != "spharm":
        parser.error("only `--mode spharm` is implemented (the reference's monomial path cannot run either)")
    if args.operating_mode == "decomposition":
        if args.com and args.crop_to is None:
            parser.error("`--com` flag requires `--crop_to BOX_SIZE` e.g. --com --crop_to 32")
        ZernikeSpherical(args).run(decompose=True)
    else:
        ZernikeSpherical(args).run(reconstruct=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
