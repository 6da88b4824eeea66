"""zernike3d_b200 - B200-native spharm 3D-Zernike decomposition / reconstruction (drop-in for that path of
charbj/zernike3d).  Compute lives in hand-written sm_100a kernels behind the C ABI in ``include/``."""
from ._lib import Z3DError  # noqa: F401

__all__ = ["Z3DError"]
