"""Top-level ``indexing`` module, importable as in the reference (``from indexing import *``, zm:22)."""
from zernike3d_b200.indexing import *  # noqa: F401,F403
from zernike3d_b200.indexing import __all__  # noqa: F401
