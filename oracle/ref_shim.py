"""TEST INFRASTRUCTURE ONLY - imports the UNMODIFIED reference (``/root/reference/zernike_master.py``)
on CPU so its own mode-'3' code can be run as ground truth when generating ``tests/golden``.

The reference cannot be imported as-is here (``import cupy`` at zernike_master.py:3 fails; ``mrcfile``,
``ray``, ``zarr``, ``numcodecs`` and ``nvidia_smi`` are not installed either).  Following SURVEY.md
§8(c) we register stand-in modules *before* the import:

* ``cupy`` re-exports numpy, so every ``cp.`` call on the mode-'3' path (zernike_master.py:2281-2323,
  2194-2217, 2596-2614) executes the reference's own statement with the reference's own dtype flow
  (float64 compute -> float32 / complex64 store -> complex64 accumulate);
* the other modules are inert stubs (they are never reached on the mode-'3' path).

Nothing in the product (``zernike3d_b200/``) imports this file, and it is never used on the GPU box
(``/root/reference`` does not exist there).
"""
from __future__ import annotations

import argparse
import sys
import types

import numpy as np

REFERENCE_DIR = "/root/reference"


def _install_stubs() -> None:
    if "cupy" in sys.modules and getattr(sys.modules["cupy"], "_z3d_stub", False):
        return
    cp = types.ModuleType("cupy")
    for name in dir(np):
        if not name.startswith("__"):
            setattr(cp, name, getattr(np, name))
    cp._z3d_stub = True
    cp.asnumpy = np.asarray

    class _Pool:
        def free_all_blocks(self):
            pass

    cp.get_default_memory_pool = lambda: _Pool()
    cp.get_default_pinned_memory_pool = lambda: _Pool()
    cuda = types.ModuleType("cupy.cuda")
    runtime = types.ModuleType("cupy.cuda.runtime")
    runtime.setDevice = lambda *_a, **_k: None
    cuda.runtime = runtime
    cuda.set_allocator = lambda *_a, **_k: None
    cuda.set_pinned_memory_allocator = lambda *_a, **_k: None
    cp.cuda = cuda
    sys.modules["cupy"] = cp
    sys.modules["cupy.cuda"] = cuda
    sys.modules["cupy.cuda.runtime"] = runtime

    import scipy.special

    cupyx = types.ModuleType("cupyx")
    cupyx_scipy = types.ModuleType("cupyx.scipy")
    cupyx_special = types.ModuleType("cupyx.scipy.special")

    def sph_harm(m, n, az, pol):  # scipy >= 1.17 removed sph_harm; same argument meaning
        return scipy.special.sph_harm_y(n, m, pol, az)

    cupyx_special.sph_harm = sph_harm
    if not hasattr(scipy.special, "sph_harm"):
        scipy.special.sph_harm = sph_harm
    cupyx.scipy = cupyx_scipy
    cupyx_scipy.special = cupyx_special
    sys.modules["cupyx"] = cupyx
    sys.modules["cupyx.scipy"] = cupyx_scipy
    sys.modules["cupyx.scipy.special"] = cupyx_special

    ray = types.ModuleType("ray")
    ray.init = lambda *_a, **_k: None
    ray.shutdown = lambda *_a, **_k: None
    ray.put = lambda x: x
    ray.get = lambda x: x
    ray.wait = lambda ids, **_k: (ids[:1], ids[1:])

    def remote(*dargs, **dkwargs):
        def wrap(fn):
            target = fn.__func__ if isinstance(fn, staticmethod) else fn
            target.remote = target
            return fn

        if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
            return wrap(dargs[0])
        return wrap

    ray.remote = remote
    sys.modules["ray"] = ray

    for name in ("mrcfile", "zarr"):
        sys.modules.setdefault(name, types.ModuleType(name))
    numcodecs = types.ModuleType("numcodecs")
    numcodecs.Blosc = type("Blosc", (), {"BITSHUFFLE": 2, "__init__": lambda self, *a, **k: None})
    sys.modules.setdefault("numcodecs", numcodecs)

    smi = types.ModuleType("nvidia_smi")
    smi.nvmlInit = lambda: None
    smi.nvmlShutdown = lambda: None
    smi.nvmlDeviceGetCount = lambda: 1
    smi.nvmlDeviceGetHandleByIndex = lambda i: i
    smi.nvmlDeviceGetName = lambda h: "stub"
    smi.nvmlDeviceGetMemoryInfo = lambda h: types.SimpleNamespace(free=10**12, total=10**12, used=0)
    sys.modules["nvidia_smi"] = smi


def load_reference():
    """Return the reference's ``zernike_master`` module, imported unmodified from ``/root/reference``
    under the private name ``_ref_zernike_master`` (the repo root holds same-named drop-in modules, so the
    reference and its ``indexing`` import are resolved against ``/root/reference`` only)."""
    import importlib.util

    if "_ref_zernike_master" in sys.modules:
        return sys.modules["_ref_zernike_master"]
    _install_stubs()
    saved = {k: sys.modules.pop(k) for k in ("indexing", "zernike_master") if k in sys.modules}
    sys.path.insert(0, REFERENCE_DIR)
    try:
        spec = importlib.util.spec_from_file_location("_ref_zernike_master", REFERENCE_DIR + "/zernike_master.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_ref_zernike_master"] = mod
        spec.loader.exec_module(mod)
        ref_ix = sys.modules.pop("indexing")
        if not ref_ix.__file__.startswith(REFERENCE_DIR):
            raise RuntimeError(f"reference imported {ref_ix.__file__} instead of its own indexing.py")
    finally:
        sys.path.remove(REFERENCE_DIR)
        sys.modules.update(saved)
    return mod


def reference_args(order: int, **over) -> argparse.Namespace:
    ns = dict(order=order, compute_mode="3", gpu_id=0, bit8=False, parity=False, verbose=False,
              cpu_tasks=1, save=False, mixed=False, apix=None, output="o.npz", save_bit8=False,
              z_ind=None, scales=None)
    ns.update(over)
    return argparse.Namespace(**ns)


class ReferenceRunner:
    """Drives the reference's ``ZernikeSpherical`` mode-'3' methods exactly as ``run()`` does
    (zernike_master.py:2878-2889 for decomposition, 2900-2915 for reconstruction)."""

    def __init__(self, order: int, dim: int):
        zm = load_reference()
        self.zm = zm
        self.Z = zm.ZernikeSpherical(reference_args(order))
        self.Z.dim = dim
        self.jobs = self.Z.CalculatePolynomials()
        self.order, self.dim = order, dim

    def decompose(self, vol: np.ndarray) -> np.ndarray:
        """complex64 moments in the reference's save order (zernike_master.py:378-382)."""
        self.Z.moments.moments = {}
        self.Z.Decompose(vol, self.dim, self.jobs)
        out = self.Z.moments._save_moments(self.Z.moments.moments)
        self.Z.moments.moments = {}
        return out

    def reconstruct(self, moments: np.ndarray) -> np.ndarray:
        """float32 map the reference writes for a single moment vector (zernike_master.py:2686-2715)."""
        zm = self.zm
        M = zm.Moments()
        for mu, mom in enumerate(moments):
            n, l, m = zm.mu_to_nlm(mu)
            M.set_moment(n, l, m, mom)
        M.add_dimension()
        M.stats(quiet=True)
        self.Z.moments = M
        captured = {}

        def capture(volume, name):
            captured[name] = np.array(volume, np.float32)

        self.Z.write_mrc_file = capture
        self.Z.Reconstruct(self.jobs)
        (out,) = captured.values()
        self.Z.moments = zm.Moments()
        return out
