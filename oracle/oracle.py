"""TEST INFRASTRUCTURE ONLY - ctypes wrapper around ``oracle/libzernike_oracle.so`` (the C restatement
of the reference's mode-'3' path, see ``zernike_oracle.c``).  Importable only from ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs; the product never imports it."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libzernike_oracle.so")
FAITHFUL, EXACT = 0, 1
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "zernike_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libzernike_oracle.so"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        L.zo_num_moments.argtypes = [ctypes.c_int]
        L.zo_num_radial.argtypes = [ctypes.c_int]
        L.zo_decompose.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, dp]
        L.zo_reconstruct.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        L.zo_basis_at.argtypes = [ctypes.c_int] * 6 + [dp, dp]
        L.zo_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def num_moments(order: int) -> int:
    return lib().zo_num_moments(order)


def num_radial(order: int) -> int:
    return lib().zo_num_radial(order)


def max_threads() -> int:
    return lib().zo_max_threads()


def set_threads(n: int) -> None:
    lib().zo_set_threads(n)


def decompose(vol: np.ndarray, order: int, mode: int = FAITHFUL, rows=None, return_f64: bool = False):
    """complex64 moments in sorted (n, l, m) order; ``rows=(lo, hi)`` restricts the voxel sum to
    array-axis-0 rows [lo, hi) (unreduced slab partial)."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    dim = vol.shape[0]
    assert vol.shape == (dim, dim, dim)
    lo, hi = (0, dim) if rows is None else rows
    nm = num_moments(order)
    o32 = np.empty(2 * nm, np.float32)
    o64 = np.empty(2 * nm, np.float64)
    rc = lib().zo_decompose(_fp(vol), dim, order, mode, lo, hi, _fp(o32), _dp(o64))
    if rc:
        raise ValueError(f"zo_decompose failed rc={rc}")
    if return_f64:
        return o64.view(np.complex128)
    return o32.view(np.complex64)


def reconstruct(moments: np.ndarray, dim: int, order: int, mode: int = FAITHFUL, rows=None) -> np.ndarray:
    mom = np.ascontiguousarray(moments, dtype=np.complex64)
    assert mom.shape == (num_moments(order),)
    lo, hi = (0, dim) if rows is None else rows
    out = np.empty((hi - lo, dim, dim), np.float32)
    rc = lib().zo_reconstruct(_fp(mom.view(np.float32)), dim, order, mode, lo, hi, _fp(out))
    if rc:
        raise ValueError(f"zo_reconstruct failed rc={rc}")
    return out


def basis_at(dim: int, order: int, i: int, j: int, k: int, mode: int = FAITHFUL):
    """(R[sorted (n,l)], Y[(l,m)] complex) at voxel [i, j, k]."""
    rad = np.empty(num_radial(order), np.float64)
    y = np.empty((order + 1) * (order + 2), np.float64)
    lib().zo_basis_at(dim, order, mode, i, j, k, _dp(rad), _dp(y))
    return rad, y.view(np.complex128)


def invariants(moments: np.ndarray, order: int) -> np.ndarray:
    """F_nl = || (|O_nl0|, |O_nlm|, |O_nlm| for m > 0) ||_2 in sorted (n, l) order
    (Moments.calculate_invariants zm:62-76)."""
    mom = np.asarray(moments)
    out, mu = [], 0
    for n in range(order + 1):
        for l in range(n % 2, n + 1, 2):
            a = np.abs(mom[mu:mu + l + 1]).astype(np.float64)
            out.append(np.sqrt(a[0] ** 2 + 2.0 * np.sum(a[1:] ** 2)))
            mu += l + 1
    return np.array(out)
