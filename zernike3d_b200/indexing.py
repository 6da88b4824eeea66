"""Flat-index maps between (n, l, m) / (n, l) / (l, m) and the reference's array positions.

Same public names and semantics as the reference's ``indexing.py`` (ix:13-228), because the order of
the complex64 vector in a saved ``.npz`` *is* this map (zernike_master.py:378-382 sorts the (n, l, m)
keys; ``Moments.load`` inverts it with ``mu_to_nlm``, zm:277-279).  The reference walks the index space
with numba loops; here every map is a closed form (plus vectorised table builders the kernels' host
side uses).

Ordering: n ascending; within n, l = n mod 2, +2, ..., n; within l, m = 0..l (m >= 0 only).
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "mu_R", "nl_to_mu_R", "mu_R_to_nl", "mu_Y", "lm_to_mu_Y", "mu_Y_to_lm", "mu_", "nlm_to_mu", "mu_to_nlm",
    "memory_requirement", "num_moments", "num_radial", "num_harmonics", "nlm_table", "nl_table",
]


# ----------------------------------------------------------------------------- counts
def num_radial(n: int) -> int:
    """Number of (n', l) pairs with n' <= n (676 at n = 50)."""
    if n < 0:
        return 0
    h = n // 2
    # sum_{k=0..n} (k//2 + 1): pairs (2j, 2j+1) contribute 2(j+1)
    return (h + 1) * (h + 2) - (0 if n % 2 else h + 1)


def num_harmonics(l: int) -> int:
    """Number of (l', m >= 0) pairs with l' <= l (1326 at l = 50)."""
    return (l + 1) * (l + 2) // 2 if l >= 0 else 0


def _moments_in_shell(n: int) -> int:
    """Number of (l, m >= 0) with l <= n, l = n mod 2: sum of (l + 1)."""
    h = n // 2
    return (h + 1) * (h + 1) if n % 2 == 0 else (h + 1) * (h + 2)


def num_moments(n: int) -> int:
    """Number of (n', l, m >= 0) triples with n' <= n (12 051 at n = 50)."""
    if n < 0:
        return 0
    k = n // 2  # shells 0..n: even shells 2j -> (j+1)^2, odd shells 2j+1 -> (j+1)(j+2)
    even = (k + 1) * (k + 2) * (2 * k + 3) // 6
    jo = (n - 1) // 2 if n >= 1 else -1  # last j with 2j+1 <= n
    odd = (jo + 1) * (jo + 2) * (jo + 3) // 3 if jo >= 0 else 0
    return even + odd


# ----------------------------------------------------------------------------- radial (n, l)
def mu_R(n):
    """Last 0-based radial index of order ``n`` (ix:13-28): 675 at n = 50."""
    return num_radial(int(n)) - 1


def nl_to_mu_R(moment):
    n, l = int(moment[0]), int(moment[1])
    return num_radial(n - 1) + (l - n % 2) // 2


def mu_R_to_nl(mu):
    mu = int(mu)
    n = max(0, int(2.0 * math.sqrt(mu + 1.0)) - 3)
    while num_radial(n) <= mu:
        n += 1
    return (n, n % 2 + 2 * (mu - num_radial(n - 1)))


# ----------------------------------------------------------------------------- spherical (l, m)
def mu_Y(l):
    """Number of spherical harmonics (m >= 0) up to degree ``l`` (ix:87-99)."""
    return num_harmonics(int(l))


def lm_to_mu_Y(moment):
    l, m = int(moment[0]), int(moment[1])
    return l * (l + 1) // 2 + m


def mu_Y_to_lm(mu):
    mu = int(mu)
    l = (math.isqrt(8 * mu + 1) - 1) // 2
    return (l, mu - l * (l + 1) // 2)


# ----------------------------------------------------------------------------- Zernike (n, l, m)
def mu_(n):
    """Last 0-based moment index of order ``n`` (ix:140-162): ``mu_(50) + 1 == 12051``."""
    return num_moments(int(n)) - 1


def nlm_to_mu(moment):
    n, l, m = int(moment[0]), int(moment[1]), abs(int(moment[2]))
    j = (l - n % 2) // 2  # l is the j-th degree of shell n
    before = j * j if n % 2 == 0 else j * (j + 1)  # sum of (l' + 1) over lower degrees of the shell
    return num_moments(n - 1) + before + m


def mu_to_nlm(mu):
    mu = int(mu)
    n = max(0, int(round((6.0 * (mu + 1)) ** (1.0 / 3.0))) - 3)
    while num_moments(n) <= mu:
        n += 1
    r = mu - num_moments(n - 1)
    if n % 2 == 0:
        j = math.isqrt(r)  # j^2 <= r
    else:
        j = (math.isqrt(4 * r + 1) - 1) // 2  # j(j+1) <= r
    before = j * j if n % 2 == 0 else j * (j + 1)
    return (n, n % 2 + 2 * j, r - before)


def memory_requirement(n, dim):
    """GB the *reference* needs for its materialised basis (ix:230-234) - kept for comparison only;
    this implementation never builds those arrays."""
    return (mu_Y(n) * (dim ** 3) * 8 + mu_R(n) * (dim ** 3) * 4) / 1e9


# ----------------------------------------------------------------------------- vectorised tables
def nlm_table(order: int) -> np.ndarray:
    """int32 ``[num_moments, 3]`` of (n, l, m) per flat index."""
    out = np.empty((num_moments(order), 3), dtype=np.int32)
    mu = 0
    for n in range(order + 1):
        for l in range(n % 2, n + 1, 2):
            out[mu:mu + l + 1, 0] = n
            out[mu:mu + l + 1, 1] = l
            out[mu:mu + l + 1, 2] = np.arange(l + 1)
            mu += l + 1
    return out


def nl_table(order: int) -> np.ndarray:
    """int32 ``[num_radial, 3]`` of (n, l, first flat moment index) per radial index."""
    rows, mu = [], 0
    for n in range(order + 1):
        for l in range(n % 2, n + 1, 2):
            rows.append((n, l, mu))
            mu += l + 1
    return np.array(rows, dtype=np.int32)
