"""Seeded synthetic inputs shared by the oracle, the parity tests and bench.py.

The reference ships no sample data, so every parity input is generated here.
``volume(seed, n)`` is the generator SURVEY.md §8(d) specifies: 12 isotropic Gaussian blobs plus
1 % white noise on an ``[n, n, n]`` float32 grid indexed exactly like ``mrc.data`` in the
reference's ``run()`` (zernike_master.py:2851-2852).
"""
from __future__ import annotations

import numpy as np


def volume(seed: int, n: int) -> np.ndarray:
    """float32 ``[n, n, n]`` test volume: blobs in a [-1, 1]^3 frame + 0.01 * N(0, 1)."""
    rng = np.random.default_rng(seed)
    ax = np.linspace(-1.0, 1.0, n)
    g0, g1, g2 = np.meshgrid(ax, ax, ax, indexing="ij", sparse=True)
    vol = np.zeros((n, n, n), dtype=np.float64)
    for _ in range(12):
        c = rng.uniform(-0.5, 0.5, size=3)
        sigma = rng.uniform(0.05, 0.2)
        amp = rng.uniform(0.5, 1.5)
        d2 = (g0 - c[0]) ** 2 + (g1 - c[1]) ** 2 + (g2 - c[2]) ** 2
        vol += amp * np.exp(-0.5 * d2 / sigma**2)
    vol += 0.01 * rng.standard_normal((n, n, n))
    return vol.astype(np.float32)


def bundle_3dva(n: int, components: int = 5, particles: int = 1000, seed: int = 2):
    """A cryoSPARC-3DVA-like bundle as read by ``_read_wiggle`` (zernike_master.py:2838-2842).

    Returns ``(particles, consensus_map, components)`` where ``particles`` is a structured array
    holding the ``components_mode_<i>/value`` float32 fields ``Moments.set_latents`` pulls
    (zernike_master.py:201-206).
    """
    consensus = volume(seed, n)
    comps = []
    for i in range(components):
        c = 0.1 * volume(seed + 1 + i, n)
        comps.append((c - c.mean()).astype(np.float32))
    comps = np.stack(comps)
    rng = np.random.default_rng(seed + 1000)
    dt = np.dtype([("uid", "<u8")] + [(f"components_mode_{i}/value", "<f4") for i in range(components)])
    part = np.zeros(particles, dtype=dt)
    part["uid"] = np.arange(particles)
    for i in range(components):
        part[f"components_mode_{i}/value"] = rng.normal(0.0, 30.0, size=particles).astype(np.float32)
    return part, consensus, comps
