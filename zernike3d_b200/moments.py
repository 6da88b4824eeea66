"""``Moments``: container + ``.npz`` wire format of the reference (zernike_master.py:26-382), unchanged.

Layout on disk (zm:334-360, README "f.keys"):
  single volume : ``dim``, ``apix``, ``arr_0`` = complex64[num_moments] in sorted (n, l, m), m >= 0
  3DVA bundle   : ``latents`` float32[P, K], ``dim``, ``apix``, ``arr_0`` (base map), ``arr_1..arr_K``
``load`` sniffs the layout by key names / count exactly like zm:271-296.  The public attributes
(``moments``, ``ndmoments``, ``ndinvariants``, ``latents``, ``scaled_moments`` ...) are the dicts the
reference exposes, so downstream code such as ``moment_analysis_3dva.py`` keeps working; array views
(``as_array`` / ``from_array``) are what the GPU path uses.
"""
from __future__ import annotations

import sys

import numpy as np
import numpy.lib.recfunctions as rf

from .indexing import mu_to_nlm, nlm_table, num_moments


class Moments:
    def __init__(self, quiet: bool = False):
        if not quiet:
            print("Created `Moments` instance")
        self.order = None
        self.voxdim = 64
        self.dimensions = None
        self.apix = 1
        self.moments = {}
        self.invariants = {}
        self.ndmoments = {}
        self.ndinvariants = {}
        self.scaled_moments = {}
        self.latents = None

    # ------------------------------------------------------------------ array <-> dict bridges
    @staticmethod
    def _keys(order: int):
        return [tuple(int(v) for v in row) for row in nlm_table(order)]

    def from_array(self, arr, order: int | None = None) -> None:
        """Fill ``self.moments`` from a flat complex64 vector in the reference's order."""
        arr = np.asarray(arr, dtype=np.complex64)
        if order is None:
            order = self.order_of_length(arr.shape[0])
        keys = self._keys(order)
        if len(keys) != arr.shape[0]:
            raise ValueError(f"{arr.shape[0]} moments do not form a complete order-{order} set ({len(keys)})")
        self.moments = dict(zip(keys, arr))

    @staticmethod
    def order_of_length(count: int) -> int:
        n = 0
        while num_moments(n) < count:
            n += 1
        if num_moments(n) != count:
            raise ValueError(f"{count} is not the size of a complete moment set")
        return n

    def as_array(self, latent_dim: int | None = None) -> np.ndarray:
        """complex64 vector (sorted keys, zm:378-382) of one stored volume (or of ``self.moments``)."""
        src = self.moments if latent_dim is None else self.ndmoments[latent_dim]
        return self._save_moments(src)

    def as_matrix(self) -> np.ndarray:
        """complex64 ``[dimensions, num_moments]`` of all stored volumes."""
        return np.stack([self._save_moments(m) for _, m in sorted(self.ndmoments.items())])

    def _ndmoment_to_ndarray(self):
        o = self.order
        arr = np.zeros((self.dimensions, o + 1, o + 1, o + 1), dtype=np.complex64)
        mask = np.zeros((o + 1, o + 1, o + 1), dtype=bool)
        for dim, moments in self.ndmoments.items():
            for (n, l, m), mom in moments.items():
                arr[dim, n, l, m] = mom
                mask[n, l, m] = True
        return arr, mask

    # ------------------------------------------------------------------ reference API
    def add_dimension(self):
        self.ndmoments[len(self.ndmoments)] = self.moments
        self.moments = {}

    def calculate_invariants(self):
        """F_nl = l2 norm of (|O_nl0|, |O_nlm|, |O_nlm| for m > 0)   (zm:62-76)."""
        for dimension, moments in self.ndmoments.items():
            acc = {}
            for (n, l, m), mom in moments.items():
                a2 = float(np.abs(mom)) ** 2
                acc[(n, l)] = acc.get((n, l), 0.0) + (a2 if m == 0 else 2.0 * a2)
            self.ndinvariants[dimension] = {nl: np.float32(np.sqrt(v)) for nl, v in acc.items()}
        self.invariants = {}

    def zcc(self, other, dimension=0, other_dimension=0):
        """Per-order correlation curve against another ``Moments`` ("Zernike FSC", ``Z_3DVA.calc_ZCC``
        moment_analysis_3dva.py:402-422): c_n = sum |A conj(B)| / sqrt(sum |A|^2 sum |B|^2) over the stored moments of order n.
        The device version for batches is ``Plan.zcc`` (``z3d_zcc``)."""
        a, b = self.ndmoments[dimension], other.ndmoments[other_dimension]
        order = max(n for (n, _, _) in a)
        nom, d1, d2 = np.zeros(order + 1), np.zeros(order + 1), np.zeros(order + 1)
        for (n, l, m), va in a.items():
            vb = np.conjugate(b[(n, l, m)])
            nom[n] += np.abs(va * vb)
            d1[n] += np.abs(va) ** 2
            d2[n] += np.abs(vb) ** 2
        with np.errstate(invalid="ignore", divide="ignore"):
            return (nom / np.sqrt(d1 * d2)).astype(np.float32)

    def apply_scaling(self, z_ind=None, scales=None):
        """Omega = [1, latent...] . Omega_dims   (zm:78-95)."""
        mat = self.as_matrix()  # [K+1, num_moments]
        if z_ind is None and scales is None:
            print("No latent coordinates (`scales=`) and no Z-index (`z_ind=`) were provided, "
                  "so returning a random latent coordinate.")
            z_ind = int(np.random.choice(len(self.latents)))
        if z_ind is not None:
            w = np.concatenate([[1.0], np.asarray(self.latents[z_ind], dtype=np.float32)]).astype(np.float32)
        else:  # --scales arrive as strings from argparse (zm:3133-3138)
            w = np.array([1.0] + [float(s) for s in scales], dtype=np.float32)
        if w.shape[0] != mat.shape[0]:
            raise ValueError(f"{w.shape[0] - 1} scale factors given, the moment file holds {mat.shape[0] - 1} components")
        scaled = np.tensordot(w, mat, axes=([0], [0])).astype(np.complex64)
        self.scaled_moments = dict(zip(self._keys(self.order), scaled))
        return scaled

    def get_moment_size(self):
        return len(self.ndmoments[0]) if self.ndmoments else len(self.moments)

    def get_order(self):
        src = self.ndmoments[0] if self.ndmoments else self.moments
        return max(src)[0]

    def get_size(self):
        if self.ndmoments:
            return sum(self._size(v) for v in self.ndmoments.values())
        return self._size(self.moments)

    def _size(self, obj):
        size = sys.getsizeof(obj)
        if isinstance(obj, dict):
            size += sum(getattr(v, "nbytes", sys.getsizeof(v)) for v in obj.values())
            size += sum(sys.getsizeof(k) for k in obj)
        return size

    def stats(self, quiet=False):
        self.dimensions = len(self.ndmoments)
        initialised = len(self.ndinvariants) != 0
        moment_size = self.get_moment_size()
        self.order = self.get_order()
        if quiet:
            return
        first = self.ndmoments[0][(0, 0, 0)] if self.ndmoments else self.moments[(0, 0, 0)]
        lines = ["Moment statistics", f"   Dimensions: {self.dimensions}", f"   Moments: {moment_size}",
                 f"   Order: {self.order}", f"   nbytes: {self.get_size()}", f"   type: {np.asarray(first).dtype}", "",
                 f"   Invariants: {initialised}"]
        if initialised:
            lines += [f"   Total: {len(self.ndinvariants[0])}",
                      f"   nbytes: {sum(self._size(v) for v in self.ndinvariants.values())}",
                      f"   type: {self.ndinvariants[0][(0, 0)].dtype}"]
        if self.latents is not None:
            lines.append(f"   particles: {self.latents.shape}")
        bar = "+" + "-" * 55 + "+"
        print(bar)
        for ln in lines:
            print("|{:<55}|".format(ln))
        print(bar)

    def check_allowed(self, n, l, m):
        if n > self.order:
            print(f"n = {n} not allowed, required: n = {self.order}")
        elif l > n:
            print(f"l = {l} not allowed, required: l < n or l = n")
        elif (n - l) % 2:
            print(f"l = {l} not allowed, n - l must be even")
        elif not -l <= m <= l:
            print(f"m = {m} not allowed, required: -l <= m <= l")
        else:
            return True
        return False

    def get_moment(self, n, l, m, latent_dim=0):
        if self.check_allowed(n, l, m):
            return self.ndmoments[latent_dim][(n, l, m)] if self.ndmoments else self.moments[(n, l, m)]

    def get_invariant(self, n, l, latent_dim=0):
        if self.check_allowed(n, l, m=0):
            return self.ndinvariants[latent_dim][(n, l)] if self.ndinvariants else self.invariants[(n, l)]

    def set_moment(self, n, l, m, moment):
        self.moments[(n, l, m)] = moment

    def set_latents(self, particles, components):
        """Pull ``components_mode_<i>/value`` out of a cryoSPARC particle array (zm:201-206)."""
        keys = [f"components_mode_{c}/value" for c in range(components)]
        sel = rf.repack_fields(particles[keys])
        self.latents = sel.copy().view("<f4").reshape((sel.shape[0], components))

    def display_moments(self, latent_dim=0, nlm=None, n_max=None):
        if nlm is not None:
            n, l, m = nlm
            if m < 0:  # Omega_{n,l,-m} = (-1)^m conj(Omega_{n,l,m})
                print(((-1) ** m) * np.conj(self.get_moment(n=n, l=l, m=abs(m), latent_dim=latent_dim)))
            else:
                print(self.get_moment(n=n, l=l, m=m, latent_dim=latent_dim))
            return
        if self.ndmoments and latent_dim not in self.ndmoments:
            print(f"Key error: There are only {self.dimensions} dimensions, you gave dim={latent_dim}")
            return
        src = self.ndmoments[latent_dim] if self.ndmoments else self.moments
        for key, val in src.items():
            if n_max is not None and key[0] > n_max:
                break
            print(key, val)

    def display_invariants(self, latent_dim=0, nl=None, n_max=None):
        if nl is not None:
            print(self.get_invariant(n=nl[0], l=nl[1], latent_dim=latent_dim))
            return
        if self.ndinvariants and latent_dim not in self.ndinvariants:
            print(f"Key error: There are only {self.dimensions} dimensions, you gave dim={latent_dim}")
            return
        src = self.ndinvariants[latent_dim] if self.ndinvariants else self.invariants
        for key, val in src.items():
            if n_max is not None and key[0] > n_max:
                break
            print(key, val)

    # ------------------------------------------------------------------ npz I/O
    def load(self, input):
        print(f"Loading {input} into Moments object")
        with np.load(input) as data:
            files = list(data.files)
            if "arr_0" in files and len(files) == 3:
                print("Detected full precision moments (single).")
                self.voxdim, self.apix = data["dim"], data["apix"]
                self.from_array(data["arr_0"])
                self.add_dimension()
            elif "arr_0" in files and len(files) > 3:
                print("Detected full precision moments (multiple).")
                self.voxdim, self.apix, self.latents = data["dim"], data["apix"], data["latents"]
                arrs = [k for k in files if k not in ("dim", "apix", "latents")]
                for key in arrs:  # arr_0, arr_1, ... in file order, like the reference
                    self.from_array(data[key])
                    self.add_dimension()
            elif "args_0" in files:
                many = len(files) > 5
                print(f"Detected 8 bit precision moments ({'multiple' if many else 'single'}).")
                self.voxdim, self.apix = data["dim"], data["apix"]
                if many:
                    self.latents = data["latents"]
                k = 0
                while f"args_{k}" in files:
                    self.from_array(self._decode_8bit(data[f"args_{k}"], data[f"phases_{k}"], data[f"deltas_{k}"]))
                    self.add_dimension()
                    k += 1
            else:
                raise ValueError(f"{input}: not a zernike3d moment file (keys {files})")
        self.stats(quiet=True)
        return len(self.ndmoments)

    def save(self, output="moments", compress=False, dim=64, apix=1):
        sets = [m for _, m in sorted(self.ndmoments.items())] if self.ndmoments else [self.moments]
        payload = {}
        for i, moments in enumerate(sets):
            if compress:
                a, p, d = self._save_compressed_moments(moments)
                payload[f"args_{i}"], payload[f"phases_{i}"], payload[f"deltas_{i}"] = a, p, d
            else:
                payload[f"arr_{i}"] = self._save_moments(moments)
        if compress:
            output = str(output).split(".")[0] + ".8bit"
        if self.ndmoments:
            np.savez_compressed(output, latents=self.latents, dim=dim, apix=apix, **payload)
        else:
            np.savez_compressed(output, dim=dim, apix=apix, **payload)

    @staticmethod
    def _decode_8bit(args, phases, deltas):
        r_vals = np.geomspace(deltas[0], deltas[1], num=256)
        r = r_vals[args]
        theta = phases * deltas[2]
        return (r * (np.cos(theta) + 1j * np.sin(theta))).astype(np.complex64)

    def _save_compressed_moments(self, moments):
        """Lossy polar 2 x uint8 encoding (zm:362-376) - kept for the ``--save_bit8`` flag only."""
        z = self._save_moments(moments)
        mag, ph = np.abs(z), np.angle(z)
        lo, hi = float(mag.min()), float(mag.max())
        lo = 1e-100 if lo == 0 else lo
        dtheta = 2 * np.pi / 256
        a8 = np.digitize(mag, np.geomspace(lo, hi, num=256), right=False).astype(np.uint8)
        p8 = np.digitize(ph, np.arange(-np.pi, np.pi, dtheta), right=False).astype(np.uint8)
        return a8, p8, np.array([lo, hi, dtheta], dtype=np.float64)

    @staticmethod
    def _save_moments(moments):
        return np.array([moments[k] for k in sorted(moments)]).astype(np.complex64)
