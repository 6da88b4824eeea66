"""Minimal MRC2014 reader / writer in numpy.

The reference reads its input with ``mrcfile.open(path).data`` / ``.voxel_size`` (zernike_master.py:
2851-2853) and writes reconstructions with ``mrc.set_data(float32)`` + ``mrc.voxel_size = apix``
(zm:2717-2721).  ``mrcfile`` is not available in this image, so this module implements just that
surface: a 1024-byte MRC2014 header, data indexed ``[z][y][x]`` (sections, rows, columns), voxel size
= cell / sampling.
"""
from __future__ import annotations

import contextlib
import io

import numpy as np

_MODES = {0: np.int8, 1: np.int16, 2: np.float32, 6: np.uint16, 12: np.float16}
_MODE_OF = {np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.float32): 2, np.dtype(np.uint16): 6,
            np.dtype(np.float16): 12}


class VoxelSize:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __getitem__(self, k):
        return getattr(self, k)

    def __repr__(self):
        return f"VoxelSize(x={self.x}, y={self.y}, z={self.z})"


class MrcFile:
    """The slice of ``mrcfile.MrcFile`` the reference uses: ``.data``, ``.voxel_size``, ``set_data``."""

    def __init__(self, data=None, voxel_size=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
        self.data = None if data is None else np.asarray(data)
        if np.isscalar(voxel_size):
            voxel_size = (voxel_size,) * 3
        self._voxel = VoxelSize(*voxel_size)
        self.origin = tuple(float(o) for o in origin)

    @property
    def voxel_size(self):
        return self._voxel

    @voxel_size.setter
    def voxel_size(self, v):
        if np.isscalar(v):
            v = (v, v, v)
        self._voxel = VoxelSize(*v)

    def set_data(self, data):
        self.data = np.ascontiguousarray(data)


def read(path: str) -> MrcFile:
    with io.open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 1024:
        raise ValueError(f"{path}: shorter than an MRC header")
    stamp = raw[212:214]
    order = ">" if stamp[:1] == b"\x11" else "<"
    i4 = np.frombuffer(raw[:1024], dtype=order + "i4")
    f4 = np.frombuffer(raw[:1024], dtype=order + "f4")
    nx, ny, nz, mode = (int(v) for v in i4[0:4])
    mx, my, mz = (int(v) for v in i4[7:10])
    cella = f4[10:13]
    nsymbt = int(i4[23])
    if mode not in _MODES:
        raise ValueError(f"{path}: unsupported MRC mode {mode}")
    if raw[208:211] != b"MAP":
        raise ValueError(f"{path}: missing 'MAP' stamp")
    dt = np.dtype(_MODES[mode]).newbyteorder(order)
    count = nx * ny * nz
    data = np.frombuffer(raw, dtype=dt, count=count, offset=1024 + nsymbt).reshape(nz, ny, nx)
    data = data.astype(dt.newbyteorder("="), copy=False)
    vox = tuple(float(c) / m if m > 0 else 0.0 for c, m in zip(cella, (mx, my, mz)))
    return MrcFile(data, vox, origin=tuple(float(v) for v in f4[49:52]))


def write(path: str, data, voxel_size=1.0) -> None:
    data = np.ascontiguousarray(data)
    if data.dtype not in _MODE_OF:
        data = data.astype(np.float32)
    if data.ndim != 3:
        raise ValueError("only 3-D volumes are supported")
    nz, ny, nx = data.shape
    if np.isscalar(voxel_size):
        voxel_size = (voxel_size,) * 3
    elif isinstance(voxel_size, VoxelSize):
        voxel_size = (voxel_size.x, voxel_size.y, voxel_size.z)
    hdr = np.zeros(256, dtype="<i4")
    f = hdr.view("<f4")
    hdr[0:3] = (nx, ny, nz)
    hdr[3] = _MODE_OF[data.dtype]
    hdr[7:10] = (nx, ny, nz)                      # mx, my, mz
    f[10:13] = (nx * voxel_size[0], ny * voxel_size[1], nz * voxel_size[2])  # cella
    f[13:16] = 90.0                               # cellb
    hdr[16:19] = (1, 2, 3)                        # mapc, mapr, maps
    d64 = data.astype(np.float64, copy=False)
    f[19], f[20], f[21] = d64.min(), d64.max(), d64.mean()
    hdr[22] = 1                                   # ispg: volume
    hdr[27] = 20141                               # nversion
    raw = bytearray(hdr.tobytes())
    raw[208:212] = b"MAP "
    raw[212:216] = b"\x44\x44\x00\x00"            # little-endian machine stamp
    f2 = np.frombuffer(raw, dtype="<f4").copy()
    f2[54] = d64.std()                            # rms
    raw[216:220] = f2[54:55].tobytes()
    with io.open(path, "wb") as fh:
        fh.write(bytes(raw))
        fh.write(data.astype(data.dtype.newbyteorder("<"), copy=False).tobytes())


@contextlib.contextmanager
def open(path: str, mode: str = "r"):  # noqa: A001 - mirrors mrcfile.open
    """``with mrc.open(p) as m: m.data`` / ``with mrc.open(p, mode='w+') as m: m.set_data(v); m.voxel_size = a``."""
    if mode == "r":
        yield read(path)
    elif mode in ("w+", "w"):
        m = MrcFile()
        yield m
        if m.data is None:
            raise ValueError("no data set on the new MRC file")
        write(path, m.data, m.voxel_size)
    else:
        raise ValueError(f"unsupported mode {mode!r}")
