"""``Plan``: thin host object around one ``z3d_plan`` (order, box size, device).

It replaces what ``ZernikeSpherical.CalculatePolynomials`` builds in the reference (zernike_master.py:
2237-2323) - but holds only recurrence-coefficient and tile tables (a few hundred KB), never an
R_nl / Y_lm voxel grid.  PyTorch owns the device buffers and streams; all arithmetic happens in the
hand-written sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


class Plan:
    INFO_KEYS = ("order", "dim", "num_moments", "passes", "ctas", "threads", "tiles_per_thread", "smem_bytes",
                 "voxels_per_chunk", "sms", "useful_accumulators", "padded_accumulators",
                 "batched_min", "batched_row_groups", "batched_row_groups_wide", "batched_wide_min", "batched_rounds",
                 "batched_rounds_wide", "stream_state", "stream_sets", "stream_splits", "stream_rows16", "stream_table_mib",
                 "chunks", "gen", "gen_sets", "gen_splits", "gen_cols", "gen_slots", "gen_min", "tables_level", "gen_maxcols",
                 "gen2", "gen2_sets", "gen2_splits", "gen2_slots", "synth", "synth_min", "synth_partial_slots")

    def __init__(self, order: int, dim: int, device: int | None = None):
        if not torch.cuda.is_available():
            raise _lib.Z3DError(_lib.ERR_NODEVICE, "no CUDA device: zernike3d_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._L = _lib.lib()
        h = ctypes.c_void_p()
        _lib.check(self._L.z3d_plan_create(int(order), int(dim), self.device, ctypes.byref(h)))
        self._h = h
        self.order, self.dim = int(order), int(dim)
        self.num_moments = self._L.z3d_num_moments(self.order)
        self.num_radial = self._L.z3d_num_radial(self.order)
        self.half = (self.dim + 1) // 2
        self._ws = None
        self._ws_synth = None
        self._info = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._L.z3d_plan_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------ helpers
    def info(self) -> dict:
        arr = (ctypes.c_int * 39)()
        _lib.check(self._L.z3d_plan_info(self._h, arr, 39))
        return dict(zip(self.INFO_KEYS, list(arr)))

    def prepare_tables(self, level: int, max_bytes: int = 0) -> int:
        """Opt in to the table-fed batched kernels (``z3d_plan_prepare_tables``): 1 = R / Q factor tables, 2 = + the streamed T
        table, 0 = back to the table-free default.  Returns the level reached."""
        got = ctypes.c_int()
        _lib.check(self._L.z3d_plan_prepare_tables(self._h, int(level), int(max_bytes), ctypes.byref(got)))
        self._ws = None
        return int(got.value)

    def workspace_bytes(self, batch: int) -> int:
        return int(self._L.z3d_workspace_bytes(self._h, int(batch)))

    def _workspace(self, batch: int, stream=None) -> torch.Tensor:
        """Scratch owned by the plan.  It is allocated on torch's current stream of the plan's device; when a call runs on
        another stream the caching allocator is told (``record_stream``), so that neither a later growth (which drops the old
        buffer) nor the allocator's reuse can hand the memory out while that stream still works on it."""
        need = self.workspace_bytes(batch)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=f"cuda:{self.device}")
        if stream is not None:
            self._ws.record_stream(stream)
        return self._ws

    def _stream(self, stream) -> ctypes.c_void_p:
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    def _check_moments(self, m: torch.Tensor, what: str) -> torch.Tensor:
        if m.dtype != torch.complex64 or not m.is_cuda or m.dim() != 2 or m.shape[1] != self.num_moments or m.device.index != self.device:
            raise ValueError(f"{what}: expected complex64 [B,{self.num_moments}] on cuda:{self.device}, got {m.dtype} "
                             f"{tuple(m.shape)} on {m.device}")
        return m.contiguous()

    def _check_volumes(self, v: torch.Tensor, what: str) -> torch.Tensor:
        N = self.dim
        if v.dtype != torch.float32 or not v.is_cuda or tuple(v.shape[1:]) != (N, N, N) or v.device.index != self.device:
            raise ValueError(f"{what}: expected a float32 tensor [B,{N},{N},{N}] on cuda:{self.device}, got {v.dtype} "
                             f"{tuple(v.shape)} on {v.device}")
        return v.contiguous()

    def _pairs(self, pairs):
        lo, hi = (0, self.half) if pairs is None else pairs
        return int(lo), int(hi)

    def profile(self, on: bool = True) -> None:
        _lib.check(self._L.z3d_profile_enable(self._h, int(on)))

    def last_kernel_ms(self) -> float:
        """Duration of the dominant kernel (k_decompose / k_reconstruct) of the last call, CUDA events on the
        launching stream; requires ``profile(True)``."""
        ms = ctypes.c_float()
        _lib.check(self._L.z3d_profile_last_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    # ------------------------------------------------------------------ device-buffer API
    def decompose(self, vol: torch.Tensor, pairs=None, out: torch.Tensor | None = None, stream=None,
                  part=None) -> torch.Tensor:
        """``vol`` float32 CUDA ``[B, N, N, N]`` (or ``[N, N, N]``) -> complex64 ``[B, num_moments]``.

        ``pairs=(lo, hi)`` restricts the voxel sum to the mirrored plane pairs lo..hi-1 of axis 0 (slab partial: only
        those planes are read); ``part=(i, n)`` instead sums the i-th of n equal ranges of the full-volume work list
        (balanced multi-GPU sharding of one replicated volume); the default is the whole volume."""
        single = vol.dim() == 3
        v = vol.unsqueeze(0) if single else vol
        v = self._check_volumes(v, "decompose")
        B = v.shape[0]
        if out is None:
            out = torch.empty((B, self.num_moments), dtype=torch.complex64, device=v.device)
        elif stream is not None:
            out.record_stream(stream)
        ws = self._workspace(B, stream)
        if part is not None:
            if pairs is not None:
                raise ValueError("give either pairs= or part=, not both")
            _lib.check(self._L.z3d_decompose_part(self._h, _ptr(v), B, int(part[0]), int(part[1]), _ptr(out), _ptr(ws),
                                                  ws.numel(), self._stream(stream)))
        else:
            lo, hi = self._pairs(pairs)
            _lib.check(self._L.z3d_decompose(self._h, _ptr(v), B, lo, hi, _ptr(out), _ptr(ws), ws.numel(),
                                             self._stream(stream)))
        return out[0] if single else out

    def decompose_allreduce(self, vol: torch.Tensor, comm: "PeerComm", out: torch.Tensor | None = None, stream=None):
        """Work-sharded analysis of replicated volume(s) fused with a one-shot all-reduce over NVLink peer memory
        (``z3d_decompose_allreduce``): every rank returns the full, bit-identical moments."""
        single = vol.dim() == 3
        v = self._check_volumes(vol.unsqueeze(0) if single else vol, "decompose_allreduce")
        B = v.shape[0]
        if out is None:
            out = torch.empty((B, self.num_moments), dtype=torch.complex64, device=v.device)
        ws = self._workspace(B, stream)
        _lib.check(self._L.z3d_decompose_allreduce(self._h, comm._c, _ptr(v), B, _ptr(out), _ptr(ws), ws.numel(),
                                                   self._stream(stream)))
        return out[0] if single else out

    def reconstruct(self, moments: torch.Tensor, pairs=None, out: torch.Tensor | None = None, stream=None,
                    batched: bool | None = None) -> torch.Tensor:
        """complex64 CUDA ``[B, num_moments]`` -> float32 ``[B, N, N, N]`` (planes outside ``pairs`` untouched).
        ``batched``: None = whole maps of a batch of at least ``info()["synth_min"]`` go through the tensor-core synthesis kernel
        in tiles of 64 (``z3d_reconstruct_batched``), everything else through the per-map kernel; True / False force one of them."""
        single = moments.dim() == 1
        m = moments.unsqueeze(0) if single else moments
        m = self._check_moments(m, "reconstruct")
        B, N = m.shape[0], self.dim
        lo, hi = self._pairs(pairs)
        if out is not None and (out.dtype != torch.float32 or not out.is_cuda or out.device.index != self.device or not out.is_contiguous()
                                or tuple(out.shape) != (B, N, N, N)):
            raise ValueError(f"reconstruct: out must be a contiguous float32 tensor [{B},{N},{N},{N}] on cuda:{self.device}")
        if self._info is None:
            self._info = self.info()
        if batched is None:   # whole maps of a large enough batch: tiles of 64 on the tensor cores (z3d_reconstruct_batched)
            batched = bool(self._info["synth"]) and B >= self._info["synth_min"] and (lo, hi) == (0, self.half)
        if batched:
            if (lo, hi) != (0, self.half):
                raise ValueError("reconstruct: the tensor-core path synthesises whole maps (pairs must cover the box)")
            if out is None:
                out = torch.empty((B, N, N, N), dtype=torch.float32, device=m.device)
            need = int(self._L.z3d_reconstruct_batched_workspace_bytes(self._h, B))
            if need == 0:
                raise _lib.Z3DError(_lib.ERR_INVALID, "reconstruct: no tensor-core synthesis layout for this plan")
            if self._ws_synth is None or self._ws_synth.numel() < need:
                self._ws_synth = None
                self._ws_synth = torch.empty(need, dtype=torch.uint8, device=f"cuda:{self.device}")
            if stream is not None:
                self._ws_synth.record_stream(stream)
                out.record_stream(stream)
            _lib.check(self._L.z3d_reconstruct_batched(self._h, _ptr(m), B, _ptr(out), _ptr(self._ws_synth), self._ws_synth.numel(),
                                                       self._stream(stream)))
            return out[0] if single else out
        if out is None:
            out = torch.zeros((B, N, N, N), dtype=torch.float32, device=m.device)
        ws = self._workspace(B, stream)
        _lib.check(self._L.z3d_reconstruct(self._h, _ptr(m), B, lo, hi, _ptr(out), _ptr(ws), ws.numel(),
                                           self._stream(stream)))
        return out[0] if single else out

    def scale_moments(self, dims: torch.Tensor, latents: torch.Tensor, stream=None) -> torch.Tensor:
        """3DVA latent scaling on the device (``Moments.apply_scaling``, zm:78-95, for many coordinates at once):
        ``dims`` complex64 CUDA ``[K + 1, num_moments]`` (base map, components), ``latents`` float32 CUDA ``[P, K]`` ->
        complex64 ``[P, num_moments]`` = base + latents @ components, ready for :meth:`reconstruct`."""
        if dims.dtype != torch.complex64 or not dims.is_cuda or dims.dim() != 2 or dims.shape[1] != self.num_moments:
            raise ValueError(f"expected complex64 CUDA [K+1,{self.num_moments}], got {dims.dtype} {tuple(dims.shape)}")
        K = dims.shape[0] - 1
        lat = latents.reshape(-1, K) if K else latents.reshape(latents.shape[0], 0)
        if lat.dtype != torch.float32 or not lat.is_cuda:
            raise ValueError("latents must be a float32 CUDA tensor [P, K]")
        dims, lat = dims.contiguous(), lat.contiguous()
        out = torch.empty((lat.shape[0], self.num_moments), dtype=torch.complex64, device=dims.device)
        _lib.check(self._L.z3d_scale_moments(self._h, _ptr(dims), K, _ptr(lat), lat.shape[0], _ptr(out), self._stream(stream)))
        return out

    def invariants(self, moments: torch.Tensor, stream=None) -> torch.Tensor:
        single = moments.dim() == 1
        m = self._check_moments(moments.unsqueeze(0) if single else moments, "invariants")
        out = torch.empty((m.shape[0], self.num_radial), dtype=torch.float32, device=m.device)
        _lib.check(self._L.z3d_invariants(self._h, _ptr(m), m.shape[0], _ptr(out), self._stream(stream)))
        return out[0] if single else out

    def zcc(self, moments_a: torch.Tensor, moments_b: torch.Tensor, stream=None) -> torch.Tensor:
        """Per-order correlation curve(s) of two moment sets on the device (``Z_3DVA.calc_ZCC``, moment_analysis_3dva.py:402-422):
        ``[num_moments]`` or ``[B, num_moments]`` complex64 each -> ``[order + 1]`` or ``[B, order + 1]`` float32."""
        single = moments_a.dim() == 1
        a = self._check_moments(moments_a.unsqueeze(0) if single else moments_a, "zcc")
        b = self._check_moments(moments_b.unsqueeze(0) if single else moments_b, "zcc")
        if a.shape != b.shape:
            raise ValueError("zcc: moment sets of different shapes")
        out = torch.empty((a.shape[0], self.order + 1), dtype=torch.float32, device=a.device)
        _lib.check(self._L.z3d_zcc(self._h, _ptr(a), _ptr(b), a.shape[0], _ptr(out), self._stream(stream)))
        return out[0] if single else out

    # ------------------------------------------------------------------ host-buffer API
    def decompose_host(self, vol, out=None):
        """numpy / pinned-tensor volumes ``[B, N, N, N]`` float32 -> numpy complex64 ``[B, num_moments]``.
        H2D, kernels and D2H are pipelined inside the library (``z3d_decompose_host``)."""
        v = vol.numpy() if isinstance(vol, torch.Tensor) else np.asarray(vol)
        single = v.ndim == 3
        v = np.ascontiguousarray(v[None] if single else v, dtype=np.float32)
        B = v.shape[0]
        if out is None:
            out = np.empty((B, self.num_moments), dtype=np.complex64)
        o = out.numpy() if isinstance(out, torch.Tensor) else out
        _lib.check(self._L.z3d_decompose_host(self._h, ctypes.c_void_p(v.ctypes.data), B,
                                              ctypes.c_void_p(o.ctypes.data)))
        return out[0] if single else out

    def decompose_host_submit(self, vol, out) -> None:
        """Streaming form of :meth:`decompose_host` (``z3d_decompose_host_submit``): enqueue one batch and return; ``vol`` (float32
        ``[B, N, N, N]``) and ``out`` (complex64 ``[B, num_moments]``) must be contiguous pinned tensors / arrays that stay alive and
        untouched until :meth:`host_wait`.  The next batch's copies overlap this batch's last kernels."""
        v = vol.numpy() if isinstance(vol, torch.Tensor) else vol
        o = out.numpy() if isinstance(out, torch.Tensor) else out
        if v.dtype != np.float32 or not v.flags.c_contiguous or v.ndim != 4 or tuple(v.shape[1:]) != (self.dim,) * 3:
            raise ValueError(f"decompose_host_submit: need contiguous float32 volumes [B,{self.dim},{self.dim},{self.dim}]")
        if o.dtype != np.complex64 or not o.flags.c_contiguous or tuple(o.shape) != (v.shape[0], self.num_moments):
            raise ValueError(f"decompose_host_submit: need a contiguous complex64 output [{v.shape[0]},{self.num_moments}]")
        self._pending = getattr(self, "_pending", [])
        self._pending.append((vol, out))
        _lib.check(self._L.z3d_decompose_host_submit(self._h, ctypes.c_void_p(v.ctypes.data), v.shape[0], ctypes.c_void_p(o.ctypes.data)))

    def host_wait(self) -> None:
        """Wait for every submitted host-buffer batch (``z3d_host_wait``)."""
        _lib.check(self._L.z3d_host_wait(self._h))
        self._pending = []

    def reconstruct_host(self, moments, out=None):
        m = moments.numpy() if isinstance(moments, torch.Tensor) else np.asarray(moments)
        single = m.ndim == 1
        m = np.ascontiguousarray(m[None] if single else m, dtype=np.complex64)
        B, N = m.shape[0], self.dim
        if out is None:
            out = np.empty((B, N, N, N), dtype=np.float32)
        o = out.numpy() if isinstance(out, torch.Tensor) else out
        _lib.check(self._L.z3d_reconstruct_host(self._h, ctypes.c_void_p(m.ctypes.data), B,
                                                ctypes.c_void_p(o.ctypes.data)))
        return out[0] if single else out


class PeerComm:
    """Peer-mapped communication buffers of the fused analysis + all-reduce path: one per rank, exchanged as CUDA IPC
    handles through ``torch.distributed`` (control plane only; the data path is the kernel's own NVLink loads)."""

    def __init__(self, plan: Plan, batch_max: int = 1, group=None):
        import torch.distributed as dist

        self._L = _lib.lib()
        self._c = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(self._L.z3d_comm_create(plan._h, int(batch_max), ctypes.byref(self._c), handle))
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        _lib.check(self._L.z3d_comm_connect(self._c, self.rank, self.world, b"".join(handles)))
        dist.barrier(group)

    def error(self) -> int:
        return int(self._L.z3d_comm_error(self._c))

    def close(self):
        c, self._c = self._c, None
        if c:
            self._L.z3d_comm_destroy(c)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def crop_bin(volumes: torch.Tensor, crop_to: int | None = None, bin: float | None = None, center=None, stream=None) -> torch.Tensor:
    """Centred crop and nearest-neighbour binning of ``[B, N, N, N]`` (or ``[N, N, N]``) float32 CUDA volumes on the device, with
    the index arithmetic of ``ZernikeSpherical.bin_crop_array_center`` / ``scipy.ndimage.zoom(order=0)`` (zm:2781-2815,
    2867-2876) - bit-identical to the host path, for ensembles that are already in device memory."""
    single = volumes.dim() == 3
    v = (volumes.unsqueeze(0) if single else volumes).contiguous()
    if v.dtype != torch.float32 or not v.is_cuda or v.shape[1] != v.shape[2] or v.shape[2] != v.shape[3]:
        raise ValueError("crop_bin: need float32 CUDA volumes [B, N, N, N]")
    n_in = int(v.shape[1])
    crop = 0 if crop_to is None else int(crop_to)
    box = n_in if crop == 0 else 2 * int((crop - crop % 2) / 2)
    n_out = box if bin is None else int(round(box * (1 / bin)))
    cz, cy, cx = (n_in // 2,) * 3 if center is None else (int(c) for c in center)
    out = torch.empty((v.shape[0], n_out, n_out, n_out), dtype=torch.float32, device=v.device)
    s = torch.cuda.current_stream(v.device) if stream is None else stream
    _lib.check(_lib.lib().z3d_crop_bin(_ptr(v), v.shape[0], n_in, cz, cy, cx, crop, n_out, _ptr(out), ctypes.c_void_p(s.cuda_stream)))
    return out[0] if single else out


def launch_count() -> int:
    return int(_lib.lib().z3d_launch_count())


def fma_peak_tflops(device: int = 0, variant: int = 0, iters: int = 4096) -> float:
    v = ctypes.c_double()
    _lib.check(_lib.lib().z3d_fma_peak(device, variant, iters, ctypes.byref(v)))
    return v.value


def tf32_peak_tflops(device: int = 0, iters: int = 2048) -> float:
    """Measured dense TF32 tensor-pipe throughput (``z3d_tf32_peak``): back-to-back tcgen05.mma kind::tf32 on every SM."""
    v = ctypes.c_double()
    _lib.check(_lib.lib().z3d_tf32_peak(device, iters, ctypes.byref(v)))
    return v.value
