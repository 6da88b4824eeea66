"""ctypes binding of the C ABI declared in ``include/zernike3d_b200.h``.

There is deliberately no fallback: if ``libz3d_b200.so`` is missing, importing the compute entry points
raises; if no sm_100 device is present, plan creation raises.  PyTorch is used by callers only to own
device buffers and streams - this module passes raw pointers.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libz3d_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NODEVICE, ERR_WORKSPACE, ERR_UNSUPPORTED = range(6)


class Z3DError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"zernike3d_b200 error {code}: {msg}")
        self.code = code


_lib = None


def build(verbose: bool = False) -> str:
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    import subprocess

    cmd = ["make", "-C", os.path.join(HERE, "csrc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libz3d_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing - build it with `make -C zernike3d_b200/csrc` "
            "(or __graft_entry__.build()); zernike3d_b200 has no CPU or PyTorch fallback")
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_void_p, c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
    L.z3d_version.restype = c_int
    L.z3d_last_error.restype = ctypes.c_char_p
    L.z3d_num_moments.argtypes = [c_int]
    L.z3d_num_radial.argtypes = [c_int]
    L.z3d_num_harmonics.argtypes = [c_int]
    L.z3d_nlm_table.argtypes = [c_int, ctypes.POINTER(c_int)]
    L.z3d_plan_create.argtypes = [c_int, c_int, c_int, ctypes.POINTER(c_void_p)]
    L.z3d_plan_destroy.argtypes = [c_void_p]
    L.z3d_plan_info.argtypes = [c_void_p, ctypes.POINTER(c_int), c_int]
    L.z3d_stream_layout.argtypes = [c_int, c_int, ctypes.POINTER(c_int), c_int]
    L.z3d_gen_layout.argtypes = [c_int, c_int, ctypes.POINTER(c_int), c_int]
    L.z3d_gen_layout_tiles.argtypes = [c_int, c_int, c_int, ctypes.POINTER(c_int), c_int]
    L.z3d_plan_prepare_tables.argtypes = [c_void_p, c_int, c_size_t, ctypes.POINTER(c_int)]
    L.z3d_zcc.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    L.z3d_crop_bin.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    L.z3d_workspace_bytes.argtypes = [c_void_p, c_int]
    L.z3d_workspace_bytes.restype = c_size_t
    L.z3d_decompose.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.z3d_decompose_part.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.z3d_comm_create.argtypes = [c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.c_char_p]
    L.z3d_comm_connect.argtypes = [c_void_p, c_int, c_int, ctypes.c_char_p]
    L.z3d_comm_destroy.argtypes = [c_void_p]
    L.z3d_comm_error.argtypes = [c_void_p]
    L.z3d_decompose_allreduce.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.z3d_reconstruct.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.z3d_reconstruct_batched.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    L.z3d_reconstruct_batched_workspace_bytes.argtypes = [c_void_p, c_int]
    L.z3d_reconstruct_batched_workspace_bytes.restype = c_size_t
    L.z3d_invariants.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    L.z3d_scale_moments.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]
    L.z3d_decompose_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    L.z3d_decompose_host_submit.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    L.z3d_host_wait.argtypes = [c_void_p]
    L.z3d_reconstruct_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    L.z3d_launch_count.restype = ctypes.c_ulonglong
    L.z3d_profile_enable.argtypes = [c_void_p, c_int]
    L.z3d_profile_last_ms.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_float)]
    L.z3d_fma_peak.argtypes = [c_int, c_int, c_int, ctypes.POINTER(ctypes.c_double)]
    L.z3d_tf32_peak.argtypes = [c_int, c_int, ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise Z3DError(rc, lib().z3d_last_error().decode())


EXPORTS = [
    "z3d_version", "z3d_last_error", "z3d_num_moments", "z3d_num_radial", "z3d_num_harmonics", "z3d_nlm_table",
    "z3d_plan_create", "z3d_plan_destroy", "z3d_plan_info", "z3d_workspace_bytes", "z3d_decompose",
    "z3d_reconstruct", "z3d_invariants", "z3d_decompose_host", "z3d_reconstruct_host", "z3d_launch_count",
    "z3d_fma_peak", "z3d_profile_enable", "z3d_profile_last_ms", "z3d_decompose_part", "z3d_comm_create", "z3d_comm_connect", "z3d_comm_destroy", "z3d_comm_error",
    "z3d_decompose_allreduce", "z3d_scale_moments", "z3d_stream_layout", "z3d_zcc", "z3d_crop_bin", "z3d_gen_layout", "z3d_gen_layout_tiles",
    "z3d_plan_prepare_tables", "z3d_tf32_peak", "z3d_reconstruct_batched", "z3d_reconstruct_batched_workspace_bytes", "z3d_decompose_host_submit", "z3d_host_wait",
]
