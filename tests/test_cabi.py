"""The C-ABI shared library loads without a GPU, exports every symbol include/zernike3d_b200.h declares, and its
host-only entry points work; compute entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from zernike3d_b200 import _lib
from zernike3d_b200 import indexing as ix


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zernike3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(z3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 17 and set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/zernike3d_b200.h but not exported"


def test_host_only_entry_points():
    L = _lib.lib()
    assert L.z3d_version() >= 100
    assert L.z3d_num_moments(50) == 12051 and L.z3d_num_radial(50) == 676 and L.z3d_num_harmonics(50) == 1326
    assert L.z3d_num_moments(-1) == -1
    n = L.z3d_num_moments(9)
    buf = (ctypes.c_int * (3 * n))()
    assert L.z3d_nlm_table(9, buf) == 0
    assert np.array_equal(np.array(buf).reshape(n, 3), ix.nlm_table(9))
    assert L.z3d_nlm_table(-3, buf) == _lib.ERR_INVALID and b"bad arguments" in L.z3d_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_device_fails_loudly():
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.z3d_plan_create(10, 16, 0, ctypes.byref(h))
    assert rc == _lib.ERR_NODEVICE and not h.value
    assert b"no CPU fallback" in L.z3d_last_error()
    from zernike3d_b200.plan import Plan

    with pytest.raises(_lib.Z3DError):
        Plan(10, 16)


def test_bad_arguments_rejected_before_touching_the_device():
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.z3d_plan_create(-1, 16, 0, ctypes.byref(h)) == _lib.ERR_INVALID
    assert L.z3d_plan_create(10, 1, 0, ctypes.byref(h)) == _lib.ERR_INVALID
    assert L.z3d_decompose(None, None, 1, 0, 1, None, None, 0, None) == _lib.ERR_INVALID
    assert L.z3d_workspace_bytes(None, 1) == 0
    assert L.z3d_scale_moments(None, None, 1, None, 1, None, None) == _lib.ERR_INVALID
    assert L.z3d_zcc(None, None, None, 1, None, None) == _lib.ERR_INVALID
    assert L.z3d_crop_bin(None, 1, 32, 16, 16, 16, 0, 16, None, None) == _lib.ERR_INVALID
    assert L.z3d_reconstruct(None, None, 1, 0, 1, None, None, 0, None) == _lib.ERR_INVALID


@pytest.mark.parametrize("sms", [148, 132, 64])
def test_streamed_T_column_set_layout_for_every_order(sms):
    """Host-only planner of the streamed-T batched kernel (z3d_stream.cuh): for every supported order the column sets respect
    the kernel's limits (<= 512 accumulator columns, <= 4 panels, <= 8 segments of <= 256 rows, one fold kind), every moment
    owns exactly one column and sets x splits fit the device."""
    L = _lib.lib()
    out = (ctypes.c_int * 8)()
    for order in list(range(0, 75)):
        assert L.z3d_stream_layout(order, sms, out, 8) == 0
        ok, nsets, nsplits, rows16, widest, bad, maxpanel, maxseg = list(out)
        nmom = L.z3d_num_moments(order)
        if not ok:                       # no layout: batches of such a plan use k_decompose_batched / the per-volume kernel
            continue
        assert bad == 0, (order, list(out))
        assert 1 <= nsets * nsplits <= sms and nmom <= rows16 <= nmom + 16 * (2 * (order + 1) + nsets)
        assert widest <= 512 and rows16 % 16 == 0 and maxpanel <= 4 and maxseg <= 8
    assert L.z3d_stream_layout(50, 148, out, 8) == 0 and list(out)[:4] == [1, 37, 4, 12832]
    assert L.z3d_stream_layout(-1, 148, out, 8) == _lib.ERR_INVALID


@pytest.mark.parametrize("sms", [148, 132])
def test_table_free_column_set_layouts_for_every_order(sms):
    """Host-only planner of the table-free tensor-core kernels (z3d_gen.cuh / z3d_synth.cuh): for every supported order the column sets of
    the analysis layout (ntile 1) and of the two-tile layout the batched synthesis runs on (ntile 2) pass the planner's self-check
    (segments of 16..256 columns laid end to end, accumulators + a-panel slots within the 512 tensor-memory columns - which for ntile 2
    implies the synthesis kernel's budget 2 ncols + 48 npanel <= 512 -, every chain row written exactly twice, every moment owning one
    column) and sets x splits fit the device."""
    L = _lib.lib()
    out = (ctypes.c_int * 11)()
    for ntile in (1, 2):
        for order in range(0, 75):
            assert L.z3d_gen_layout_tiles(order, sms, ntile, out, 11) == 0
            ok, nsets, nsplits, cols_all, widest, bad, maxpanel, maxseg, maxtask, maxseed, est = list(out)
            if not ok:          # orders whose sets outnumber the SMs fall back to the per-volume kernels
                assert order > 40
                continue
            nmom = L.z3d_num_moments(order)
            assert bad == 0, (ntile, order, list(out))
            assert 1 <= nsets * nsplits <= sms and cols_all >= nmom and cols_all % 16 == 0
            assert widest % 16 == 0 and widest <= 512 // ntile and 1 <= maxpanel <= 8 and maxseg <= 8 and maxtask <= 256
            if ntile == 2:
                assert 2 * widest + 48 * 1 <= 512
    assert L.z3d_gen_layout_tiles(50, 148, 1, out, 11) == 0 and list(out)[:3] == [1, 37, 4]
    assert L.z3d_gen_layout_tiles(50, 148, 2, out, 11) == 0 and list(out)[:3] == [1, 74, 2]
    assert L.z3d_gen_layout_tiles(50, 148, 3, out, 11) == _lib.ERR_INVALID
