"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/dagnn_b200.h declares
(no compute calls without a GPU); host-side queries (layouts, workspace sizes, error strings) behave."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dagnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dagnn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built_lib):
    from dagnn_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 14
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "libdagnn_sm100.so does not export %s" % name
        assert name in _lib.EXPORTS, "%s is declared in the header but has no ctypes binding" % name
    assert sorted(_lib.EXPORTS) == declared
    assert built_lib.dagnn_abi_version() == _lib.ABI_VERSION == 10


def test_pack_layout_and_workspace_queries(built_lib):
    from dagnn_b200 import _lib
    L = _lib.DagnnPackLayout()
    assert built_lib.dagnn_pack_layout(256, 256, 0, 1, 0, C.byref(L)) == 0     # first of several layers
    assert (L.Hq, L.Mc, L.Kin64, L.Kh64, L.HP) == (256, 768, 256, 256, 256)
    assert L.imgh_off - L.imgx_off == (L.Mc // 64) * (L.Kin64 // 64) * 2 * 4096 // 2      # fp16 images counted in 4-byte units
    assert L.total_floats - L.imgh_off == 2 * (L.Mc // 64) * (L.Kh64 // 64) * 2 * 4096 // 2   # [W_hh ; W_ih of the next layer]
    assert L.imgx_off % 256 == 0
    assert built_lib.dagnn_pack_layout(256, 256, 0, 0, 1, C.byref(L)) == 0     # last layer: no next-layer block, no input image
    assert L.imgh_off == L.imgx_off and L.total_floats - L.imgh_off == (L.Mc // 64) * (L.Kh64 // 64) * 2 * 4096 // 2
    assert built_lib.dagnn_pack_layout(8, 501, 8, 1, 0, C.byref(L)) == 0        # D-VAE NA: Din = 8, H = 501, 8 vertex-id columns
    assert (L.Hq, L.Mc, L.Kin64, L.Kh64, L.HP) == (504, 1536, 64, 512, 512)
    assert built_lib.dagnn_pack_layout(0, 16, 0, 1, 1, C.byref(L)) != 0         # bad argument -> error code + message
    assert b"pack_layout" in built_lib.dagnn_last_error()
    ws = built_lib.dagnn_sweep_workspace_bytes(2, 2, 256, 256, 16478, 24491, 256)
    assert ws >= 256 + 4 * (16478 * 4 + 2 * 16478 * 768 * 4)
    assert built_lib.dagnn_sweep_workspace_bytes(3, 2, 256, 256, 10, 10, 256) == 0   # dirs out of range
    assert built_lib.dagnn_schedule_workspace_bytes(1000, 2000, 256) > 0
    assert built_lib.dagnn_sweep_trace_bytes(10) == 11 * 256 * 16 * 8
    assert built_lib.dagnn_levels_workspace_bytes(1000, 257) >= 4 * (2000 + 257)
    assert built_lib.dagnn_levels_workspace_bytes(1000, 0) == 0                   # bad argument
    # argument validation comes before any CUDA call: error code + message, nothing launched
    n0 = built_lib.dagnn_launch_count()
    assert built_lib.dagnn_levels_build(None, 10, 0, 4, None, None, None, None, 0, None) != 0
    assert b"levels" in built_lib.dagnn_last_error() and built_lib.dagnn_launch_count() == n0


def test_no_cpu_path():
    """The product refuses CPU tensors instead of computing on the host."""
    import torch
    from dagnn_b200 import _lib, data as D, ogb
    B = D.make_code2_batch(2, 5)
    enc = ogb.ASTNodeEncoder(16, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(50, 5, 16, 16, None, encoder=enc, out_wx=False, out_pool_all=False).eval()
    with pytest.raises(_lib.DagnnError):
        with torch.no_grad():
            m(B)
    with pytest.raises(NotImplementedError):
        ogb.DAGNN(50, 5, 16, 16, None, encoder=enc, agg="gated_sum")
    with pytest.raises(ValueError):
        ogb.DAGNN(50, 5, 32, 16, None, encoder=enc, agg_x=True)               # dagnn.py:27-28
