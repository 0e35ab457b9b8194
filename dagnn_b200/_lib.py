"""ctypes binding of libdagnn_sm100.so (the C ABI declared in include/dagnn_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libdagnn_sm100.so")
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["abi.cu", "schedule.cu", "levels.cu", "embed_readout.cu", "pack.cu", "sweep.cu", "sweep_cluster.cu", "sweep_bwd.cu", "gemm.cu", "rows.cu", "tc_selftest.cu"]
HEADERS = ["common.cuh", "sync.cuh", "tc.cuh"]

MAX_LAYERS = 8
MAX_DIRS = 2
MAX_READOUT_BLOCKS = 20
ABI_VERSION = 10

vp = C.c_void_p


class DagnnSchedule(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("E", C.c_int64), ("B", C.c_int64),
        ("dirs", C.c_int32), ("max_levels", C.c_int32),
        ("perm", vp * MAX_DIRS), ("pos", vp * MAX_DIRS), ("lvl_off", vp * MAX_DIRS), ("rowptr", vp * MAX_DIRS),
        ("col", vp * MAX_DIRS), ("eid", vp * MAX_DIRS), ("eattr", vp * MAX_DIRS),
        ("gptr", vp), ("gdepth", vp), ("summary", vp),
    ]


class DagnnPackLayout(C.Structure):
    _fields_ = [
        ("Din", C.c_int32), ("H", C.c_int32), ("nvid", C.c_int32), ("first_layer", C.c_int32), ("last_layer", C.c_int32),
        ("Hq", C.c_int32), ("Mc", C.c_int32), ("Kin64", C.c_int32), ("Kh64", C.c_int32), ("HP", C.c_int32),
        ("bias_off", C.c_int64), ("wk_off", C.c_int64), ("attnc_off", C.c_int64), ("vidk_off", C.c_int64),
        ("imgx_off", C.c_int64), ("imgh_off", C.c_int64), ("total_floats", C.c_int64),
    ]


class DagnnSweepArgs(C.Structure):
    _fields_ = [
        ("sched", C.POINTER(DagnnSchedule)),
        ("num_layers", C.c_int32),
        ("Din", C.c_int32), ("H", C.c_int32), ("nvid", C.c_int32),
        ("X", vp), ("ldx", C.c_int64), ("X_image", vp),
        ("Hs", (vp * MAX_LAYERS) * MAX_DIRS),
        ("ldh", C.c_int64),
        ("packed", (vp * MAX_LAYERS) * MAX_DIRS),
        ("use_edge_attr", C.c_int32),
        ("workspace", vp), ("workspace_bytes", C.c_size_t), ("trace", vp),
    ]


class DagnnCellParams(C.Structure):
    _fields_ = [("weight_ih", vp), ("weight_hh", vp), ("bias_ih", vp), ("bias_hh", vp), ("attn_w", vp), ("edge_w", vp),
                ("Dq", C.c_int32), ("reserved", C.c_int32)]


class DagnnCellGrads(C.Structure):
    _fields_ = [("weight_ih", vp), ("weight_hh", vp), ("bias_ih", vp), ("bias_hh", vp), ("attn_w", vp), ("edge_w", vp)]


class DagnnSweepBwdArgs(C.Structure):
    _fields_ = [
        ("sched", C.POINTER(DagnnSchedule)),
        ("num_layers", C.c_int32), ("Din", C.c_int32), ("H", C.c_int32), ("nvid", C.c_int32),
        ("X", vp), ("ldx", C.c_int64),
        ("Hs", (vp * MAX_LAYERS) * MAX_DIRS), ("dHs", (vp * MAX_LAYERS) * MAX_DIRS), ("ldh", C.c_int64),
        ("params", (DagnnCellParams * MAX_LAYERS) * MAX_DIRS), ("grads", (DagnnCellGrads * MAX_LAYERS) * MAX_DIRS),
        ("dX", vp), ("lddx", C.c_int64),
        ("use_edge_attr", C.c_int32), ("num_levels", C.c_int32),
        ("lvl_off_host", vp * MAX_DIRS),
        ("workspace", vp), ("workspace_bytes", C.c_size_t),
    ]


class DagnnReadoutBlock(C.Structure):
    _fields_ = [
        ("src", vp), ("ld", C.c_int64), ("width", C.c_int32), ("index_mode", C.c_int32), ("dir", C.c_int32),
        ("filter", C.c_int32), ("filter_lvl", vp), ("out_col", C.c_int32), ("reserved", C.c_int32),
    ]


EXPORTS = {
    "dagnn_abi_version": (C.c_int, []),
    "dagnn_last_error": (C.c_char_p, []),
    "dagnn_launch_count": (C.c_int64, []),
    "dagnn_operand_image_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "dagnn_embed_f32": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, vp, C.c_int64, vp, vp]),
    "dagnn_levels_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "dagnn_levels_build": (C.c_int, [vp, C.c_int64, C.c_int64, C.c_int32, vp, vp, vp, vp, C.c_size_t, vp]),
    "dagnn_schedule_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int32]),
    "dagnn_schedule_build": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.POINTER(DagnnSchedule), vp, C.c_size_t, vp]),
    "dagnn_pack_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(DagnnPackLayout)]),
    "dagnn_pack_params_f32": (C.c_int, [vp, vp, vp, vp, vp, C.c_int32, vp, vp, C.POINTER(DagnnPackLayout), vp, vp]),
    "dagnn_sweep_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32]),
    "dagnn_sweep_trace_bytes": (C.c_size_t, [C.c_int32]),
    "dagnn_sweep_forward_f32": (C.c_int, [C.POINTER(DagnnSweepArgs), vp]),
    "dagnn_tc_selftest_f16x3": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp]),
    "dagnn_tc_selftest_ts": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp]),
    "dagnn_linear_f32": (C.c_int, [vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp]),
    "dagnn_gemm_f32": (C.c_int, [vp, C.c_int64, C.c_int32, vp, C.c_int64, C.c_int32, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, vp]),
    "dagnn_sweep_backward_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64]),
    "dagnn_sweep_backward_f32": (C.c_int, [C.POINTER(DagnnSweepBwdArgs), vp]),
    "dagnn_readout_backward_f32": (C.c_int, [C.POINTER(DagnnSchedule), C.POINTER(DagnnReadoutBlock), C.POINTER(vp), C.c_int32, C.c_int32,
                                             vp, vp, C.c_int64, vp]),
    "dagnn_embed_backward_f32": (C.c_int, [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, vp, C.c_int64, vp, vp, vp, vp]),
    "dagnn_dvae_rows_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "dagnn_dvae_rows_build": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int64, vp, vp, vp, vp, C.c_size_t, vp]),
    "dagnn_readout_f32": (C.c_int, [C.POINTER(DagnnSchedule), C.POINTER(DagnnReadoutBlock), C.c_int32, C.c_int32, vp,
                                    C.c_int64, vp]),
    "dagnn_states_to_node_order_f32": (C.c_int, [C.POINTER(DagnnSchedule), C.c_int32, vp, C.c_int64, C.c_int32, vp,
                                                 C.c_int64, vp]),
}


class DagnnError(RuntimeError):
    pass


def nvcc_command(out_path: str = LIB_PATH):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-I", os.path.join(ROOT, "include"), "-shared", "-Xcompiler", "-fPIC", "-o", out_path] + \
           os.environ.get("DAGNN_NVCC_FLAGS", "").split() + [os.path.join(CSRC, s) for s in SOURCES]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into dagnn_b200/libdagnn_sm100.so (in-tree)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "dagnn_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise DagnnError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_LIB = None


def lib():
    """Load the shared library (once). Raises if it has not been built — there is no fallback path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DagnnError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(nvcc, sm_100a). dagnn_b200 has no CPU / eager fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.dagnn_abi_version() != ABI_VERSION:
            raise DagnnError("ABI version mismatch: library %d, binding %d" % (L.dagnn_abi_version(), ABI_VERSION))
        _LIB = L
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dagnn_last_error()
        raise DagnnError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count() -> int:
    return int(lib().dagnn_launch_count())
