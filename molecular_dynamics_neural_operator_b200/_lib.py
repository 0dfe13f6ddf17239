"""ctypes binding of libgkn_b200.so — the only way the Python host code reaches the CUDA kernels.

Signatures mirror ``include/gkn_b200.h`` one to one.  There is no fallback: if the library is
missing and cannot be built, or an entry point returns non-zero, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_int, c_int32, c_int64, c_longlong, c_void_p

from . import _build

_LIB = None

P = c_void_p  # device pointer / stream

# name -> (restype, argtypes); must list every symbol declared in include/gkn_b200.h
SIGNATURES = {
    "gkn_version": (c_int, []),
    "gkn_last_error": (ctypes.c_char_p, []),
    "gkn_launch_count": (c_longlong, []),
    "gkn_device_info": (c_int, [ctypes.POINTER(c_int)] * 3),
    "gkn_graph_prepare_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "gkn_graph_prepare": (c_int, [P, c_int64, c_int64, P, P, P, P, P, P, P, P, P, c_int64, P]),
    "gkn_plan_from_csr": (c_int, [P, P, c_int64, c_int64, P, P, P, P, P, P, P]),
    "gkn_radius_graph_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "gkn_radius_graph_count": (c_int, [P, c_int64, c_int64, c_double, P, P, P, c_int64, P]),
    "gkn_radius_graph_fill": (
        c_int,
        [P, c_int64, c_int64, c_double, P, c_int64, P, P, P, c_int64, P],
    ),
    "gkn_contract_fwd_bf16": (c_int, [P, P, P, P, P, c_int64, c_int32, c_int32, P]),
    "gkn_contract_dx_bf16": (c_int, [P, P, P, P, P, c_int64, c_int32, c_int32, P]),
    "gkn_contract_fwd_f32": (c_int, [P, P, P, P, P, c_int64, c_int32, c_int32, P]),
    "gkn_contract_dx_f32": (c_int, [P, P, P, P, P, c_int64, c_int32, c_int32, P]),
    "gkn_node_msg_bf16": (
        c_int,
        [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int32, P, c_int32, c_int32, P],
    ),
    "gkn_node_du_bf16": (c_int, [P, P, P, c_int64, c_int64, c_int64, P, c_int32, c_int32, P]),
    "gkn_node_dh_bf16": (
        c_int,
        [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, P, c_int32, c_int32, P],
    ),
    "gkn_node_du_mn_bf16": (c_int, [P, P, P, c_int64, c_int64, c_int64, P, c_int32, c_int32, P]),
    "gkn_node_dh_mn_bf16": (
        c_int,
        [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, P, c_int32, c_int32, P],
    ),
    "gkn_lstm_blocks": (c_int, [c_int64]),
    "gkn_lstm_fwd": (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int32, P]),
    "gkn_lstm_bwd": (c_int, [P, P, P, P, P, P, P, P, P, c_int64, c_int64, c_int32, P]),
    "gkn_relu_bits_bf16": (c_int, [P, c_int64, c_int32, P, P]),
    "gkn_node_dh_mn_relu_bf16": (
        c_int, [P, P, P, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, P, P, P, c_int32, P,
                c_int64, c_int32, c_int32, P]),
    "gkn_node_gather_bf16": (c_int, [P, P, c_int64, P, P, c_int32, P]),
    "gkn_node_fused_fwd_bf16": (
        c_int,
        [P, P, P, P, P, c_int64, P, P, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, P],
    ),
    "gkn_node_fused_bwd_grid": (c_int, [c_int32, c_int32, ctypes.POINTER(c_int32),
                                        ctypes.POINTER(c_int32)]),
    "gkn_node_fused_bwd_bf16": (
        c_int,
        [P, P, P, P, P, P, P, c_int64, P, P, P, c_int32, c_int64, c_int64, c_int64, c_int32,
         c_int32, P],
    ),
    "gkn_sum_partials": (c_int, [P, c_int32, c_int64, P, P, P]),
    "gkn_scatter_mean_fwd": (
        c_int,
        [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int32, c_int32, P],
    ),
    "gkn_scatter_src_sum": (c_int, [P, P, P, P, P, P, P, c_int64, c_int64, c_int32, P]),
    "gkn_segment_matvec": (
        c_int,
        [P, P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int32, c_int32, P],
    ),
    "gkn_shard_reduce_update": (
        c_int,
        [P, P, ctypes.c_uint64, ctypes.c_uint64, c_int32, P, P, P, P, P, c_int64, c_int64, c_int64,
         c_int32, c_int32, P],
    ),
    "gkn_edge_attr_gather": (c_int, [P, P, c_int64, c_int64, P, P, P]),
    "gkn_xty_workspace_bytes": (c_int64, [c_int32]),
    "gkn_xty_f32": (c_int, [P, P, c_int64, c_int32, P, P, P, c_int64, P]),
    "gkn_relu_grad_scale": (c_int, [P, P, P, P, P, c_int64, c_int32, c_int32, P]),
    "gkn_edge_outer_sum": (
        c_int,
        [P, P, P, P, P, c_int32, c_int64, c_int64, c_int64, c_int32, c_int32, P],
    ),
    "gkn_edge_grad_workspace_bytes": (c_int64, [c_int64, c_int32]),
    "gkn_edge_grad_max_slots": (c_int, [P, c_int64, P, P]),
    "gkn_edge_grad_dh_bf16": (
        c_int,
        [P, P, P, P, P, P, c_int64, c_int64, c_int32, c_int32, c_int32, P, c_int64, P],
    ),
    "gkn_bias_act_cast_bf16": (c_int, [P, P, P, c_int64, c_int64, c_int32, P]),
    "gkn_edge_layer0": (c_int, [P, P, P, P, c_int32, c_int64, c_int32, c_int32, P]),
    "gkn_mask_cast_colsum_workspace_bytes": (c_int64, [c_int64, c_int32]),
    "gkn_mask_cast_colsum": (
        c_int,
        [P, c_int32, P, c_int32, P, c_int32, P, c_int32, P, c_int32, c_int64, c_int64, P, c_int64, P],
    ),
}


def load():
    """Load (building first if needed) the CUDA library.  Fails loudly; never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if not _build.is_current():
        if _build.find_nvcc() is not None:
            path = _build.build()
        elif not path.exists():
            raise RuntimeError(
                f"{path} is missing and nvcc is unavailable: the GKN CUDA extension is required "
                "(there is no CPU or PyTorch fallback)"
            )
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = ABI/header mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def last_error() -> str:
    return load().gkn_last_error().decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libgkn_b200 {what} failed (code {rc}): {last_error()}")


def launch_count() -> int:
    return int(load().gkn_launch_count())
