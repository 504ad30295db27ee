"""ctypes binding of libagenda_b200.so (the C ABI declared in include/agenda_b200.h).

There is NO fallback: if the shared library has not been built, or a tensor is not on a CUDA device, the call
raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a, in-tree).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libagenda_b200.so")

F32, BF16 = 0, 1


class AgendaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libagenda_b200 error {code}: {message}")
        self.code = code


# name -> (argtypes); every entry point returns int.  Mirrors include/agenda_b200.h one to one.
_SIGNATURES = {
    "agenda_attn_self_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                             c_void_p],
    "agenda_attn_self_fwd_strided": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     ctypes.c_longlong, c_float, c_void_p],
    "agenda_attn_self_fwd_variant": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int,
                                     c_void_p],
    "agenda_attn_self_fwd_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p],
    "agenda_attn_cross_fwd_heat": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_float, ctypes.POINTER(c_int32), c_int, c_int, c_void_p, c_int, c_void_p],
    "agenda_attn_cross_fwd_heat_heads": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_float, ctypes.POINTER(c_int32), c_int, c_int, c_void_p, c_int,
                                         c_void_p],
    "agenda_attn_cross_fwd_heat_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_float, ctypes.POINTER(c_int32), c_int, c_int, c_void_p, c_int,
                                       c_void_p],
    "agenda_attn_fwd_masked": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                               c_void_p, c_int, ctypes.POINTER(c_int32), c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "agenda_linear_split_f32": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "agenda_linear_split_pack_w": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "agenda_linear_split_f32_packed": [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p],
    "agenda_linear_split_f32_heads": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "agenda_pack_context_kv": [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "agenda_attn_cross_fwd_heat_x3": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                      ctypes.POINTER(c_int32), c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "agenda_attn_cross_fwd_heat_x3_hm": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                         ctypes.POINTER(c_int32), c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "agenda_attn_cross_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_float, ctypes.POINTER(c_int32), c_int, c_int,
                              c_void_p],
    "agenda_attn_cross_bwd_tc": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_int, c_float, ctypes.POINTER(c_int32), c_int, c_int, c_void_p],
    "agenda_attn_self_fwd_lse": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p],
    "agenda_attn_self_bwd_lse": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "agenda_attn_self_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                             c_int, c_int, c_int, c_int, c_float, c_void_p],
    "agenda_heat_upsample_accum": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "agenda_heat_upsample_accum_heads": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "agenda_heat_finalize": [c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "agenda_heat_normalize_u8": [c_void_p, c_void_p, c_int, c_int, c_void_p],
    "agenda_resize_bicubic_u8": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "agenda_heat_to_u8_image": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "agenda_stack_heatmaps_u8": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "agenda_heat_postprocess_stack": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p],
    "agenda_groupnorm_nhwc": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                              c_int, c_void_p],
    "agenda_add_bias_residual": [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p],
    "agenda_geglu": [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p],
    "agenda_layernorm": [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_float, c_void_p],
    "agenda_ccl_bbox": [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
}
EXPORTS = (["agenda_version", "agenda_last_error", "agenda_device_ok", "agenda_context_blob_bytes",
            "agenda_attn_self_bwd_workspace_bytes", "agenda_groupnorm_workspace_bytes",
            "agenda_attn_cross_bwd_tc_workspace_bytes", "agenda_attn_self_fwd_emits_lse", "agenda_linear_split_pack_bytes"] + list(_SIGNATURES))

_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built (run __graft_entry__.build()). "
            "agenda_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.agenda_version.restype = c_int
    lib.agenda_version.argtypes = []
    lib.agenda_last_error.restype = c_char_p
    lib.agenda_last_error.argtypes = []
    lib.agenda_device_ok.restype = c_int
    lib.agenda_device_ok.argtypes = []
    lib.agenda_context_blob_bytes.restype = ctypes.c_longlong
    lib.agenda_context_blob_bytes.argtypes = [c_int, c_int, c_int]
    lib.agenda_attn_self_bwd_workspace_bytes.restype = ctypes.c_longlong
    lib.agenda_attn_self_bwd_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.agenda_attn_cross_bwd_tc_workspace_bytes.restype = ctypes.c_longlong
    lib.agenda_attn_cross_bwd_tc_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.agenda_linear_split_pack_bytes.restype = ctypes.c_longlong
    lib.agenda_linear_split_pack_bytes.argtypes = [c_int, c_int, c_int]
    lib.agenda_attn_self_fwd_emits_lse.restype = c_int
    lib.agenda_attn_self_fwd_emits_lse.argtypes = [c_int, c_int]
    lib.agenda_groupnorm_workspace_bytes.restype = ctypes.c_longlong
    lib.agenda_groupnorm_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


launches = 0          # number of kernel launches issued through the C ABI (every entry point launches exactly one)
event_sink = None     # bench.py: {entry_point_name: [(start_event, end_event, args), ...]} to time single kernels


def call(name: str, *args) -> None:
    global launches
    lib = load()
    sink = event_sink.get(name) if event_sink is not None else None
    if sink is not None:
        import torch
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise AgendaError(rc, lib.agenda_last_error().decode("utf-8", "replace"))
    launches += 1
    if sink is not None:
        end.record()
        sink.append((start, end, args))
