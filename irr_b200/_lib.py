"""ctypes binding of libirr_b200.so — the C ABI declared in include/irr_b200.h.

There is no CPU or library fallback: if the shared library is missing or a call returns non-zero, a RuntimeError
is raised (the reference turns a failed launch into AT_ERROR -> RuntimeError, correlation_cuda.cc:78-80).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IRR_B200_LIB: load an experimental build of the same sources instead (kernel A/B runs only)
LIB_PATH = os.environ.get("IRR_B200_LIB") or os.path.join(_HERE, "libirr_b200.so")

c_fp = C.c_void_p  # device pointers are passed as integers (tensor.data_ptr())


class ConvSeg(C.Structure):
    """``irr_conv_seg`` of include/irr_b200.h (one output-channel segment of irr_conv2d_fwd_multi)."""
    _fields_ = [("n_begin", C.c_int), ("addend_pre", C.c_int), ("leaky_slope", C.c_float), ("alpha", C.c_float),
                ("addend", C.c_void_p), ("addend_bs", C.c_longlong), ("y", C.c_void_p), ("y_bs", C.c_longlong)]


c_ll = C.c_longlong
c_i = C.c_int
c_f = C.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES).  Must match include/irr_b200.h exactly;
# tests/test_abi.py cross-checks this table against the header and the exported symbols.
PROTOTYPES = {
    "irr_abi_version": [],
    "irr_last_error": [],
    "irr_device_info": [C.POINTER(c_i), C.POINTER(c_i), C.POINTER(c_i)],
    "irr_correlation_fwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_fp],
    "irr_warp_correlation_fwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i,
                                 c_i, c_f, c_i, c_i, c_f, c_i, c_fp],
    "irr_correlation_workspace_bytes": [c_i, c_i, c_i, c_i, c_i],
    "irr_warp_correlation_fwd_ws": [c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i,
                                    c_i, c_f, c_i, c_i, c_f, c_i, c_fp, C.c_size_t, c_i, c_fp],
    "irr_warp_correlation_fwd_dt": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_i, c_i, c_i, c_i,
                                    c_i, c_i, c_f, c_i, c_i, c_f, c_i, c_fp, C.c_size_t, c_i, c_fp],
    "irr_round_bf16_store_fwd": [c_fp, c_ll, c_i, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_fp],
    "irr_warp_fwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_fp, c_ll, c_fp, c_i, c_i, c_i, c_i, c_i, c_i,
                     c_f, c_i, c_i, c_i, c_fp],
    "irr_warp_bwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_f,
                     c_i, c_fp],
    "irr_correlation_generic_fwd": [c_fp, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fp],
    "irr_correlation_generic_out_shape": [c_i, c_i, c_i, c_i, c_i, c_i, c_i, C.POINTER(c_i), C.POINTER(c_i),
                                          C.POINTER(c_i)],
    "irr_conv2d_packed_bytes": [c_i, c_i, c_i, c_i],
    "irr_conv2d_math_supported": [c_i, c_i, c_i, c_i, c_i, c_i],
    "irr_conv2d_direct_supported": [c_i, c_i, c_i],
    "irr_conv2d_pack_weights": [c_fp, c_fp, c_i, c_i, c_i, c_i, c_fp],
    "irr_conv2d_fwd": [c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f,
                       c_f, c_i, c_i, c_i, c_fp],
    "irr_conv2d_workspace_bytes": [c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i],
    "irr_conv2d_fwd_ws": [c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f,
                          c_f, c_i, c_fp, C.c_size_t, c_i, c_i, c_fp],
    "irr_conv2d_fwd_dual": [c_fp, c_ll, c_fp, c_fp, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f,
                            c_f, c_i, c_fp, c_ll, c_fp, c_ll, c_f, c_f, c_i, c_fp, C.c_size_t, c_i, c_i, c_fp],
    "irr_conv2d_fwd_multi": [c_fp, c_ll, c_fp, c_fp, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, C.POINTER(ConvSeg), c_i, c_i, c_fp,
                             C.c_size_t, c_i, c_i, c_fp],
    "irr_resize_bilinear_ac_fwd": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_i, c_i, c_fp],
    "irr_scale_channels_fwd": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_ll, c_f, c_f, c_fp],
    "irr_round_bf16_fwd": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_ll, c_fp],
    "irr_upsample_nearest2x_fwd": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_fp],
    "irr_sub_spatial_mean_fwd": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_fp],
    "irr_channel_l2norm_fwd": [c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_ll, c_fp],
    "irr_correlation_bwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_fp],
    "irr_eval_metrics_fwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_fp, c_i, c_i, c_i, c_fp],
    "irr_refine_gather_fwd": [c_fp, c_ll, c_fp, c_ll, c_fp, c_ll, c_i, c_i, c_i, c_i, c_i, c_fp],
}
_RESTYPES = {"irr_last_error": C.c_char_p, "irr_conv2d_packed_bytes": C.c_size_t,
             "irr_conv2d_workspace_bytes": C.c_size_t, "irr_correlation_workspace_bytes": C.c_size_t}

_lib = None


def load() -> C.CDLL:
    """Load (once) and type the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"irr_b200: native library {LIB_PATH} is missing — build it with "
            f"`make -C irr_b200/csrc` (or __graft_entry__.build()).  There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_i)
    if lib.irr_abi_version() != 2:
        raise RuntimeError("irr_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().irr_last_error()
        raise RuntimeError(f"irr_b200.{what} failed (rc={rc}): {msg.decode() if msg else ''}")
