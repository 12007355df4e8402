"""ctypes binding of libtstereo.so — the C-ABI boundary declared in include/tstereo.h.

There is no fallback: if the shared library is missing or a symbol is absent the import of
any op fails loudly.  `SIGNATURES` is the single Python-side statement of the ABI and is checked
against the header by tests/test_abi.py.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSTEREO_LIB", os.path.join(_HERE, "libtstereo.so"))

P, I, F, LL = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> (restype, argtypes); order and types mirror include/tstereo.h exactly
SIGNATURES = {
    "tstereo_version": (I, []),
    "tstereo_last_error": (C.c_char_p, []),
    "tstereo_build_id": (C.c_char_p, []),
    "tstereo_launch_count": (LL, []),
    "tstereo_block_cost_scratch_floats": (LL, [I, I, I, I, I]),
    "tstereo_block_cost_shift": (I, [P, P, P, P, I, I, I, I, I, P]),
    "tstereo_block_cost_warp": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "tstereo_group_cost_shift": (I, [P, P, P, P, I, I, I, I, I, P]),
    "tstereo_block_cost_shift_s": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "tstereo_group_cost_warp": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "tstereo_cost_conv_wpack_floats": (LL, [I, I, I]),
    "tstereo_cost_conv_warp": (I, [P, P, P, P, P, LL, LL, LL, P, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_cost_conv_shift": (I, [P, P, P, P, LL, LL, LL, P, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_cost_taps": (I, [P, P, P, P, P, P, LL, LL, LL, P, I, I, I, I, I, I, P]),
    "tstereo_conv_hw3": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, I, I, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_conv_hw3_tc2_wpack_floats": (LL, [I, I, I]),
    "tstereo_conv_hw3_tc2": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_conv_hw3s2_tc2_wpack_floats": (LL, [I, I, I]),
    "tstereo_conv_hw3s2_tc2": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_deconv_hw_tc2_wpack_floats": (LL, [I, I, I]),
    "tstereo_deconv_hw_tc2": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_conv_d_tc2_wpack_floats": (LL, [I, I, I, I]),
    "tstereo_conv_d_tc2": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_split_pack": (I, [P, LL, LL, LL, P, I, I, I, I, I, P]),
    "tstereo_conv_hw3_s": (I, [P, LL, LL, LL, P, P, LL, LL, LL, P, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_conv_hw3s2_s": (I, [P, LL, LL, LL, P, P, LL, LL, LL, P, P, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_deconv_hw_s": (I, [P, LL, LL, LL, P, P, LL, LL, LL, P, P, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_conv_d_s": (I, [P, LL, LL, LL, P, P, LL, LL, LL, P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_conv_d": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, I, I, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_deconv_hw": (I, [P, LL, LL, LL, P, LL, LL, LL, P, P, I, I, I, I, I, I, I, I, P]),
    "tstereo_copy_planes": (I, [P, P, LL, LL, I, I, I, P]),
    "tstereo_resize_add_act": (I, [P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_resize_add_act_s": (I, [P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "tstereo_pool5": (I, [P, LL, LL, P, P, LL, LL, I, I, I, I, I, P]),
    "tstereo_merge_memory": (I, [P, P, P, P, P, P, P, LL, LL, P, I, I, I, I, I, I, P]),
    "tstereo_heads": (I, [P, P, P, P, I, I, I, I, I, F, P]),
    "tstereo_predict_disp": (I, [P, P, P, P, P, P, I, I, I, I, P]),
    "tstereo_range_samples": (I, [P, F, P, P, P, I, I, I, I, I, P]),
    "tstereo_convex_upsample": (I, [P, P, P, P, P, I, I, I, P]),
    "tstereo_unet_upsample": (I, [P, P, P, I, I, I, I, I, P]),
    "tstereo_bilinear_resize": (I, [P, P, F, F, I, I, I, I, I, I, I, I, P]),
    "tstereo_normalize_u8": (I, [P, P, LL, LL, I, I, I, P, P, P]),
    "tstereo_disp_error": (I, [P, P, F, F, I, I, LL, P, P]),
    "tstereo_loss_smooth_l1": (I, [P, P, I, I, I, I, I, F, F, I, P, P]),
    "tstereo_loss_wasserstein": (I, [P, P, P, P, I, I, I, I, I, I, F, F, I, P, P]),
    "tstereo_pose_prep": (I, [P, P, P, P, F, P, I, P]),
    "tstereo_reproject_disp": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "tstereo_project_to_3d": (I, [P, P, P, P, I, I, I, I, P]),
    "tstereo_splat_metric": (I, [P, P, P, I, I, I, I, P]),
    "tstereo_update_map_scratch_floats": (LL, [I, I, I, I, I]),
    "tstereo_update_map": (I, [P, I, I, P, P, P, P, P, P, I, P, I, I, P, P, P, P, I, I, I, P]),
    "tstereo_softsplat": (I, [P, P, P, P, P, I, I, I, I, P]),
}



class SplitStruct(C.Structure):
    """`tstereo_split` of include/tstereo.h: an S-format (fp16 hi / lo) activation."""
    _fields_ = [("ptr", P), ("sB", LL), ("sD", LL), ("sP", LL), ("sC8", LL), ("C8", I), ("parts", I), ("nb", I)]


_lib = None


def _built_id(path: str) -> str:
    """Build id baked into a libtstereo.so, read from the file bytes (no dlopen of a stale library)."""
    try:
        data = open(path, "rb").read()
    except OSError:
        return ""
    i = data.find(b"TSTEREO_BUILD_ID=")
    return data[i + 17:i + 33].decode(errors="replace") if i >= 0 else ""


class TStereoError(RuntimeError):
    """A libtstereo entry point returned a negative TSTEREO_E_* code."""


def load() -> C.CDLL:
    """Load libtstereo.so once and type every exported symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    default = "TSTEREO_LIB" not in os.environ
    if default:
        from . import build as _build
        want = _build.source_id()
        if not os.path.exists(LIB_PATH) or _built_id(LIB_PATH) != want:
            _build.build()                      # stale or missing: recompile in-tree (nvcc needs no GPU)
        if _built_id(LIB_PATH) != want:
            raise ImportError(f"{LIB_PATH} is stale and could not be rebuilt")
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m temporalstereo_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.tstereo_version() < 200:
        raise ImportError(f"libtstereo.so version {lib.tstereo_version()} too old")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; non-zero -> TStereoError(tstereo_last_error())."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise TStereoError(f"{name} failed ({rc}): {lib.tstereo_last_error().decode()}")
