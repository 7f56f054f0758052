"""ctypes binding of libstemb200.so (the C ABI declared in include/stemb200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.  The library is built
in-tree by ``__graft_entry__.build()`` (or ``make -C spatiotemporalentropymodel_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# STEMB200_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("STEMB200_LIB") or os.path.join(_HERE, "libstemb200.so")

DT_F16, DT_F32 = 0, 1
EPI_LINEAR, EPI_SFT, EPI_ADD = 0, 1, 2


class ConvDesc(C.Structure):
    """Mirror of ``stemb200_conv_desc`` (include/stemb200.h)."""

    _fields_ = [
        ("batch", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32), ("n_src", C.c_int32),
        ("c_in", C.c_int32 * 3), ("c_out", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("stride", C.c_int32), ("transposed", C.c_int32), ("tap_mask", C.c_uint32), ("epilogue", C.c_int32),
        ("lrelu_slope", C.c_float), ("out_dtype", C.c_int32), ("sq_scale", C.c_float),
        ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("direct_store", C.c_int32), ("row_taps", C.c_int32),
    ]


class ArDesc(C.Structure):
    """Mirror of ``stemb200_ar_desc``."""

    _fields_ = [("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("l1", C.c_int32),
                ("l2", C.c_int32), ("slope", C.c_float), ("n_scales", C.c_int32), ("scale_bound", C.c_float)]


class StemLibError(RuntimeError):
    pass


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/stemb200.h declares
SIGNATURES = {
    "stemb200_version": (C.c_char_p, []),
    "stemb200_last_error": (C.c_char_p, []),
    "stemb200_launch_count": (C.c_uint64, []),
    "stemb200_conv2d_packed_k": (_i64, [C.POINTER(ConvDesc)]),
    "stemb200_conv2d_pack_weight": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp]),
    "stemb200_conv2d_fwd": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp]),
    "stemb200_conv2d_gdn_fwd": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_vp), _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "stemb200_conv2d_gdn_last_fwd": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _vp]),
    "stemb200_conv2d_gc_fwd": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(_vp), _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _vp,
                                         _vp, _vp]),
    "stemb200_synthesis_col_index": (C.c_int, [_i32, _i32, _i32]),
    "stemb200_synthesis_col2im": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32,
                                            _vp]),
    "stemb200_synthesis_col2im_u8": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32,
                                               _vp]),
    "stemb200_nchw_f32_to_nhwc_f16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_nhwc_f16_to_nchw_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_nhwc_f32_to_nchw_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_im2col_k5s2_c3": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_frame_to_nhwc8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_frame_u8_to_nhwc8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_frame_to_nhwc4": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_frame_u8_to_nhwc4": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_conv_first_gdn_fwd": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _vp, _vp]),
    "stemb200_im2col_k3s1_c4": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "stemb200_im2col_k3s1_c4_u8": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "stemb200_im2col_k5s2_c3_u8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_avgpool_nhwc_f16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_qmap_pool": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "stemb200_latent_stage": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "stemb200_gaussian_conditional_fwd": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _f32,
                                                    _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stemb200_gaussian_conditional_fwd_cond32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _f32, _f32,
                                                           _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stemb200_gaussian_conditional_flat": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _f32, _f32, _vp, _vp, _vp,
                                                     _vp, _vp, _vp]),
    "stemb200_entropy_bottleneck_fwd": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp,
                                                  _vp]),
    "stemb200_synthesis_tail": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "stemb200_synthesis_tail_u8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "stemb200_cast_f16_to_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "stemb200_ar_packed_floats": (_i64, [C.POINTER(ArDesc)]),
    "stemb200_ar_workspace_bytes": (_i64, [C.POINTER(ArDesc)]),
    "stemb200_ar_encode": (C.c_int, [C.POINTER(ArDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stemb200_ar_decode": (C.c_int, [C.POINTER(ArDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp,
                                     _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "stemb200_rans_encode_host": (_i64, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _i64]),
    "stemb200_rans_decode_host": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "stemb200_pmf_to_quantized_cdf_host": (C.c_int, [C.POINTER(C.c_float), _i32, _i32, C.POINTER(C.c_int32)]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libstemb200.so and bind every declared symbol. Raises StemLibError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StemLibError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().stemb200_last_error().decode("utf-8", "replace")
        raise StemLibError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().stemb200_launch_count())
