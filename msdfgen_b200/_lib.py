"""Loader for the in-tree C-ABI library (msdfgen_b200/libmsdfgen_b200.so) — fails loudly when it is missing.

The library is built by ``__graft_entry__.build()`` (``make -C msdfgen_b200/csrc``).  There is no CPU or PyTorch fallback:
if the shared object is absent, or no CUDA device is present when a compute call is made, callers get an exception.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200MSDF_LIB selects an alternative build of the SAME library (kernel A/B experiments); default is the in-tree build.
LIB_PATH = os.environ.get("B200MSDF_LIB") or os.path.join(HERE, "libmsdfgen_b200.so")

# Every symbol include/msdfgen_b200.h declares (checked by tests/test_abi.py without a GPU).
EXPORTS = [
    "b200msdf_version", "b200msdf_last_error", "b200msdf_default_config", "b200msdf_create", "b200msdf_destroy",
    "b200msdf_set_stream", "b200msdf_synchronize", "b200msdf_launch_count", "b200msdf_bind_host_thread", "b200msdf_debug_exact_root_calls", "b200msdf_host_alloc", "b200msdf_host_free",
    "b200msdf_generate", "b200msdf_generate_batch", "b200msdf_generate_batch_u8", "b200msdf_error_correction", "b200msdf_fast_error_correction",
    "b200msdf_sign_correction", "b200msdf_rasterize", "b200msdf_batch_sign_correction", "b200msdf_render_sdf", "b200msdf_simulate_8bit", "b200msdf_estimate_sdf_error",
    "b200msdf_batch_upload", "b200msdf_batch_upload_regions", "b200msdf_batch_launch", "b200msdf_batch_launch_stages", "b200msdf_batch_device_output", "b200msdf_batch_device_stencil",
    "b200msdf_batch_output_floats", "b200msdf_batch_glyph_offset", "b200msdf_batch_download", "b200msdf_batch_free",
]


class Shapes(C.Structure):
    _fields_ = [("n_glyphs", C.c_int32), ("n_contours", C.c_int32), ("n_segments", C.c_int32),
                ("glyph_contour_begin", C.POINTER(C.c_uint32)), ("contour_segment_begin", C.POINTER(C.c_uint32)),
                ("points", C.POINTER(C.c_double)), ("orders", C.POINTER(C.c_uint8)), ("colors", C.POINTER(C.c_uint8)),
                ("y_downward", C.POINTER(C.c_uint8))]


class Transform(C.Structure):
    _fields_ = [("scale_x", C.c_double), ("scale_y", C.c_double), ("translate_x", C.c_double), ("translate_y", C.c_double),
                ("dm_scale", C.c_double), ("dm_translate", C.c_double)]


class Bitmap(C.Structure):
    _fields_ = [("pixels", C.POINTER(C.c_float)), ("width", C.c_int32), ("height", C.c_int32), ("row_stride", C.c_int32),
                ("y_downward", C.c_int32)]


class BitmapU8(C.Structure):
    _fields_ = [("pixels", C.POINTER(C.c_uint8)), ("width", C.c_int32), ("height", C.c_int32), ("row_stride", C.c_int32),
                ("y_downward", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("overlap_support", C.c_int32), ("ec_mode", C.c_int32), ("ec_distance_check_mode", C.c_int32),
                ("min_deviation_ratio", C.c_double), ("min_improve_ratio", C.c_double), ("ec_buffer", C.POINTER(C.c_uint8)),
                ("scanline_pass", C.c_int32), ("fill_rule", C.c_int32), ("sdf_zero_value", C.c_float)]


_lib = None


def load():
    """Returns the ctypes handle; raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C msdfgen_b200/csrc`). msdfgen_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
    lib.b200msdf_version.restype = i32
    lib.b200msdf_last_error.restype = C.c_char_p
    lib.b200msdf_default_config.argtypes = [C.POINTER(Config)]
    lib.b200msdf_default_config.restype = None
    lib.b200msdf_create.argtypes = [i32, C.POINTER(vp)]
    lib.b200msdf_destroy.argtypes = [vp]
    lib.b200msdf_destroy.restype = None
    lib.b200msdf_set_stream.argtypes = [vp, vp, i32]
    lib.b200msdf_synchronize.argtypes = [vp]
    lib.b200msdf_launch_count.argtypes = [vp]
    lib.b200msdf_launch_count.restype = u64
    lib.b200msdf_debug_exact_root_calls.argtypes = [vp, i32, C.POINTER(u64)]
    lib.b200msdf_debug_exact_root_calls.restype = i32
    lib.b200msdf_bind_host_thread.argtypes = [vp]
    lib.b200msdf_bind_host_thread.restype = i32
    lib.b200msdf_host_alloc.argtypes = [C.POINTER(vp), sz]
    lib.b200msdf_host_free.argtypes = [vp]
    lib.b200msdf_host_free.restype = None
    lib.b200msdf_generate.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), C.POINTER(Bitmap), C.POINTER(Config)]
    lib.b200msdf_generate_batch.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(Bitmap), C.POINTER(Config)]
    lib.b200msdf_generate_batch_u8.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(BitmapU8), C.POINTER(Config)]
    lib.b200msdf_error_correction.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(Bitmap), C.POINTER(Config)]
    lib.b200msdf_fast_error_correction.argtypes = [vp, i32, i32, C.POINTER(Transform), C.POINTER(Bitmap), C.c_double]
    lib.b200msdf_sign_correction.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(Bitmap), C.c_float, i32]
    lib.b200msdf_rasterize.argtypes = [vp, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(Bitmap), i32]
    lib.b200msdf_batch_sign_correction.argtypes = [vp, vp, C.c_float, i32]
    lib.b200msdf_render_sdf.argtypes = [vp, C.POINTER(Bitmap), i32, C.POINTER(Bitmap), i32, C.c_double, C.c_double, C.c_float]
    lib.b200msdf_simulate_8bit.argtypes = [vp, C.POINTER(Bitmap), i32]
    lib.b200msdf_batch_upload.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(C.c_int32), i32,
                                          C.POINTER(C.c_uint8), C.POINTER(vp)]
    lib.b200msdf_batch_upload_regions.argtypes = [vp, i32, C.POINTER(Shapes), C.POINTER(Transform), i32, C.POINTER(C.c_int32), i32,
                                                  C.POINTER(C.c_int32), i32, C.POINTER(vp)]
    lib.b200msdf_batch_launch.argtypes = [vp, vp, C.POINTER(Config)]
    lib.b200msdf_batch_launch_stages.argtypes = [vp, vp, C.POINTER(Config), C.c_uint]
    lib.b200msdf_batch_device_output.argtypes = [vp]
    lib.b200msdf_batch_device_output.restype = vp
    lib.b200msdf_batch_device_stencil.argtypes = [vp]
    lib.b200msdf_batch_device_stencil.restype = vp
    lib.b200msdf_batch_output_floats.argtypes = [vp]
    lib.b200msdf_batch_output_floats.restype = sz
    lib.b200msdf_batch_glyph_offset.argtypes = [vp, i32]
    lib.b200msdf_batch_glyph_offset.restype = sz
    lib.b200msdf_batch_download.argtypes = [vp, vp, C.POINTER(C.c_float)]
    lib.b200msdf_batch_free.argtypes = [vp, vp]
    lib.b200msdf_batch_free.restype = None
    _lib = lib
    return lib


class B200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise B200Error(f"msdfgen_b200 rc={rc}: {load().b200msdf_last_error().decode(errors='replace')}")
