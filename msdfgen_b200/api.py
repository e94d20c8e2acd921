"""Python mirror of the reference's generator interface, bound to the C ABI (include/msdfgen_b200.h) with ctypes.

The names follow msdfgen.h (/root/reference/msdfgen.h:47-56): ``generateSDF / generatePSDF / generateMSDF / generateMTSDF``
take a shape, an SDFTransformation and a generator config and fill a float bitmap; ``msdfErrorCorrection`` and the
``msdfFast*ErrorCorrection`` variants mirror core/msdf-error-correction.h.  Batched variants take a FlatShapes atlas.
All compute happens in the CUDA library; this module only marshals numpy arrays across the ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .flat import FlatShapes

SDF, PSDF, MSDF, MTSDF = 1, 2, 3, 4
CHANNELS = {SDF: 1, PSDF: 1, MSDF: 3, MTSDF: 4}
EC_DISABLED, EC_INDISCRIMINATE, EC_EDGE_PRIORITY, EC_EDGE_ONLY = 0, 1, 2, 3
DCM_DO_NOT_CHECK, DCM_CHECK_AT_EDGE, DCM_ALWAYS_CHECK = 0, 1, 2
DEFAULT_RATIO = 1.11111111111111111
FILL_NONZERO, FILL_ODD, FILL_POSITIVE, FILL_NEGATIVE = 0, 1, 2, 3     # FillRule, core/Scanline.h:9-14


@dataclass
class ErrorCorrectionConfig:
    """core/generator-config.h:13-46"""
    mode: int = EC_EDGE_PRIORITY
    distanceCheckMode: int = DCM_CHECK_AT_EDGE
    minDeviationRatio: float = DEFAULT_RATIO
    minImproveRatio: float = DEFAULT_RATIO


@dataclass
class MSDFGeneratorConfig:
    """core/generator-config.h:49-64 (GeneratorConfig + errorCorrection)"""
    overlapSupport: bool = True
    errorCorrection: ErrorCorrectionConfig = None
    # extension: the command line's scanline pass (main.cpp:1223-1297) run on the device between generation and error correction
    scanlinePass: bool = False
    fillRule: int = FILL_NONZERO
    sdfZeroValue: float = 0.5

    def __post_init__(self):
        if self.errorCorrection is None:
            self.errorCorrection = ErrorCorrectionConfig()

    def c(self, ec_buffer=None):
        cfg = _lib.Config(int(self.overlapSupport), self.errorCorrection.mode, self.errorCorrection.distanceCheckMode,
                          self.errorCorrection.minDeviationRatio, self.errorCorrection.minImproveRatio, None,
                          int(self.scanlinePass), int(self.fillRule), float(self.sdfZeroValue))
        if ec_buffer is not None:
            cfg.ec_buffer = ec_buffer.ctypes.data_as(C.POINTER(C.c_uint8))
        return cfg


GeneratorConfig = MSDFGeneratorConfig


@dataclass
class SDFTransformation:
    """Projection(scale, translate) + DistanceMapping(Range(lower, upper)) — core/SDFTransformation.h, DistanceMapping.cpp:13."""
    scale: tuple
    translate: tuple
    range_lower: float
    range_upper: float

    @staticmethod
    def from_array(a):
        a = np.asarray(a, dtype=np.float64)
        return SDFTransformation((a[0], a[1]), (a[2], a[3]), a[4], a[5])

    def c(self):
        lower, upper = np.float64(self.range_lower), np.float64(self.range_upper)
        return _lib.Transform(self.scale[0], self.scale[1], self.translate[0], self.translate[1], float(np.float64(1.0) / (upper - lower)), float(-lower))


def _transforms(xf):
    """Accepts SDFTransformation | array[6] | array[n,6] (sx, sy, tx, ty, range_lower, range_upper) -> (ctypes array, per_glyph)."""
    if isinstance(xf, SDFTransformation):
        return (_lib.Transform * 1)(xf.c()), 0
    a = np.asarray(xf, dtype=np.float64)
    if (a.ndim == 1 and a.size == 4) or (a.ndim == 2 and a.shape[1] == 4):
        # projection only (scanline passes): any distance range will do
        a = a.reshape(-1, 4)
        a = np.concatenate([a, np.tile([-1.0, 1.0], (len(a), 1))], axis=1)
    a = a.reshape(-1, 6)
    arr = (_lib.Transform * len(a))()
    for i, r in enumerate(a):
        arr[i] = SDFTransformation.from_array(r).c()
    return arr, (1 if len(a) > 1 else 0)


def _shapes_struct(s: FlatShapes):
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    return _lib.Shapes(s.n_glyphs, s.n_contours, s.n_segments, p(s.gbegin, C.c_uint32), p(s.cbegin, C.c_uint32), p(s.points, C.c_double),
                       p(s.orders, C.c_uint8), p(s.colors, C.c_uint8), p(s.ydown, C.c_uint8))


def _bitmap(a, ydown=False):
    h, w, n = a.shape
    assert a.dtype == np.float32
    if a.size:
        assert a.strides[2] == 4 and a.strides[1] == 4 * n and a.strides[0] % 4 == 0
    return _lib.Bitmap(a.ctypes.data_as(C.POINTER(C.c_float)), w, h, a.strides[0] // 4, int(ydown))


_BITMAP_DTYPE = np.dtype([("pixels", np.uint64), ("width", np.int32), ("height", np.int32), ("row_stride", np.int32), ("y_downward", np.int32)])


def _bitmaps(outputs, n, ydown=None):
    """b200msdf_bitmap[n] for a packed float32[n,h,w,N] block (vectorised) or a list of per-glyph arrays."""
    if isinstance(outputs, np.ndarray) and outputs.ndim == 4:
        assert outputs.dtype == np.float32 and outputs.shape[0] == n
        assert outputs.size == 0 or (outputs.strides[3] == 4 and outputs.strides[2] == 4 * outputs.shape[3])
        rec = np.empty(max(n, 1), dtype=_BITMAP_DTYPE)
        rec["pixels"][:n] = outputs.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(outputs.strides[0])
        rec["width"], rec["height"], rec["row_stride"] = outputs.shape[2], outputs.shape[1], outputs.strides[1] // 4
        rec["y_downward"] = 0
        if ydown is not None:
            rec["y_downward"][:n] = np.asarray(ydown, dtype=np.int32)
        arr = (_lib.Bitmap * max(n, 1)).from_buffer(rec)
        arr._keep = (rec, outputs)
        return arr
    bms = (_lib.Bitmap * max(n, 1))()
    for g in range(n):
        bms[g] = _bitmap(outputs[g], bool(ydown[g]) if ydown is not None else False)
    return bms


class Context:
    """One per thread and GPU (b200msdf_ctx): owns the stream and the pooled device / pinned buffers."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self._pinned = []
        _lib.check(self.lib.b200msdf_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            for p in self._pinned:
                self.lib.b200msdf_host_free(p)
            self._pinned = []
            self.lib.b200msdf_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream, own=False):
        """Run on the caller's cudaStream_t (0 = the CUDA default stream, which is torch's default); own=True returns to the context's stream."""
        _lib.check(self.lib.b200msdf_set_stream(self.h, C.c_void_p(cuda_stream), int(own)))

    def synchronize(self):
        _lib.check(self.lib.b200msdf_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.b200msdf_launch_count(self.h))

    def pinned(self, shape, dtype):
        """numpy array backed by pinned host memory from b200msdf_host_alloc; freed when the context is closed."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        ptr = C.c_void_p()
        _lib.check(self.lib.b200msdf_host_alloc(C.byref(ptr), max(count * dtype.itemsize, 1)))
        self._pinned.append(ptr)
        buf = (C.c_uint8 * max(count * dtype.itemsize, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def exact_root_calls(self, reset=False):
        """Runs of the literal (correctly rounded) cubic solver on this device since the last reset."""
        v = C.c_uint64(0)
        _lib.check(self.lib.b200msdf_debug_exact_root_calls(self.h, int(reset), C.byref(v)))
        return int(v.value)

    def bind_host_thread(self):
        """Pin the calling thread to the CPUs of the GPU's NUMA node (call before allocating pinned buffers); returns the CPU count."""
        n = self.lib.b200msdf_bind_host_thread(self.h)
        if n <= 0:
            raise RuntimeError(self.lib.b200msdf_last_error().decode())
        return n

    def release_pinned(self):
        """Frees every block handed out by pinned() (the arrays must no longer be used)."""
        self.synchronize()
        for p in self._pinned:
            self.lib.b200msdf_host_free(p)
        self._pinned = []

    # ---- reference-shaped single-shape calls ----------------------------------------------------------------------
    def generate(self, mode, output, shape: FlatShapes, transformation, config: MSDFGeneratorConfig = None, out_ydown=False,
                 ec_buffer=None):
        config = config or MSDFGeneratorConfig()
        assert shape.n_glyphs == 1 and output.shape[2] == CHANNELS[mode]
        xf, _ = _transforms(transformation)
        bm = _bitmap(output, out_ydown)
        s = _shapes_struct(shape)
        cfg = config.c(ec_buffer)
        _lib.check(self.lib.b200msdf_generate(self.h, mode, C.byref(s), xf, C.byref(bm), C.byref(cfg)))
        return output

    def generate_batch(self, mode, outputs, shapes: FlatShapes, transformations, config: MSDFGeneratorConfig = None, out_ydown=None):
        """outputs: float32[nGlyphs,h,w,N] (one block) or a list of per-glyph float32[h,w,N] arrays."""
        config = config or MSDFGeneratorConfig()
        xf, per = _transforms(transformations)
        n = shapes.n_glyphs
        bms = _bitmaps(outputs, n, out_ydown)
        s = _shapes_struct(shapes)
        cfg = config.c()
        _lib.check(self.lib.b200msdf_generate_batch(self.h, mode, C.byref(s), xf, per, bms, C.byref(cfg)))
        return outputs

    def generate_batch_u8(self, mode, outputs, shapes: FlatShapes, transformations, config: MSDFGeneratorConfig = None, out_ydown=None):
        """8-bit atlas output: outputs is a list of uint8[h,w,N] views (e.g. rectangles of one atlas page; row strides free)."""
        config = config or MSDFGeneratorConfig()
        xf, per = _transforms(transformations)
        n = shapes.n_glyphs
        bms = (_lib.BitmapU8 * max(n, 1))()
        for g in range(n):
            a = outputs[g]
            h, w, ch = a.shape
            assert a.dtype == np.uint8 and ch == CHANNELS[mode] and (a.size == 0 or (a.strides[2] == 1 and a.strides[1] == ch))
            bms[g] = _lib.BitmapU8(a.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, a.strides[0] if a.size else w * ch,
                                   int(bool(out_ydown[g])) if out_ydown is not None else 0)
        s = _shapes_struct(shapes)
        cfg = config.c()
        _lib.check(self.lib.b200msdf_generate_batch_u8(self.h, mode, C.byref(s), xf, per, bms, C.byref(cfg)))
        return outputs

    def error_correction(self, sdfs, shapes: FlatShapes, transformations, config: MSDFGeneratorConfig = None, sdf_ydown=None, ec_buffer=None):
        config = config or MSDFGeneratorConfig()
        xf, per = _transforms(transformations)
        n = shapes.n_glyphs
        bms = _bitmaps(sdfs, n, sdf_ydown)
        s = _shapes_struct(shapes)
        cfg = config.c(ec_buffer)
        _lib.check(self.lib.b200msdf_error_correction(self.h, sdfs[0].shape[2], C.byref(s), xf, per, bms, C.byref(cfg)))
        return sdfs

    def fast_error_correction(self, sdf, transformation, edge: bool, min_deviation_ratio=DEFAULT_RATIO):
        xf, _ = _transforms(transformation)
        bm = _bitmap(sdf)
        _lib.check(self.lib.b200msdf_fast_error_correction(self.h, sdf.shape[2], int(edge), xf, C.byref(bm), min_deviation_ratio))
        return sdf

    def sign_correction(self, sdfs, shapes: FlatShapes, transformations, sdf_zero_value=0.5, fill_rule=FILL_NONZERO, sdf_ydown=None):
        """distanceSignCorrection (core/rasterization.h:14-18) on every [h,w,N] field of the list, in place."""
        xf, per = _transforms(transformations)
        bms = _bitmaps(sdfs, shapes.n_glyphs, sdf_ydown)
        s = _shapes_struct(shapes)
        _lib.check(self.lib.b200msdf_sign_correction(self.h, sdfs[0].shape[2], C.byref(s), xf, per, bms, sdf_zero_value, fill_rule))
        return sdfs

    def rasterize(self, outputs, shapes: FlatShapes, transformations, fill_rule=FILL_NONZERO, out_ydown=None):
        """rasterize (core/rasterization.h:12) into every [h,w,1] output of the list."""
        xf, per = _transforms(transformations)
        bms = _bitmaps(outputs, shapes.n_glyphs, out_ydown)
        s = _shapes_struct(shapes)
        _lib.check(self.lib.b200msdf_rasterize(self.h, C.byref(s), xf, per, bms, fill_rule))
        return outputs

    def estimate_sdf_error(self, sdfs, shapes: FlatShapes, transformations, scanlines_per_row=4, fill_rule=FILL_NONZERO):
        """estimateSDFError (core/sdf-error-estimation.h:19) for every [h,w,N] field of the list -> float64[n_glyphs]."""
        xf, per = _transforms(transformations)
        bms = _bitmaps(sdfs, shapes.n_glyphs)
        s = _shapes_struct(shapes)
        out = np.zeros(shapes.n_glyphs, np.float64)
        _lib.check(self.lib.b200msdf_estimate_sdf_error(self.h, sdfs[0].shape[2], C.byref(s), xf, per, bms, int(scanlines_per_row), int(fill_rule),
                                                        out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def render_sdf(self, output, sdf, sdf_px_range=(0.0, 0.0), sd_threshold=0.5):
        """renderSDF (core/render-sdf.h:11-16): output float32 [oh,ow,NO] <- sdf float32 [h,w,NI]; sdf_px_range a number or (lower, upper)."""
        lo, hi = (-0.5 * sdf_px_range, 0.5 * sdf_px_range) if np.isscalar(sdf_px_range) else sdf_px_range     # Range(symmetricalWidth)
        bo, bs = _bitmap(output), _bitmap(sdf)
        _lib.check(self.lib.b200msdf_render_sdf(self.h, C.byref(bo), output.shape[2], C.byref(bs), sdf.shape[2], float(lo), float(hi), float(sd_threshold)))
        return output

    def simulate_8bit(self, bitmap):
        """simulate8bit (core/render-sdf.h:19-21), in place."""
        bm = _bitmap(bitmap)
        _lib.check(self.lib.b200msdf_simulate_8bit(self.h, C.byref(bm), bitmap.shape[2]))
        return bitmap

    # ---- HBM-resident batches --------------------------------------------------------------------------------------
    def upload(self, mode, shapes: FlatShapes, transformations, dims, out_ydown=None, origins=None):
        """origins: (x0, y0) pixel origin(s) of the output block(s) inside their full image (region batches, SDF/PSDF or EC off)."""
        return Batch(self, mode, shapes, transformations, dims, out_ydown, origins)


class Batch:
    """b200msdf_batch: shapes + geometry resident in HBM, with device output buffers."""

    def __init__(self, ctx: Context, mode, shapes: FlatShapes, transformations, dims, out_ydown=None, origins=None):
        self.ctx, self.mode, self.n = ctx, mode, shapes.n_glyphs
        xf, per = _transforms(transformations)
        d = np.ascontiguousarray(dims, dtype=np.int32).reshape(-1, 2)
        self.dims = d
        yd = np.ascontiguousarray(out_ydown, dtype=np.uint8) if out_ydown is not None else None
        s = _shapes_struct(shapes)
        self.h = C.c_void_p()
        if origins is not None:
            o = np.ascontiguousarray(origins, dtype=np.int32).reshape(-1, 2)
            _lib.check(ctx.lib.b200msdf_batch_upload_regions(ctx.h, mode, C.byref(s), xf, per, d.ctypes.data_as(C.POINTER(C.c_int32)), int(len(d) > 1),
                                                             o.ctypes.data_as(C.POINTER(C.c_int32)), int(len(o) > 1), C.byref(self.h)))
            return
        _lib.check(ctx.lib.b200msdf_batch_upload(ctx.h, mode, C.byref(s), xf, per, d.ctypes.data_as(C.POINTER(C.c_int32)), int(len(d) > 1),
                                                 yd.ctypes.data_as(C.POINTER(C.c_uint8)) if yd is not None else None, C.byref(self.h)))

    def launch(self, config: MSDFGeneratorConfig = None):
        cfg = (config or MSDFGeneratorConfig()).c()
        _lib.check(self.ctx.lib.b200msdf_batch_launch(self.ctx.h, self.h, C.byref(cfg)))

    def launch_stages(self, stages, config: MSDFGeneratorConfig = None):
        """stages: bit 0 windings/prepare, bit 1 distance field, bit 2 error correction, bit 3 scanline pass (per-kernel timing)."""
        cfg = (config or MSDFGeneratorConfig()).c()
        _lib.check(self.ctx.lib.b200msdf_batch_launch_stages(self.ctx.h, self.h, C.byref(cfg), stages))

    def sign_correction(self, sdf_zero_value=0.5, fill_rule=FILL_NONZERO):
        """distanceSignCorrection on the batch's most recent device result, in place."""
        _lib.check(self.ctx.lib.b200msdf_batch_sign_correction(self.ctx.h, self.h, sdf_zero_value, fill_rule))

    @property
    def output_floats(self):
        return int(self.ctx.lib.b200msdf_batch_output_floats(self.h))

    @property
    def device_output(self):
        return self.ctx.lib.b200msdf_batch_device_output(self.h)

    def glyph_offset(self, g):
        return int(self.ctx.lib.b200msdf_batch_glyph_offset(self.h, g))

    def download(self, dst=None):
        dst = np.empty(self.output_floats, dtype=np.float32) if dst is None else dst
        _lib.check(self.ctx.lib.b200msdf_batch_download(self.ctx.h, self.h, dst.ctypes.data_as(C.POINTER(C.c_float))))
        return dst

    def glyph(self, flat, g):
        w, h = self.dims[g if len(self.dims) > 1 else 0]
        n = CHANNELS[self.mode]
        o = self.glyph_offset(g)
        return flat[o:o + w * h * n].reshape(h, w, n)

    def free(self):
        if self.h:
            self.ctx.lib.b200msdf_batch_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- module-level functions named like the reference's -------------------------------------------------------------------
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def generateSDF(output, shape, transformation, config=None, **kw):
    return default_context().generate(SDF, output, shape, transformation, config, **kw)


def generatePSDF(output, shape, transformation, config=None, **kw):
    return default_context().generate(PSDF, output, shape, transformation, config, **kw)


generatePseudoSDF = generatePSDF


def generateMSDF(output, shape, transformation, config=None, **kw):
    return default_context().generate(MSDF, output, shape, transformation, config, **kw)


def generateMTSDF(output, shape, transformation, config=None, **kw):
    return default_context().generate(MTSDF, output, shape, transformation, config, **kw)


def msdfErrorCorrection(sdf, shape, transformation, config=None, **kw):
    return default_context().error_correction([sdf], shape, transformation, config, **kw)[0]


def distanceSignCorrection(sdf, shape, projection, sdfZeroValue=0.5, fillRule=FILL_NONZERO, **kw):
    """core/rasterization.h:14-18; `projection` may be a Projection-like (scale, translate) pair or an SDFTransformation."""
    return default_context().sign_correction([sdf], shape, projection, sdfZeroValue, fillRule, **kw)[0]


def rasterize(output, shape, projection, fillRule=FILL_NONZERO, **kw):
    """core/rasterization.h:12"""
    return default_context().rasterize([output], shape, projection, fillRule, **kw)[0]


def renderSDF(output, sdf, sdfPxRange=0.0, sdThreshold=0.5):
    """core/render-sdf.h:11-16"""
    return default_context().render_sdf(output, sdf, sdfPxRange, sdThreshold)


def estimateSDFError(sdf, shape, transformation, scanlinesPerRow, fillRule=FILL_NONZERO):
    """core/sdf-error-estimation.h:19-21"""
    return float(default_context().estimate_sdf_error([sdf], shape, transformation, scanlinesPerRow, fillRule)[0])


def simulate8bit(bitmap):
    """core/render-sdf.h:19-21"""
    return default_context().simulate_8bit(bitmap)


def msdfFastDistanceErrorCorrection(sdf, transformation, minDeviationRatio=DEFAULT_RATIO):
    return default_context().fast_error_correction(sdf, transformation, False, minDeviationRatio)


def msdfFastEdgeErrorCorrection(sdf, transformation, minDeviationRatio=DEFAULT_RATIO):
    return default_context().fast_error_correction(sdf, transformation, True, minDeviationRatio)
