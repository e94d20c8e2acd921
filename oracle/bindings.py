"""ctypes bindings for the test oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two checkers share one call surface (`Oracle`):

* ``kind="reference"`` — ``oracle/_ref/libmsdfgen_ref{,_omp}.so``: the UNMODIFIED reference core compiled from
  /root/reference by ``oracle/Makefile`` plus ``oracle/ref_wrapper.cpp`` (a pure marshalling wrapper).
* ``kind="port"`` — ``oracle/libmsdf_oracle.so``: the plain-C restatement ``oracle/msdf_oracle.c``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module.  Nothing under ``msdfgen_b200/`` does.

Shapes are exchanged flat (numpy): ``points`` float64[nSeg,8], ``orders`` uint8[nSeg] (1/2/3), ``colors`` uint8[nSeg],
``cbegin`` uint32[nContours+1]; a batch adds ``gbegin`` uint32[nGlyphs+1] (contour ranges).  Transforms are
``(scale_x, scale_y, translate_x, translate_y, range_lower, range_upper)``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference"

MODE = {"sdf": 1, "psdf": 2, "msdf": 3, "mtsdf": 4}
CHANNELS = {1: 1, 2: 1, 3: 3, 4: 4}

# ErrorCorrectionConfig defaults (/root/reference/core/generator-config.h:19-46, MSDFErrorCorrection.cpp:22-23)
EC_DISABLED, EC_INDISCRIMINATE, EC_EDGE_PRIORITY, EC_EDGE_ONLY = 0, 1, 2, 3
DCM_DO_NOT_CHECK, DCM_CHECK_AT_EDGE, DCM_ALWAYS_CHECK = 0, 1, 2
DEFAULT_RATIO = 1.11111111111111111


@dataclass
class Config:
    overlap: bool = True
    ec_mode: int = EC_EDGE_PRIORITY
    ec_dcm: int = DCM_CHECK_AT_EDGE
    min_deviation_ratio: float = DEFAULT_RATIO
    min_improve_ratio: float = DEFAULT_RATIO


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def build(which=("port", "ref")):
    """Compile the checkers (called by __graft_entry__.build()). The reference leg is skipped when /root/reference is absent."""
    if "port" in which:
        subprocess.run(["make", "-C", HERE, "port"], check=True, capture_output=True)
    if "ref" in which and os.path.isdir(REF_DIR):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_REF_LIBS = {"reference": "libmsdfgen_ref.so", "reference_omp": "libmsdfgen_ref_omp.so", "reference_gp": "libmsdfgen_ref_gp.so"}


def ref_available(omp=False, kind=None):
    kind = kind or ("reference_omp" if omp else "reference")
    return os.path.exists(os.path.join(HERE, "_ref", _REF_LIBS[kind]))


def port_available():
    return os.path.exists(os.path.join(HERE, "libmsdf_oracle.so"))


class Oracle:
    """One checker instance. kind: 'reference' (single thread) | 'reference_omp' (the reference's own row-parallel OpenMP
    build) | 'reference_gp' (single-threaded reference core under the wrapper's outer OpenMP loop over glyphs) | 'port'."""

    def __init__(self, kind="reference"):
        self.kind = kind
        if kind == "port":
            path = os.path.join(HERE, "libmsdf_oracle.so")
            if not os.path.exists(path):
                build(("port",))
            self.lib = C.CDLL(path)
            self.pfx = "oracle_"
        else:
            path = os.path.join(HERE, "_ref", _REF_LIBS[kind])
            if not os.path.exists(path):
                if not os.path.isdir(REF_DIR):
                    raise FileNotFoundError(f"{path} missing and {REF_DIR} absent: the reference oracle cannot be built here")
                build(("ref",))
            self.lib = C.CDLL(path)
            self.pfx = "ref_"
        self.is_ref = self.pfx == "ref_"

    # -- helpers -----------------------------------------------------------------------------------------------------
    def _xf(self, xf):
        """(sx, sy, tx, ty, lower, upper) -> what the library wants: the reference takes the Range, the port the mapping."""
        xf = _f64(xf).reshape(-1, 6).copy()
        if not self.is_ref:
            lower, upper = xf[:, 4].copy(), xf[:, 5].copy()
            xf[:, 4] = 1.0 / (upper - lower)      # DistanceMapping(Range): scale, /root/reference/core/DistanceMapping.cpp:13
            xf[:, 5] = -lower                     # translate
        return xf

    def info(self):
        if not self.is_ref:
            return {"kind": "port", "omp": False, "threads": 1}
        omp, thr = C.c_int(0), C.c_int(0)
        ver = self.lib.ref_info(C.byref(omp), C.byref(thr))
        return {"kind": "reference", "version": ver, "omp": bool(omp.value), "omp_level": omp.value, "threads": thr.value}

    def set_threads(self, n):
        """OpenMP team size of this library's parallel regions (no-op for the single-threaded builds); returns the size in effect."""
        if not self.is_ref:
            return 1
        self.lib.ref_set_threads.restype = C.c_int
        return int(self.lib.ref_set_threads(C.c_int(int(n))))

    # -- generators --------------------------------------------------------------------------------------------------
    def generate(self, mode, cbegin, points, orders, colors, xf, w, h, cfg: Config = Config(), shape_ydown=False,
                 out_ydown=False, want_stencil=False):
        mode = MODE.get(mode, mode)
        n = CHANNELS[mode]
        cbegin, points, orders, colors = _u32(cbegin), _f64(points), _u8(orders), _u8(colors)
        out = np.full((h, w, n), np.nan, dtype=np.float32)
        st = np.zeros((h, w), dtype=np.uint8) if want_stencil else None
        x = self._xf(xf)
        fn = getattr(self.lib, self.pfx + "generate")
        fn.restype = C.c_int
        rc = fn(C.c_int(mode), C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8),
                _p(colors, C.c_uint8), C.c_int(int(shape_ydown)), _p(x, C.c_double), _p(out, C.c_float), C.c_int(w), C.c_int(h),
                C.c_int(w * n), C.c_int(int(out_ydown)), C.c_int(int(cfg.overlap)), C.c_int(cfg.ec_mode), C.c_int(cfg.ec_dcm),
                C.c_double(cfg.min_deviation_ratio), C.c_double(cfg.min_improve_ratio), _p(st, C.c_uint8))
        if rc:
            raise RuntimeError(f"{self.pfx}generate rc={rc}")
        return (out, st) if want_stencil else out

    def generate_batch(self, mode, gbegin, cbegin, points, orders, colors, xf, w, h, cfg: Config = Config(), ydown=None,
                       glyph_parallel=False):
        """Loop of generate* over same-size glyphs; returns (out[nGlyphs,h,w,N], seconds inside the generator loop)."""
        import time
        mode = MODE.get(mode, mode)
        n = CHANNELS[mode]
        gbegin, cbegin, points, orders, colors = _u32(gbegin), _u32(cbegin), _f64(points), _u8(orders), _u8(colors)
        ng = len(gbegin) - 1
        out = np.empty((ng, h, w, n), dtype=np.float32)
        x = self._xf(xf)
        per = 1 if len(x) > 1 else 0
        yd = _u8(ydown) if ydown is not None else None
        fn = getattr(self.lib, self.pfx + "generate_batch")
        fn.restype = C.c_int
        args = [C.c_int(mode), C.c_int(ng), _p(gbegin, C.c_uint32), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8),
                _p(colors, C.c_uint8), _p(yd, C.c_uint8), _p(x, C.c_double), C.c_int(per), _p(out, C.c_float), C.c_int(w), C.c_int(h),
                C.c_int(int(cfg.overlap)), C.c_int(cfg.ec_mode), C.c_int(cfg.ec_dcm), C.c_double(cfg.min_deviation_ratio),
                C.c_double(cfg.min_improve_ratio)]
        if self.is_ref:
            secs = C.c_double(0)
            rc = fn(*args, C.c_int(int(glyph_parallel)), C.byref(secs))
            dt = secs.value
        else:
            t0 = time.perf_counter()
            rc = fn(*args)
            dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError(f"{self.pfx}generate_batch rc={rc}")
        return out, dt

    def error_correction(self, sdf, cbegin, points, orders, colors, xf, cfg: Config = Config(), shape_ydown=False, sdf_ydown=False,
                         want_stencil=False):
        """msdfErrorCorrection on a copy of sdf[h,w,N]; returns corrected copy (and the stencil)."""
        sdf = np.array(sdf, dtype=np.float32, order="C", copy=True)
        h, w, n = sdf.shape
        cbegin, points, orders, colors = _u32(cbegin), _f64(points), _u8(orders), _u8(colors)
        st = np.zeros((h, w), dtype=np.uint8) if want_stencil else None
        x = self._xf(xf)
        fn = getattr(self.lib, self.pfx + "error_correction")
        fn.restype = C.c_int
        args = [C.c_int(n), C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8),
                _p(colors, C.c_uint8), C.c_int(int(shape_ydown)), _p(x, C.c_double), _p(sdf, C.c_float), C.c_int(w), C.c_int(h),
                C.c_int(w * n), C.c_int(int(sdf_ydown)), C.c_int(int(cfg.overlap)), C.c_int(cfg.ec_mode), C.c_int(cfg.ec_dcm),
                C.c_double(cfg.min_deviation_ratio), C.c_double(cfg.min_improve_ratio)]
        if not self.is_ref:
            args.append(C.c_int(0))
        args.append(_p(st, C.c_uint8))
        rc = fn(*args)
        if rc:
            raise RuntimeError(f"{self.pfx}error_correction rc={rc}")
        return (sdf, st) if want_stencil else sdf

    def fast_error_correction(self, sdf, xf, edge: bool, min_deviation_ratio=DEFAULT_RATIO):
        """msdfFastDistanceErrorCorrection (edge=False) / msdfFastEdgeErrorCorrection (edge=True)."""
        sdf = np.array(sdf, dtype=np.float32, order="C", copy=True)
        h, w, n = sdf.shape
        x = self._xf(xf)
        if self.is_ref:
            self.lib.ref_fast_error_correction.restype = C.c_int
            rc = self.lib.ref_fast_error_correction(C.c_int(n), C.c_int(int(edge)), _p(x, C.c_double), _p(sdf, C.c_float), C.c_int(w),
                                                    C.c_int(h), C.c_int(w * n), C.c_double(min_deviation_ratio))
        else:
            cb = np.zeros(1, dtype=np.uint32)
            self.lib.oracle_error_correction.restype = C.c_int
            rc = self.lib.oracle_error_correction(
                C.c_int(n), C.c_int(0), _p(cb, C.c_uint32), None, None, None, C.c_int(0), _p(x, C.c_double), _p(sdf, C.c_float),
                C.c_int(w), C.c_int(h), C.c_int(w * n), C.c_int(0), C.c_int(1), C.c_int(EC_EDGE_ONLY if edge else EC_INDISCRIMINATE),
                C.c_int(0), C.c_double(min_deviation_ratio), C.c_double(DEFAULT_RATIO), C.c_int(1), None)
        if rc:
            raise RuntimeError("fast_error_correction failed")
        return sdf

    def sign_correction(self, sdf, cbegin, points, orders, xf, zero_value=0.5, fill_rule=0, shape_ydown=False, sdf_ydown=False):
        """distanceSignCorrection (core/rasterization.h) on a copy of sdf[h,w,N]; fill_rule 0 NONZERO 1 ODD 2 POSITIVE 3 NEGATIVE."""
        sdf = np.array(sdf, dtype=np.float32, order="C", copy=True)
        h, w, n = sdf.shape
        cbegin, points, orders = _u32(cbegin), _f64(points), _u8(orders)
        x = _f64(xf).reshape(-1)[:4].copy()
        fn = getattr(self.lib, self.pfx + "sign_correction")
        fn.restype = C.c_int
        rc = fn(C.c_int(n), C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8), None,
                C.c_int(int(shape_ydown)), _p(x, C.c_double), _p(sdf, C.c_float), C.c_int(w), C.c_int(h), C.c_int(w * n),
                C.c_int(int(sdf_ydown)), C.c_float(zero_value), C.c_int(fill_rule))
        if rc:
            raise RuntimeError("sign_correction failed")
        return sdf

    def rasterize(self, cbegin, points, orders, xf, w, h, fill_rule=0, shape_ydown=False, out_ydown=False):
        """rasterize (core/rasterization.h) -> float32 [h,w,1] of 0/1."""
        out = np.zeros((h, w, 1), np.float32)
        cbegin, points, orders = _u32(cbegin), _f64(points), _u8(orders)
        x = _f64(xf).reshape(-1)[:4].copy()
        fn = getattr(self.lib, self.pfx + "rasterize")
        fn.restype = C.c_int
        rc = fn(C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8), C.c_int(int(shape_ydown)),
                _p(x, C.c_double), _p(out, C.c_float), C.c_int(w), C.c_int(h), C.c_int(int(out_ydown)), C.c_int(fill_rule))
        if rc:
            raise RuntimeError("rasterize failed")
        return out

    def render_sdf(self, sdf, ow, oh, out_channels, px_range=(0.0, 0.0), threshold=0.5):
        """renderSDF (core/render-sdf.h:11-16): sdf float32 [h,w,N] -> float32 [oh,ow,out_channels]."""
        sdf = np.ascontiguousarray(sdf, dtype=np.float32)
        sh, sw, n = sdf.shape
        out = np.full((oh, ow, out_channels), np.nan, np.float32)
        fn = getattr(self.lib, self.pfx + "render_sdf")
        fn.restype = C.c_int
        rc = fn(C.c_int(out_channels), _p(out, C.c_float), C.c_int(ow), C.c_int(oh), C.c_int(n), _p(sdf, C.c_float), C.c_int(sw), C.c_int(sh),
                C.c_double(px_range[0]), C.c_double(px_range[1]), C.c_float(threshold))
        if rc:
            raise RuntimeError("render_sdf: unsupported channel combination")
        return out

    def estimate_sdf_error(self, sdf, cbegin, points, orders, xf, scanlines_per_row=4, fill_rule=0, shape_ydown=False):
        """estimateSDFError of the compiled reference (core/sdf-error-estimation.h:19-21) for one [h,w,N] field."""
        assert self.is_ref
        sdf = np.ascontiguousarray(sdf, dtype=np.float32)
        h, w, n = sdf.shape
        cbegin, points, orders = _u32(cbegin), _f64(points), _u8(orders)
        x = _f64(np.asarray(xf, dtype=np.float64)[:4])
        fn = self.lib.ref_estimate_sdf_error
        fn.restype = C.c_double
        return float(fn(C.c_int(n), C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8),
                        C.c_int(int(shape_ydown)), _p(x, C.c_double), _p(sdf, C.c_float), C.c_int(w), C.c_int(h), C.c_int(scanlines_per_row), C.c_int(fill_rule)))

    def simulate_8bit(self, bitmap):
        """simulate8bit (core/render-sdf.h:19-21) on a copy."""
        b = np.array(bitmap, dtype=np.float32, order="C", copy=True)
        h, w, n = b.shape
        fn = getattr(self.lib, self.pfx + "simulate_8bit")
        fn.restype = None
        fn(_p(b, C.c_float), C.c_int(w), C.c_int(h), C.c_int(n))
        return b

    # -- pieces ------------------------------------------------------------------------------------------------------
    def distance(self, mode, overlap, cbegin, points, orders, colors, pts):
        """Raw fp64 shape distance (oneShotDistance) at arbitrary points -> float64[nPts, N]."""
        mode = MODE.get(mode, mode)
        n = CHANNELS[mode]
        cbegin, points, orders, colors, pts = _u32(cbegin), _f64(points), _u8(orders), _u8(colors), _f64(pts).reshape(-1, 2)
        out = np.empty((len(pts), n), dtype=np.float64)
        fn = getattr(self.lib, "ref_oneshot_distance" if self.is_ref else "oracle_distance")
        fn.restype = C.c_int
        rc = fn(C.c_int(mode), C.c_int(int(overlap)), C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double),
                _p(orders, C.c_uint8), _p(colors, C.c_uint8), C.c_int(len(pts)), _p(pts, C.c_double), _p(out, C.c_double))
        if rc:
            raise RuntimeError("distance failed")
        return out

    def segment_distances(self, points, orders, pts):
        """EdgeSegment::signedDistance of every segment at every point -> float64[nPts, nSeg, 3] = (distance, dot, param)."""
        points, orders, pts = _f64(points), _u8(orders), _f64(pts).reshape(-1, 2)
        out = np.empty((len(pts), len(orders), 3), dtype=np.float64)
        fn = getattr(self.lib, self.pfx + "segment_distances")
        fn.restype = C.c_int
        fn(C.c_int(len(orders)), _p(points, C.c_double), _p(orders, C.c_uint8), C.c_int(len(pts)), _p(pts, C.c_double), _p(out, C.c_double))
        return out

    def solve_cubic(self, a, b, c, d):
        x = (C.c_double * 3)()
        fn = getattr(self.lib, self.pfx + "solve_cubic")
        fn.restype = C.c_int
        n = fn(x, C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(d))
        return n, [x[i] for i in range(max(n, 0))]

    def solve_quadratic(self, a, b, c):
        x = (C.c_double * 2)()
        fn = getattr(self.lib, self.pfx + "solve_quadratic")
        fn.restype = C.c_int
        n = fn(x, C.c_double(a), C.c_double(b), C.c_double(c))
        return n, [x[i] for i in range(max(n, 0))]

    def windings(self, cbegin, points, orders):
        cbegin, points, orders = _u32(cbegin), _f64(points), _u8(orders)
        out = np.zeros(len(cbegin) - 1, dtype=np.int32)
        fn = getattr(self.lib, self.pfx + "windings")
        fn(C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8), _p(out, C.c_int))
        return out

    def float_to_byte(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty(a.shape, dtype=np.uint8)
        fn = getattr(self.lib, self.pfx + "pixel_float_to_byte")
        fn.restype = None
        fn(_p(a, C.c_float), _p(out, C.c_uint8), C.c_size_t(a.size))
        return out

    # -- reference-only helpers (shape preparation by the reference's own code) --------------------------------------------
    def prepare(self, cbegin, points, orders, colors=None, normalize=True, coloring=1, angle=3.0, seed=0, ydown=False):
        """Shape::normalize() + edgeColoring{Simple,InkTrap,ByDistance} by the reference; returns (cbegin, points, orders, colors)."""
        assert self.is_ref, "prepare() exists only on the compiled reference"
        cbegin, points, orders = _u32(cbegin), _f64(points), _u8(orders)
        colors = _u8(colors) if colors is not None else np.full(len(orders), 7, dtype=np.uint8)
        cap_s, cap_c = 3 * len(orders) + 16 * len(cbegin) + 16, len(cbegin) + 4
        ocb = np.zeros(cap_c + 1, dtype=np.uint32)
        opts = np.zeros((cap_s, 8), dtype=np.float64)
        oord = np.zeros(cap_s, dtype=np.uint8)
        ocol = np.zeros(cap_s, dtype=np.uint8)
        nc = C.c_int(0)
        self.lib.ref_prepare.restype = C.c_int
        ns = self.lib.ref_prepare(C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8),
                                  _p(colors, C.c_uint8), C.c_int(int(ydown)), C.c_int(int(normalize)), C.c_int(coloring), C.c_double(angle),
                                  C.c_ulonglong(seed), C.c_int(cap_s), C.c_int(cap_c), _p(ocb, C.c_uint32), _p(opts, C.c_double),
                                  _p(oord, C.c_uint8), _p(ocol, C.c_uint8), C.byref(nc))
        if ns < 0:
            raise RuntimeError("ref_prepare capacity")
        return ocb[: nc.value + 1].copy(), opts[:ns].copy(), oord[:ns].copy(), ocol[:ns].copy()

    def validate(self, cbegin, points, orders):
        assert self.is_ref
        cbegin, points, orders = _u32(cbegin), _f64(points), _u8(orders)
        return bool(self.lib.ref_validate(C.c_int(len(cbegin) - 1), _p(cbegin, C.c_uint32), _p(points, C.c_double), _p(orders, C.c_uint8)))

    def read_shape_description(self, text):
        """readShapeDescription(const char *) -> (cbegin, points, orders, colors, ydown, colors_specified)."""
        assert self.is_ref
        cap_s, cap_c = 4096, 512
        ocb = np.zeros(cap_c + 1, dtype=np.uint32)
        opts = np.zeros((cap_s, 8), dtype=np.float64)
        oord = np.zeros(cap_s, dtype=np.uint8)
        ocol = np.zeros(cap_s, dtype=np.uint8)
        nc, yd, cs = C.c_int(0), C.c_int(0), C.c_int(0)
        self.lib.ref_read_shape_description.restype = C.c_int
        ns = self.lib.ref_read_shape_description(text.encode(), C.c_int(cap_s), C.c_int(cap_c), _p(ocb, C.c_uint32), _p(opts, C.c_double),
                                                 _p(oord, C.c_uint8), _p(ocol, C.c_uint8), C.byref(nc), C.byref(yd), C.byref(cs))
        if ns < 0:
            raise ValueError(f"readShapeDescription failed ({ns})")
        return ocb[: nc.value + 1].copy(), opts[:ns].copy(), oord[:ns].copy(), ocol[:ns].copy(), bool(yd.value), bool(cs.value)
