"""Host-side shape preparation (Shape::normalize, edgeColoringSimple / InkTrap, Contour::winding, Shape::getBounds) on flat
batches, via libmsdfgen_b200_host.so — the C++ mirror of the reference's data model (msdfgen_b200/host/msdfgen.h).

This is what every MSDF caller runs before generate* (/root/reference/README.md:106-108, main.cpp:1125,1255); it is host
work (O(edges)), not part of the device path, and it has no dependency on the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .flat import FlatShapes

HERE = os.path.dirname(os.path.abspath(__file__))
# B200MSDF_HOST_LIB selects an alternative build of the SAME host library (development); default is the in-tree build.
HOST_LIB_PATH = os.environ.get("B200MSDF_HOST_LIB") or os.path.join(HERE, "libmsdfgen_b200_host.so")
_host = None

COLORING = {None: 0, "none": 0, "simple": 1, "inktrap": 2}


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build() (make -C msdfgen_b200/host)")
        # the host library links the C-ABI library by $ORIGIN rpath
        _host = C.CDLL(HOST_LIB_PATH)
    return _host


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def prepare(shapes: FlatShapes, normalize=True, coloring="simple", angle_threshold=3.0, seed=0) -> FlatShapes:
    """shape.normalize(); edgeColoringSimple(shape, angle_threshold, seed) for every glyph of the batch."""
    lib = host_lib()
    h = C.c_void_p()
    nc, ns = C.c_int(0), C.c_int(0)
    rc = lib.b200host_prepare_begin(C.c_int(shapes.n_glyphs), _p(shapes.gbegin, C.c_uint32), _p(shapes.cbegin, C.c_uint32),
                                    _p(shapes.points, C.c_double), _p(shapes.orders, C.c_uint8), _p(shapes.colors, C.c_uint8),
                                    _p(shapes.ydown, C.c_uint8), C.c_int(int(normalize)), C.c_int(COLORING[coloring]),
                                    C.c_double(angle_threshold), C.c_ulonglong(seed), C.byref(h), C.byref(nc), C.byref(ns))
    if rc:
        raise RuntimeError("b200host_prepare_begin failed")
    try:
        gb = np.zeros(shapes.n_glyphs + 1, np.uint32)
        cb = np.zeros(nc.value + 1, np.uint32)
        # b200host_prepare_fetch writes every element; touching the pages here (one thread) keeps the worker threads of the
        # fetch from serialising on first-touch page faults
        pts = np.empty((ns.value, 8), np.float64)
        pts.fill(0.0)
        od = np.empty(ns.value, np.uint8)
        cl = np.empty(ns.value, np.uint8)
        lib.b200host_prepare_fetch(h, _p(gb, C.c_uint32), _p(cb, C.c_uint32), _p(pts, C.c_double), _p(od, C.c_uint8), _p(cl, C.c_uint8))
    finally:
        lib.b200host_prepare_free(h)
    return FlatShapes(gb, cb, pts, od, cl, shapes.ydown.copy())


def windings(shapes: FlatShapes):
    out = np.zeros(shapes.n_contours, np.int32)
    host_lib().b200host_windings(C.c_int(shapes.n_contours), _p(shapes.cbegin, C.c_uint32), _p(shapes.points, C.c_double),
                                 _p(shapes.orders, C.c_uint8), _p(out, C.c_int))
    return out


def validate(shapes: FlatShapes):
    return bool(host_lib().b200host_validate(C.c_int(shapes.n_contours), _p(shapes.cbegin, C.c_uint32), _p(shapes.points, C.c_double),
                                             _p(shapes.orders, C.c_uint8)))


def bounds(shapes: FlatShapes, border=0.0, miter_limit=0.0, polarity=0):
    """Shape::getBounds over all contours of the batch -> (l, b, r, t)."""
    out = np.zeros(4, np.float64)
    host_lib().b200host_bounds(C.c_int(shapes.n_contours), _p(shapes.cbegin, C.c_uint32), _p(shapes.points, C.c_double),
                               _p(shapes.orders, C.c_uint8), C.c_double(border), C.c_double(miter_limit), C.c_int(polarity), _p(out, C.c_double))
    return out
