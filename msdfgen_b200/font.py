"""TrueType glyph ingest (msdfgen_b200/host/font_ingest.cpp, the FreeType-free counterpart of msdfgen's ext/import-font):
a font's glyphs as ONE flat batch, ready for prep.prepare() and Context.generate_batch()."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .flat import FlatShapes
from .prep import host_lib, _p

SCALING = {None: 0, "none": 0, "em": 1, "legacy": 2}


class Font:
    """A TrueType font held by the host library.  Mirrors loadFontData / getFontMetrics / getGlyphIndex / loadGlyph of
    /root/reference/ext/import-font.h:64-93; glyphs come back as flat arrays instead of Shape objects."""

    def __init__(self, source):
        data = source if isinstance(source, (bytes, bytearray)) else open(source, "rb").read()
        self._lib = host_lib()
        self._lib.b200host_font_open.restype = C.c_void_p
        self._lib.b200host_font_glyphs_begin.restype = C.c_void_p
        buf = np.frombuffer(bytes(data), np.uint8)
        self._h = self._lib.b200host_font_open(_p(buf, C.c_uint8), C.c_int(len(buf)))
        if not self._h:
            raise ValueError("not a TrueType font with glyf outlines (CFF / collections / malformed tables are not supported)")
        info = np.zeros(8, np.int32)
        self._lib.b200host_font_info(C.c_void_p(self._h), _p(info, C.c_int))
        (self.n_glyphs, self.units_per_em, self.ascender, self.descender, self.line_height, self.underline_position,
         self.underline_thickness) = (int(v) for v in info[:7])
        self.has_kerning = bool(info[7])

    def close(self):
        if self._h:
            self._lib.b200host_font_close(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        self.close()

    def scale(self, scaling="em"):
        return {0: 1.0, 1: 1.0 / self.units_per_em, 2: 1 / 64.0}[SCALING[scaling]]

    def glyph_indices(self, unicodes):
        """getGlyphIndex for every code point (0 = not in the font's Unicode character map)."""
        u = np.ascontiguousarray(unicodes, np.uint32)
        out = np.zeros(len(u), np.uint32)
        self._lib.b200host_font_glyph_indices(C.c_void_p(self._h), C.c_int(len(u)), _p(u, C.c_uint32), _p(out, C.c_uint32))
        return out

    def load_glyphs(self, glyphs, scaling="em"):
        """loadGlyph for every glyph index -> (FlatShapes with one glyph per index, all edges white; advances; loaded flags)."""
        g = np.ascontiguousarray(glyphs, np.uint32)
        nc, ns = C.c_int(0), C.c_int(0)
        h = self._lib.b200host_font_glyphs_begin(C.c_void_p(self._h), C.c_int(len(g)), _p(g, C.c_uint32), C.c_int(SCALING[scaling]),
                                                 C.byref(nc), C.byref(ns))
        if not h:
            raise RuntimeError("b200host_font_glyphs_begin failed")
        try:
            gb = np.zeros(len(g) + 1, np.uint32)
            cb = np.zeros(nc.value + 1, np.uint32)
            pts = np.zeros((ns.value, 8), np.float64)
            od = np.zeros(ns.value, np.uint8)
            cl = np.zeros(ns.value, np.uint8)
            adv = np.zeros(len(g), np.float64)
            ok = np.zeros(len(g), np.uint8)
            self._lib.b200host_font_glyphs_fetch(C.c_void_p(h), _p(gb, C.c_uint32), _p(cb, C.c_uint32), _p(pts, C.c_double), _p(od, C.c_uint8),
                                                 _p(cl, C.c_uint8), _p(adv, C.c_double), _p(ok, C.c_uint8))
        finally:
            self._lib.b200host_font_glyphs_free(C.c_void_p(h))
        return FlatShapes(gb, cb, pts, od, cl, np.zeros(len(g), np.uint8)), adv, ok.astype(bool)

    def kerning(self, left, right):
        v = C.c_int(0)
        return int(v.value) if self._lib.b200host_font_kerning(C.c_void_p(self._h), C.c_uint32(left), C.c_uint32(right), C.byref(v)) else 0
