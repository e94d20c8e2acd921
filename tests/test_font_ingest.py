"""TrueType glyph ingest (msdfgen_b200/host/font_ingest.cpp): the FreeType-free counterpart of msdfgen's
ext/import-font.  The image ships one TrueType font (an icon font inside the CUDA toolkit's HTML documentation: 676 outlines of
up to 39 contours and 364 segments, abutting contours, shared edges) and Pillow bundles a FreeType of its own — so the ingest is
checked against FreeType's rendering of the same glyphs, and the real outlines then go through the host-compiled texel path
(tests/devcore) against the compiled reference.  tests/test_zz_real_font.py is the GPU leg."""
import glob

import numpy as np
import pytest

from msdfgen_b200 import prep
from msdfgen_b200.flat import FlatShapes

CANDIDATES = sorted(glob.glob("/usr/local/cuda*/compute-sanitizer/docs/_static/css/fonts/fontawesome-webfont.ttf") +
                    glob.glob("/opt/nvidia/nsight-compute/*/docs/_static/css/fonts/fontawesome-webfont.ttf"))


@pytest.fixture(scope="module")
def font():
    if not CANDIDATES:
        pytest.skip("no TrueType font in this image")
    from msdfgen_b200.font import Font
    return Font(CANDIDATES[0])


def test_tables_and_character_map(font):
    assert font.n_glyphs == 707 and font.units_per_em == 1792
    assert (font.ascender, font.descender, font.line_height) == (1536, -256, 1792)
    gi = font.glyph_indices(np.arange(0xf000, 0xf2e1))
    assert (gi > 0).sum() == 679 and gi[0] == 13 and len(set(gi[gi > 0])) == 679
    assert font.glyph_indices([0x41, 0x10ffff, 0x20]).tolist()[:2] == [0, 0]


def test_flat_batch_is_well_formed(font):
    shapes, advances, loaded = font.load_glyphs(np.arange(font.n_glyphs), "none")
    assert loaded.all() and shapes.n_glyphs == 707
    assert set(np.unique(shapes.orders)) == {1, 2}                      # TrueType: lines and quadratics
    assert prep.validate(shapes)                                        # every contour closed, edge to edge
    assert (advances > 0).sum() > 600 and advances.max() <= 2 * 1792
    # no zero-length edges, no empty contours; coordinates are integers (implied points use FreeType's integer halving)
    assert (np.diff(shapes.cbegin) > 0).all()
    ends = np.array([shapes.points[i, 2 * o:2 * o + 2] for i, o in enumerate(shapes.orders)])
    assert (np.abs(ends - shapes.points[:, :2]).sum(1) > 0).all()
    assert np.array_equal(shapes.points, np.round(shapes.points))
    # scalings: every point is `scale*vector.x`; the factory's lowering of degenerate quadratics runs on the SCALED points (as in
    # the reference), so a control point that is collinear in integers may stop being so after scaling — orders may differ
    em, _, _ = font.load_glyphs([40, 361], "em")
    raw, _, _ = font.load_glyphs([40, 361], "none")
    assert np.array_equal(em.cbegin, raw.cbegin) and np.array_equal(em.points[:, :2], (1.0 / 1792) * raw.points[:, :2])
    assert (em.orders >= raw.orders).all()
    # bad glyph indices fail softly
    bad, _, ok = font.load_glyphs([5, 100000], "em")
    assert ok.tolist() == [True, False] and bad.n_glyphs == 2


def test_outlines_match_freetype_rendering(font, port):
    """Inside / outside of the ingested outline (the oracle's scanline rasteriser, non-zero rule) against FreeType's anti-aliased
    rendering of the same glyph in Pillow, on pixels FreeType renders clearly black or white.  FreeType hints the outline (stems
    snap to the pixel grid), so a fraction of a percent along edges may differ; a missing contour, a wrong implied point or a
    wrong start rule would flip whole regions."""
    ImageFont = pytest.importorskip("PIL.ImageFont")
    size = 96
    pil = ImageFont.truetype(CANDIDATES[0], size, layout_engine=ImageFont.Layout.BASIC)
    ascent, _ = pil.getmetrics()
    shapes, _, _ = font.load_glyphs(np.arange(font.n_glyphs), "none")
    scale = size / font.units_per_em
    cps = np.arange(0xf000, 0xf2e1)
    compared = different = 0
    for cp, g in list(zip(cps, font.glyph_indices(cps)))[::5]:
        cb, pts, od, cl, _ = shapes.glyph(int(g))
        mask = pil.getmask(chr(int(cp)), mode="L")
        w, h = mask.size
        if g == 0 or len(od) == 0 or w == 0 or h == 0:
            continue
        left, top = pil.getbbox(chr(int(cp)))[:2]
        xf = (scale, scale, -left / scale, -(ascent - top - h) / scale, -1 / scale, 1 / scale)
        inside = port.rasterize(cb, pts, od, xf, w, h, 0).reshape(h, w)[::-1] > 0.5
        coverage = np.asarray(mask, np.uint8).reshape(h, w)
        clear = (coverage > 200) | (coverage < 55)
        different += int(((inside != (coverage > 127)) & clear).sum())
        compared += int(clear.sum())
    assert compared > 500000 and different <= 0.005 * compared, (different, compared)


from test_devcore import devcore  # noqa: E402,F401  (the fixture that builds tests/devcore)


def test_real_glyphs_through_the_texel_path(font, ref, devcore):
    """Normalised, coloured real outlines through k1a's filter + k1b's walk (the big-shape form for glyphs above 96 segments),
    compiled for the host: every float equals the reference's, including glyph 361, whose two abutting contours share three
    quadratic edges traversed in opposite directions (prepare_record's twin rule sends those to the literal solver)."""
    from test_devcore import _filtered, _frame, _texels, _texel_bits, FAST_GUARDED, FOLDED_UNGUARDED
    lib = devcore
    raw, _, _ = font.load_glyphs(np.arange(font.n_glyphs), "em")
    shapes = prep.prepare(raw)
    rng = np.random.default_rng(3)
    values = 0
    for g in sorted(set(range(0, font.n_glyphs, 9)) | {361}):
        cb, pts, od, cl, _ = shapes.glyph(g)
        if len(od) == 0:
            continue
        s = FlatShapes.single(cb, pts, od, cl)
        w = h = 32 if g % 2 else 48
        xf = _frame(s, w, h)
        texels = _texels(xf, w, h, rng)
        for mode, overlap in ((3, True), (4, True), (1, True), (3, False), (2, False)):
            want = ref.distance(mode, overlap, s.cbegin, s.points, s.orders, s.colors, texels)
            got, _ = _filtered(lib, mode, overlap, FOLDED_UNGUARDED if len(od) > 96 else FAST_GUARDED, g % 3 - 1, s, xf, w, h, texels)
            assert np.array_equal(_texel_bits(got, xf), _texel_bits(want, xf)), (g, len(od), mode, overlap)
            values += got.size
    print(f"[font] {values} values of real glyphs identical")


def test_shape_interface_equals_flat_ingest(font):
    """msdfgen's calls (initializeFreetype, loadFont, loadGlyph by code point, getFontMetrics) from a C++ caller give the same
    edges as the batch ingest, and normalize + edgeColoringSimple on the Shape equal prep.prepare on the flat glyph."""
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src, out = os.path.join(root, "tests", "dropin", "font_caller.cpp"), os.path.join(root, "tests", "dropin", "_build")
    host = os.path.join(root, "msdfgen_b200")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libfont_caller.so")
    deps = [src, os.path.join(host, "libmsdfgen_b200_host.so"), os.path.join(host, "host", "import-font.h")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-ffp-contract=off", "-std=c++11", "-fPIC", "-shared", f"-I{host}/host", "-o", lib, src, f"-L{host}",
                        "-lmsdfgen_b200_host", "-lmsdfgen_b200", f"-Wl,-rpath,{host}"], check=True, capture_output=True)
    caller = C.CDLL(lib)
    for cp in (0xf000, 0xf02a, 0xf013, 0xf1ea):
        g = int(font.glyph_indices([cp])[0])
        for scaling, name in ((0, "none"), (1, "em"), (2, "legacy")):
            flat, adv, _ = font.load_glyphs([g], name)
            for prepare in (0, 1):
                want = prep.prepare(flat) if prepare else flat
                cap = 4 * len(flat.orders) + 16
                cb, pts = np.zeros(cap + 1, np.uint32), np.zeros((cap, 8))
                od, cl = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
                advance, metrics = C.c_double(0), np.zeros(6)
                nc = caller.font_caller_glyph(CANDIDATES[0].encode(), C.c_uint(cp), C.c_int(scaling), C.c_int(prepare), C.c_int(cap), C.c_int(cap),
                                              cb.ctypes.data_as(C.POINTER(C.c_uint32)), pts.ctypes.data_as(C.POINTER(C.c_double)),
                                              od.ctypes.data_as(C.POINTER(C.c_uint8)), cl.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(advance),
                                              metrics.ctypes.data_as(C.POINTER(C.c_double)))
                assert nc == want.n_contours, (hex(cp), name, prepare, nc)
                ns = int(cb[nc])
                assert np.array_equal(cb[:nc + 1], want.cbegin) and np.array_equal(od[:ns], want.orders) and np.array_equal(cl[:ns], want.colors)
                assert np.array_equal(pts[:ns].view(np.uint64), want.points.view(np.uint64))
                assert advance.value == adv[0]
                assert np.allclose(metrics, np.array([1792, 1536, -256, 1792, 0, 0]) * font.scale(name), rtol=0, atol=0)


def test_real_glyphs_error_correction_on_the_cpu(font, ref, devcore):
    """The default error correction (EDGE_PRIORITY, CHECK_DISTANCE_AT_EDGE) of real outlines through the device's classification
    code on the host: corrected floats and stencil bytes equal the reference's."""
    from oracle.bindings import Config, EC_DISABLED
    from test_devcore import _frame, _error_correction
    raw, _, _ = font.load_glyphs(np.arange(font.n_glyphs), "em")
    shapes = prep.prepare(raw)
    for g in range(0, font.n_glyphs, 4):
        cb, pts, od, cl, _ = shapes.glyph(g)
        if len(od) == 0:
            continue
        s = FlatShapes.single(cb, pts, od, cl)
        sx, sy, tx, ty = _frame(s, 32, 32)
        xf = (sx, sy, tx, ty, -2.0 / sx, 2.0 / sx)
        for name in ("msdf", "mtsdf") if g % 16 == 0 else ("msdf",):
            field = ref.generate(name, cb, pts, od, cl, xf, 32, 32, Config(ec_mode=EC_DISABLED))
            want, want_stencil = ref.error_correction(field, cb, pts, od, cl, xf, Config(), want_stencil=True)
            got, stencil = _error_correction(devcore, field, s, xf, True, 2, 1)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and np.array_equal(stencil, want_stencil), (g, name)


def test_c_port_equals_compiled_reference_on_real_glyphs(font, ref, port):
    """The GPU leg (tests/test_zz_real_font.py) checks against the C restatement, which is what travels to the GPU box: pin it to
    the compiled reference on the same real outlines, default error correction included."""
    from test_devcore import _frame
    raw, _, _ = font.load_glyphs(np.arange(font.n_glyphs), "em")
    shapes = prep.prepare(raw)
    for g in range(0, font.n_glyphs, 11):
        cb, pts, od, cl, _ = shapes.glyph(g)
        if len(od) == 0:
            continue
        sx, sy, tx, ty = _frame(FlatShapes.single(cb, pts, od, cl), 32, 32)
        xf = (sx, sy, tx, ty, -2.0 / sx, 2.0 / sx)
        for name in ("msdf", "mtsdf"):
            a, b = port.generate(name, cb, pts, od, cl, xf, 32, 32), ref.generate(name, cb, pts, od, cl, xf, 32, 32)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (g, name)


# ---- synthetic fonts: the paths the image's font does not exercise ------------------------------------------------------------------

def _against_freetype(data, chars, port, size=120):
    """(clear pixels compared, pixels that differ) between the ingested outlines' fill and FreeType's rendering of the same bytes"""
    import io
    from PIL import ImageFont
    from msdfgen_b200.font import Font
    f = Font(data)
    pil = ImageFont.truetype(io.BytesIO(data), size, layout_engine=ImageFont.Layout.BASIC)
    ascent, _ = pil.getmetrics()
    scale = size / f.units_per_em
    compared = different = 0
    for ch in chars:
        g = int(f.glyph_indices([ord(ch)])[0])
        assert g != 0
        shapes, _, ok = f.load_glyphs([g], "none")
        assert ok[0] and len(shapes.orders)
        mask = pil.getmask(ch, mode="L")
        w, h = mask.size
        left, top = pil.getbbox(ch)[:2]
        xf = (scale, scale, -left / scale, -(ascent - top - h) / scale, -1 / scale, 1 / scale)
        inside = port.rasterize(shapes.cbegin, shapes.points, shapes.orders, xf, w, h, 0).reshape(h, w)[::-1] > 0.5
        coverage = np.asarray(mask, np.uint8).reshape(h, w)
        clear = (coverage > 215) | (coverage < 40)
        different += int(((inside != (coverage > 127)) & clear).sum())
        compared += int(clear.sum())
    return compared, different


def test_contours_that_start_on_control_points(port):
    """FT_Outline_Decompose's start rules: first point off the curve and last point on it (start there), both off (start halfway,
    integer halving), control points only (every on-curve point implied); plus an ordinary contour as the control."""
    pytest.importorskip("PIL.ImageFont")
    from ttf_builder import build, simple, ON, OFF
    blob = [(500, 60, OFF), (900, 120, ON), (960, 500, OFF), (880, 900, OFF), (470, 950, ON), (90, 880, OFF), (55, 430, OFF), (130, 101, ON)]
    glyphs = [simple([]),
              simple([blob]),                                                                   # first off, last on
              simple([blob[:-1]]),                                                              # first off, last off
              simple([[(x, y, OFF) for x, y, _ in blob]]),                                      # control points only
              simple([blob[1:] + blob[:1]]),                                                    # first on (control)
              simple([[(501, 61, OFF), (901, 121, OFF), (961, 501, OFF), (101, 901, OFF), (33, 333, OFF)], blob[1:5]])]   # odd sums: halving truncates
    data = build(glyphs, {ord("A") + i: i + 1 for i in range(5)})
    compared, different = _against_freetype(data, "ABCDE", port)
    assert compared > 30000 and different <= 0.004 * compared, (different, compared)
    # the implied points themselves: integer halving towards zero, as FT_Outline_Decompose does it
    from msdfgen_b200.font import Font
    shapes, _, _ = Font(data).load_glyphs([5], "none")
    starts = {tuple(p) for p in shapes.points[:, :2]}
    assert (701.0, 91.0) in starts and (931.0, 311.0) in starts and (267.0, 197.0) in starts        # (501+901)/2, (61+121)/2, ... (33+501)/2, (333+61)/2
    # a left side bearing that disagrees with the outline's xMin moves the outline (FreeType puts the origin where `hmtx` says)
    moved = build(glyphs, {ord("A") + i: i + 1 for i in range(5)}, lsb=[0, 0, 100, 55, 130, 33])
    compared, different = _against_freetype(moved, "ABCDE", port)
    assert different <= 0.004 * compared, (different, compared)
    a, _, _ = Font(moved).load_glyphs([1, 2], "none")
    b, _, _ = Font(data).load_glyphs([1, 2], "none")
    assert np.array_equal(a.points[:, 0::2][b.points[:, 0::2] != 0], (b.points[:, 0::2] - np.repeat([[55], [-45]], [5, len(a.orders) - 5], 0))[b.points[:, 0::2] != 0])


def test_composite_glyphs(port):
    """Components with plain offsets, a uniform scale, separate x / y scales, a 2x2 matrix (rotation + shear), nesting, and point
    matching — assembled in 16.16 arithmetic like FreeType's loader, rendered by FreeType for comparison."""
    pytest.importorskip("PIL.ImageFont")
    from ttf_builder import build, simple, composite, ON, OFF
    bar = simple([[(0, 0, ON), (0, 600, ON), (150, 600, ON), (150, 0, ON)]])
    drop = simple([[(300, 50, ON), (100, 300, OFF), (300, 700, ON), (500, 300, OFF)]])
    glyphs = [simple([]), bar, drop,
              composite([{"glyph": 1, "dx": 100, "dy": 100}, {"glyph": 1, "dx": 500, "dy": 250}, {"glyph": 2, "dx": 250, "dy": 120}]),
              composite([{"glyph": 2, "dx": 40, "dy": 30, "scale": 0.61803}, {"glyph": 2, "dx": 420, "dy": 350, "scale": 0.75}]),
              composite([{"glyph": 2, "dx": 10, "dy": 10, "xscale": 1.5, "yscale": 0.5}, {"glyph": 1, "dx": 700, "dy": 300, "xscale": 0.4, "yscale": 1.1}]),
              composite([{"glyph": 2, "dx": 500, "dy": 100, "matrix": (0.8, 0.5, -0.45, 0.85)}, {"glyph": 1, "dx": 50, "dy": 50, "matrix": (1.0, 0.0, 0.37, 1.0)}]),
              composite([{"glyph": 4, "dx": 0, "dy": 0}, {"glyph": 3, "dx": 60, "dy": 40, "scale": 0.5}]),          # nested composites
              composite([{"glyph": 1, "dx": 200, "dy": 100}, {"glyph": 2, "match": (2, 0)}])]                       # drop's point 0 on bar's point 2
    data = build(glyphs, {ord("A") + i: i + 3 for i in range(6)})
    compared, different = _against_freetype(data, "ABCDEF", port)
    assert compared > 40000 and different <= 0.004 * compared, (different, compared)
    from msdfgen_b200.font import Font
    f = Font(data)
    plain, _, _ = f.load_glyphs([3], "none")
    assert plain.n_contours == 3 and len(plain.orders) == 4 + 4 + 2
    scaled, _, _ = f.load_glyphs([4], "none")
    # mulFix(300, round(0.61803 * 16384) * 4) + 40: the 2.14 scale widened to 16.16, products rounded to nearest
    s = round(0.61803 * 16384) * 4
    assert scaled.points[0, 0] == ((300 * s + 0x8000) >> 16) + 40 and scaled.points[0, 1] == ((50 * s + 0x8000) >> 16) + 30
    matched, _, _ = f.load_glyphs([8], "none")
    assert {(350.0, 700.0)} <= {tuple(p) for p in matched.points[:, :2]}                  # bar's point 2 (150, 600) + (200, 100)


def test_character_maps_and_kerning():
    from ttf_builder import build, simple, ON
    from msdfgen_b200.font import Font
    box = simple([[(0, 0, ON), (0, 500, ON), (500, 500, ON), (500, 0, ON)]])
    cmap = {0x41: 1, 0x42: 2, 0x3b1: 3, 0x1f600: 4, 0x20: 5}
    kern = {(1, 2): -80, (2, 1): 35, (3, 4): -7}
    for cmap12 in (False, True):
        f = Font(build([simple([]), box, box, box, box, simple([])], cmap, advances=[500, 600, 610, 620, 630, 250], kern=kern, cmap12=cmap12))
        want = [1, 2, 3, 4 if cmap12 else 0, 5, 0]
        assert f.glyph_indices([0x41, 0x42, 0x3b1, 0x1f600, 0x20, 0x43]).tolist() == want
        assert f.has_kerning and f.kerning(1, 2) == -80 and f.kerning(2, 1) == 35 and f.kerning(3, 4) == -7 and f.kerning(1, 3) == 0
        _, adv, _ = f.load_glyphs([1, 4, 5], "em")
        assert adv.tolist() == [0.6, 0.63, 0.25]
        assert (f.ascender, f.descender, f.line_height, f.underline_position, f.underline_thickness) == (800, -200, 1000, -125, 50)
