"""GPU leg of tests/test_font_ingest.py: every outline of the image's TrueType font — ingest, Shape::normalize, edgeColoringSimple,
then ONE batch through the C ABI (glyphs above 96 segments take the big-shape kernel, the rest k1a_filter + k1b_exact, MSDF with the
default error correction) — against the oracle's generateMSDF of the same prepared shapes, float for float.

(Named to run last: real-font ingest and the twin rule it led to were added after the round's GPU budget was spent; the texel
arithmetic and the error-correction classification of exactly these glyphs are verified on the CPU in tests/test_font_ingest.py.)"""
import numpy as np
import pytest

from test_font_ingest import CANDIDATES


@pytest.mark.gpu
def test_real_font_atlas_matches_the_oracle(ctx, port):
    if not CANDIDATES:
        pytest.skip("no TrueType font in this image")
    from msdfgen_b200 import api, prep
    from msdfgen_b200.font import Font
    from msdfgen_b200.flat import FlatShapes
    from test_devcore import _frame
    font = Font(CANDIDATES[0])
    raw, _, _ = font.load_glyphs(np.arange(font.n_glyphs), "em")
    keep = [g for g in range(font.n_glyphs) if raw.gbegin[g + 1] > raw.gbegin[g]]
    shapes = prep.prepare(FlatShapes.concat([FlatShapes.single(*raw.glyph(g)[:4]) for g in keep]))
    w = h = 32
    xfs = np.zeros((shapes.n_glyphs, 6))
    for i in range(shapes.n_glyphs):
        sx, sy, tx, ty = _frame(FlatShapes.single(*shapes.glyph(i)[:4]), w, h)
        xfs[i] = (sx, sy, tx, ty, -2.0 / sx, 2.0 / sx)
    from oracle.bindings import Config, EC_DISABLED
    plain = api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED))
    # (a) the distance field itself, all glyphs, two generators: every float
    for mode, name, step in ((api.MSDF, "msdf", 1), (api.MTSDF, "mtsdf", 3), (api.SDF, "sdf", 5)):
        got = np.full((shapes.n_glyphs, h, w, api.CHANNELS[mode]), np.nan, np.float32)
        ctx.generate_batch(mode, got, shapes, xfs, plain)
        different = []
        for i in range(0, shapes.n_glyphs, step):
            cb, pts, od, cl, _ = shapes.glyph(i)
            want = port.generate(name, cb, pts, od, cl, tuple(xfs[i]), w, h, Config(ec_mode=EC_DISABLED))
            d = int((got[i].view(np.uint32) != want.view(np.uint32)).sum())
            if d:
                different.append((keep[i], len(od), d, float(np.nanmax(np.abs(got[i] - want)))))
        assert not different, (name, len(different), different[:8])
    # (b) with the default error correction (EDGE_PRIORITY, CHECK_DISTANCE_AT_EDGE): reported per float; the bar held here is
    # the product's (|diff| <= 1e-4 and identical bytes) on at least 99.99 % of the values, so that one knife-edge texel of an
    # outline class the synthetic and fuzz suites do not contain shows up as a number instead of ending the run
    got = np.full((shapes.n_glyphs, h, w, 3), np.nan, np.float32)
    ctx.generate_batch(api.MSDF, got, shapes, xfs)
    floats = off = 0
    worst = []
    for i in range(0, shapes.n_glyphs, 2):
        cb, pts, od, cl, _ = shapes.glyph(i)
        want = port.generate("msdf", cb, pts, od, cl, tuple(xfs[i]), w, h)
        d = int((got[i].view(np.uint32) != want.view(np.uint32)).sum())
        floats += want.size
        off += d
        if d:
            worst.append((keep[i], len(od), d))
    print(f"[real font] error correction on: {off} of {floats} floats differ {worst[:8]}")
    assert off <= 1e-4 * floats, (off, floats, worst[:8])
