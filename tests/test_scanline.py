"""CPU tests of the scanline pass (distanceSignCorrection / rasterize, core/rasterization.cpp): the plain-C restatement
against the committed golden vectors of the compiled reference and, where oracle/_ref is available, live against it."""
import numpy as np
import pytest

from golden_io import load_scanline_cases

from msdfgen_b200.synth import synth_glyph, glyph_transform
from oracle.bindings import Config

CASES = load_scanline_cases()


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_port_matches_scanline_golden(port, case):
    if case["kind"] == "raster":
        out = port.rasterize(case["cbegin"], case["points"], case["orders"], case["xf"], case["w"], case["h"], case["rule"],
                             shape_ydown=case["shape_ydown"], out_ydown=case["sdf_ydown"])
    else:
        out = port.sign_correction(case["field"], case["cbegin"], case["points"], case["orders"], case["xf"], case["zero"], case["rule"],
                                   shape_ydown=case["shape_ydown"], sdf_ydown=case["sdf_ydown"])
    assert _same(out, case["out"]), f"{(out != case['out']).sum()} differing floats"


def test_scanline_golden_is_current_reference(ref):
    for case in CASES[::4]:
        if case["kind"] == "raster":
            out = ref.rasterize(case["cbegin"], case["points"], case["orders"], case["xf"], case["w"], case["h"], case["rule"],
                                shape_ydown=case["shape_ydown"], out_ydown=case["sdf_ydown"])
        else:
            out = ref.sign_correction(case["field"], case["cbegin"], case["points"], case["orders"], case["xf"], case["zero"], case["rule"],
                                      shape_ydown=case["shape_ydown"], sdf_ydown=case["sdf_ydown"])
        assert _same(out, case["out"])


def test_port_vs_reference_random_shapes(port, ref):
    for i in range(16):
        g = synth_glyph(9100 + i, cubic_fraction=0.4 if i % 2 else 0.0)
        cb, pts, od, cl = ref.prepare(g.cbegin, g.points, g.orders, normalize=bool(i % 3))
        xf = glyph_transform(20)
        for mode in ("sdf", "msdf", "mtsdf"):
            f = ref.generate(mode, cb, pts, od, cl, xf, 20, 20, Config(ec_mode=0))
            if i % 4 == 1:
                f = (1 - f).astype(np.float32)
            for rule in range(4):
                assert _same(ref.sign_correction(f, cb, pts, od, xf, 0.5, rule), port.sign_correction(f, cb, pts, od, xf, 0.5, rule))
        for rule in range(4):
            assert _same(ref.rasterize(cb, pts, od, xf, 20, 20, rule), port.rasterize(cb, pts, od, xf, 20, 20, rule))


def test_sign_correction_properties(port):
    """Idempotent; a correctly signed field is a fixed point; an inverted field is restored (NONZERO on normalised glyphs)."""
    for i in range(4):
        g = synth_glyph(9300 + i)
        from msdfgen_b200 import prep
        cb, pts, od, cl, _ = prep.prepare(g).glyph(0)
        xf = glyph_transform(24)
        f = port.generate("mtsdf", cb, pts, od, cl, xf, 24, 24, Config(ec_mode=0))
        once = port.sign_correction(f, cb, pts, od, xf)
        assert _same(port.sign_correction(once, cb, pts, od, xf), once)
        inv = (2 * np.float32(0.5) - once).astype(np.float32)
        assert np.allclose(port.sign_correction(inv, cb, pts, od, xf), once, atol=1e-6, rtol=0)
        fill = port.rasterize(cb, pts, od, xf, 24, 24)[..., 0] > 0
        med = np.median(once[..., :3], axis=2)
        assert np.array_equal(med[med != 0.5] > 0.5, fill[med != 0.5])
        assert np.array_equal(once[..., 3][once[..., 3] != 0.5] > 0.5, fill[once[..., 3] != 0.5])
