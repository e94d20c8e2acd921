"""CPU tests of the test oracle itself: the plain-C restatement (oracle/msdf_oracle.c) against the committed golden
vectors and — where the compiled reference is available — live against oracle/_ref, bit for bit."""
import numpy as np
import pytest

from golden_io import load_cases
from oracle.bindings import Config
from util import compare, to_bytes

from msdfgen_b200.synth import synth_glyph, glyph_transform

CASES = load_cases()


def _cfg(c):
    return Config(overlap=c["overlap"], ec_mode=c["ec_mode"], ec_dcm=c["ec_dcm"], min_deviation_ratio=c["min_dev"], min_improve_ratio=c["min_imp"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_port_matches_golden(port, case):
    out, st = port.generate(case["mode"], case["cbegin"], case["points"], case["orders"], case["colors"], case["xf"], case["w"], case["h"],
                            _cfg(case), shape_ydown=case["shape_ydown"], out_ydown=case["out_ydown"], want_stencil=True)
    diff_bits, max_abs, over = compare(out, case["out"])
    assert diff_bits == 0, f"{diff_bits} differing floats, max {max_abs}"
    if case["mode"] in ("msdf", "mtsdf") and case["ec_mode"] != 0:
        assert np.array_equal(st, case["stencil"])


def test_golden_is_current_reference(ref):
    """The committed fixture still equals what the compiled reference produces (guards against a stale fixture)."""
    for case in CASES[::5]:
        out = ref.generate(case["mode"], case["cbegin"], case["points"], case["orders"], case["colors"], case["xf"], case["w"], case["h"],
                           _cfg(case), shape_ydown=case["shape_ydown"], out_ydown=case["out_ydown"])
        assert compare(out, case["out"])[0] == 0


def test_port_vs_reference_random_shapes(port, ref):
    for i in range(12):
        g = synth_glyph(7000 + i, cubic_fraction=0.3 if i % 2 else 0.0)
        cb, pts, od, cl = ref.prepare(g.cbegin, g.points, g.orders)
        for mode in ("sdf", "psdf", "msdf", "mtsdf"):
            cfg = Config(overlap=bool(i % 3))
            a, sa = ref.generate(mode, cb, pts, od, cl, glyph_transform(20), 20, 20, cfg, want_stencil=True)
            b, sb = port.generate(mode, cb, pts, od, cl, glyph_transform(20), 20, 20, cfg, want_stencil=True)
            assert compare(a, b)[0] == 0
            assert np.array_equal(sa, sb)


def test_port_pieces_vs_reference(port, ref):
    rng = np.random.default_rng(5)
    g = synth_glyph(4242, cubic_fraction=0.3)
    cb, pts, od, cl = ref.prepare(g.cbegin, g.points, g.orders)
    q = rng.uniform(-0.3, 1.3, (200, 2))
    a = ref.segment_distances(pts, od, q)
    b = port.segment_distances(pts, od, q)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    for mode in ("sdf", "psdf", "msdf", "mtsdf"):
        for ov in (True, False):
            da = ref.distance(mode, ov, cb, pts, od, cl, q)
            db = port.distance(mode, ov, cb, pts, od, cl, q)
            assert np.array_equal(da.view(np.uint64), db.view(np.uint64))
    assert np.array_equal(ref.windings(cb, pts, od), port.windings(cb, pts, od))
    for _ in range(300):
        co = rng.normal(size=4) * 10.0 ** rng.integers(-6, 7, 4)
        if rng.uniform() < 0.1:
            co[0] = 0
        assert ref.solve_cubic(*co) == port.solve_cubic(*co)
        assert ref.solve_quadratic(*co[:3]) == port.solve_quadratic(*co[:3])


def test_error_correction_entry_points(port, ref):
    g = synth_glyph(99)
    cb, pts, od, cl = ref.prepare(g.cbegin, g.points, g.orders)
    xf = glyph_transform(24)
    raw = ref.generate("msdf", cb, pts, od, cl, xf, 24, 24, Config(ec_mode=0))
    for ecm in (1, 2, 3):
        for dcm in (0, 1, 2):
            cfg = Config(ec_mode=ecm, ec_dcm=dcm)
            a, sa = ref.error_correction(raw, cb, pts, od, cl, xf, cfg, want_stencil=True)
            b, sb = port.error_correction(raw, cb, pts, od, cl, xf, cfg, want_stencil=True)
            assert compare(a, b)[0] == 0 and np.array_equal(sa, sb)
    for edge in (False, True):
        assert compare(ref.fast_error_correction(raw, xf, edge), port.fast_error_correction(raw, xf, edge))[0] == 0


def test_pixel_float_to_byte(port):
    x = np.concatenate([np.linspace(-0.5, 1.5, 4001), [np.inf, -np.inf, 0.0, 1.0, 0.5]]).astype(np.float32)
    assert np.array_equal(port.float_to_byte(x), to_bytes(x))


def test_empty_shape_is_minus_infinity(port):
    out = port.generate("msdf", np.zeros(1, np.uint32), np.zeros((0, 8)), np.zeros(0, np.uint8), np.zeros(0, np.uint8), glyph_transform(4), 4, 4)
    assert np.all(np.isneginf(out))
