"""GPU parity tests of the scanline pass (run on the B200 box with -m gpu): b200msdf_sign_correction / b200msdf_rasterize /
config.scanline_pass through the C ABI against the golden vectors of the compiled reference (distanceSignCorrection and
rasterize, core/rasterization.cpp) and the C restatement.  The pass only flips floats (x -> 2*zero - x), so the bar here is
bit equality."""
import numpy as np
import pytest

from golden_io import load_scanline_cases
from oracle.bindings import Config
from util import to_bytes

from msdfgen_b200 import api, prep
from msdfgen_b200.flat import FlatShapes
from msdfgen_b200.synth import synth_atlas, synth_glyph, synth_big_shape, glyph_transform

pytestmark = pytest.mark.gpu

CASES = load_scanline_cases()
MODE = {"sdf": api.SDF, "psdf": api.PSDF, "msdf": api.MSDF, "mtsdf": api.MTSDF}
EC_OFF = api.ErrorCorrectionConfig(mode=api.EC_DISABLED)


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def ndiff(a, b):
    return int((np.ascontiguousarray(a).view(np.uint32) != np.ascontiguousarray(b).view(np.uint32)).sum())


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_scanline_golden_vectors(ctx, case):
    shape = FlatShapes.single(case["cbegin"], case["points"], case["orders"], case["colors"], ydown=case["shape_ydown"])
    xf = case["xf"][:4]
    if case["kind"] == "raster":
        out = np.full((case["h"], case["w"], 1), np.nan, np.float32)
        ctx.rasterize([out], shape, xf, case["rule"], out_ydown=[case["sdf_ydown"]])
    else:
        out = case["field"].copy()
        ctx.sign_correction([out], shape, xf, case["zero"], case["rule"], sdf_ydown=[case["sdf_ydown"]])
    assert same(out, case["out"]), f"{case['name']}: {ndiff(out, case['out'])} differing floats"


def test_sign_correction_batch_vs_port(ctx, port):
    """A mixed atlas batch (glyphs of different sizes, half of the fields inverted) for 1, 3 and 4 channels and every rule."""
    n = 96
    atlas = prep.prepare(synth_atlas(n, seed=5150, cubic_fraction=0.2))
    rng = np.random.default_rng(3)
    sizes = [(int(rng.integers(9, 40)), int(rng.integers(9, 40))) for _ in range(n)]        # (w, h)
    xfs = np.stack([glyph_transform(min(w, h)) for w, h in sizes])
    for mode in (api.SDF, api.MSDF, api.MTSDF):
        ch = api.CHANNELS[mode]
        fields = [np.empty((h, w, ch), np.float32) for w, h in sizes]
        ctx.generate_batch(mode, fields, atlas, xfs, api.MSDFGeneratorConfig(errorCorrection=EC_OFF))
        for g in range(0, n, 2):
            fields[g][...] = 1 - fields[g]
        for rule in range(4):
            got = [f.copy() for f in fields]
            ctx.sign_correction(got, atlas, xfs[:, :4], 0.5, rule)
            for g in range(n):
                cb, pts, od, cl, _ = atlas.glyph(g)
                want = port.sign_correction(fields[g], cb, pts, od, xfs[g], 0.5, rule)
                assert same(got[g], want), f"mode {mode} rule {rule} glyph {g}: {ndiff(got[g], want)} differing floats"


def test_scanline_pipeline_equals_reference_composition(ctx, port):
    """config.scanline_pass: generate (error correction postponed) -> distanceSignCorrection -> msdfErrorCorrection in ONE device
    batch equals the three reference calls the command line makes (main.cpp:1223-1297)."""
    n = 48
    atlas = prep.prepare(synth_atlas(n, seed=6161, cubic_fraction=0.15))
    xf = glyph_transform(32)
    for mode_name in ("msdf", "mtsdf", "sdf"):
        mode = MODE[mode_name]
        ch = api.CHANNELS[mode]
        for rule in (api.FILL_NONZERO, api.FILL_ODD):
            cfg = api.MSDFGeneratorConfig(scanlinePass=True, fillRule=rule)
            got = np.full((n, 32, 32, ch), np.nan, np.float32)
            ctx.generate_batch(mode, got, atlas, xf, cfg)
            pinned = ctx.pinned((n, 32, 32, ch), np.float32)
            ctx.generate_batch(mode, pinned, atlas, xf, cfg)
            assert same(got, np.asarray(pinned))
            off = api.MSDFGeneratorConfig(scanlinePass=True, fillRule=rule, errorCorrection=EC_OFF)
            got_off = np.full((n, 32, 32, ch), np.nan, np.float32)
            ctx.generate_batch(mode, got_off, atlas, xf, off)
            ctx.generate_batch(mode, pinned, atlas, xf, off)        # zero-copy destination: the scanline kernel makes the final store
            assert same(got_off, np.asarray(pinned))
            for g in range(n):
                cb, pts, od, cl, _ = atlas.glyph(g)
                want = port.generate(mode_name, cb, pts, od, cl, xf, 32, 32, Config(ec_mode=0))
                want = port.sign_correction(want, cb, pts, od, xf, 0.5, rule)
                assert same(got_off[g], want), f"{mode_name} rule {rule} glyph {g} (ec off): {ndiff(got_off[g], want)}"
                if ch >= 3:
                    want = port.error_correction(want, cb, pts, od, cl, xf, Config())
                assert same(got[g], want), f"{mode_name} rule {rule} glyph {g}: {ndiff(got[g], want)}"
    # 8-bit output goes through the same pass
    cfg = api.MSDFGeneratorConfig(scanlinePass=True, fillRule=api.FILL_ODD)
    f = np.empty((n, 32, 32, 3), np.float32)
    ctx.generate_batch(api.MSDF, f, atlas, xf, cfg)
    b = np.empty((n, 32, 32, 3), np.uint8)
    ctx.generate_batch_u8(api.MSDF, list(b), atlas, xf, cfg)
    assert np.array_equal(b, to_bytes(f))


def test_resident_batch_sign_correction(ctx):
    atlas = prep.prepare(synth_atlas(40, seed=42))
    xf = glyph_transform(24)
    off = api.MSDFGeneratorConfig(errorCorrection=EC_OFF)
    for mode in (api.PSDF, api.MTSDF):
        ch = api.CHANNELS[mode]
        host = np.empty((40, 24, 24, ch), np.float32)
        ctx.generate_batch(mode, host, atlas, xf, off)
        ctx.sign_correction(list(host), atlas, xf, 0.5, api.FILL_NEGATIVE)
        b = ctx.upload(mode, atlas, xf, [24, 24])
        b.launch_stages(3, off)
        b.sign_correction(0.5, api.FILL_NEGATIVE)
        flat = b.download()
        for g in range(40):
            assert same(b.glyph(flat, g), host[g])
        b.free()


def test_big_shapes_and_ragged_tiles(ctx, port):
    """Shapes beyond the shared-memory record budget, rows that straddle tiles (w not dividing 1024, w > 1024, w = 1)."""
    big = prep.prepare(synth_big_shape(n_contours=12, segs_per_contour=30, seed=5))
    cb, pts, od, cl, _ = big.glyph(0)
    for (w, h), mode_name in (((77, 53), "mtsdf"), ((2500, 3), "sdf"), ((1, 1500), "sdf"), ((1024, 2), "msdf"), ((1025, 2), "msdf")):
        xf = glyph_transform(w, h, px_range=6.0)
        raw = np.empty((h, w, api.CHANNELS[MODE[mode_name]]), np.float32)
        ctx.generate(MODE[mode_name], raw, big, xf, api.MSDFGeneratorConfig(errorCorrection=EC_OFF))
        raw = (1 - raw).astype(np.float32)
        for rule in (0, 1):
            got = raw.copy()
            ctx.sign_correction([got], big, xf, 0.5, rule)
            want = port.sign_correction(raw, cb, pts, od, xf, 0.5, rule)
            assert same(got, want), f"{w}x{h} {mode_name} rule {rule}: {ndiff(got, want)} differing floats"
            assert ndiff(got, raw) > 0
        ras = np.empty((h, w, 1), np.float32)
        ctx.rasterize([ras], big, xf)
        assert same(ras, port.rasterize(cb, pts, od, xf, w, h))


def test_regions_sign_correction(ctx):
    """Row bands of one image as a region batch: single-channel fields can be sign-corrected per band; multi-channel ones need
    their neighbours and are refused."""
    shape = prep.prepare(synth_big_shape(20, 20))
    size = 96
    xf = glyph_transform(size, px_range=8.0)
    off = api.MSDFGeneratorConfig(errorCorrection=EC_OFF, scanlinePass=True, fillRule=api.FILL_ODD)
    full = np.empty((size, size, 1), np.float32)
    ctx.generate(api.SDF, full, shape, xf, off)
    for r0, r1 in ((0, 40), (40, 96)):
        b = ctx.upload(api.SDF, shape, xf, [size, r1 - r0], origins=[0, r0])
        b.launch(off)
        part = b.download()[: (r1 - r0) * size].reshape(r1 - r0, size, 1)
        assert same(part, full[r0:r1])
        b.free()
    b = ctx.upload(api.MSDF, shape, xf, [size, 40], origins=[0, 0])
    with pytest.raises(Exception):
        b.launch(off)
    b.free()


def test_scanline_argument_checks(ctx):
    g = prep.prepare(synth_glyph(3))
    f = np.zeros((8, 8, 2), np.float32)
    with pytest.raises(Exception):
        ctx.sign_correction([f], g, glyph_transform(8))            # 2 channels
    with pytest.raises(Exception):
        ctx.sign_correction([np.zeros((8, 8, 3), np.float32)], g, glyph_transform(8), 0.5, 7)     # fill rule
    with pytest.raises(Exception):
        ctx.generate(api.MSDF, np.zeros((8, 8, 3), np.float32), g, glyph_transform(8), api.MSDFGeneratorConfig(scanlinePass=True, fillRule=9))
    # zero-sized fields and empty batches are no-ops
    ctx.sign_correction([np.zeros((0, 5, 3), np.float32)], g, glyph_transform(8))
