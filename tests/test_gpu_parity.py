"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, against the committed
golden vectors of the compiled reference, the C restatement, and — at BASELINE.json's full sizes — size-independent
properties.  Bar (north_star): |delta| <= 1e-4 on every float and identical bytes after pixelFloatToByte; in practice the
fp64 device path is bit-identical, so the tests also count differing float bit patterns and keep that count at zero
wherever no transcendental is involved (it is reported, not asserted, elsewhere)."""
import numpy as np
import pytest

from golden_io import load_cases
from oracle.bindings import Config
from util import compare, to_bytes

from msdfgen_b200 import api, prep
from msdfgen_b200.flat import FlatShapes
from msdfgen_b200.synth import synth_atlas, synth_glyph, synth_big_shape, glyph_transform

pytestmark = pytest.mark.gpu

CASES = load_cases()
MODE = {"sdf": api.SDF, "psdf": api.PSDF, "msdf": api.MSDF, "mtsdf": api.MTSDF}
TOL = 1e-4


def _cfg(c):
    return api.MSDFGeneratorConfig(overlapSupport=c["overlap"], errorCorrection=api.ErrorCorrectionConfig(
        mode=c["ec_mode"], distanceCheckMode=c["ec_dcm"], minDeviationRatio=c["min_dev"], minImproveRatio=c["min_imp"]))


def _ocfg(cfg: api.MSDFGeneratorConfig):
    e = cfg.errorCorrection
    return Config(overlap=cfg.overlapSupport, ec_mode=e.mode, ec_dcm=e.distanceCheckMode, min_deviation_ratio=e.minDeviationRatio,
                  min_improve_ratio=e.minImproveRatio)


def assert_parity(got, want, what=""):
    bits, max_abs, over = compare(got, want)
    assert over == 0 and max_abs <= TOL, f"{what}: {over} texels over {TOL}, max |delta| {max_abs}, {bits} differing floats"
    assert np.array_equal(to_bytes(got), to_bytes(want)), f"{what}: 8-bit quantisation differs"
    return bits


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_vectors(ctx, case):
    """generate{SDF,PSDF,MSDF,MTSDF} through b200msdf_generate vs the reference's own output, incl. the EC stencil."""
    mode = MODE[case["mode"]]
    shape = FlatShapes.single(case["cbegin"], case["points"], case["orders"], case["colors"], ydown=case["shape_ydown"])
    out = np.full((case["h"], case["w"], api.CHANNELS[mode]), np.nan, np.float32)
    st = np.full((case["h"], case["w"]), 255, np.uint8)
    ec = mode >= api.MSDF and case["ec_mode"] != 0
    ctx.generate(mode, out, shape, case["xf"], _cfg(case), out_ydown=case["out_ydown"], ec_buffer=st if ec else None)
    assert_parity(out, case["out"], case["name"])
    if ec:
        assert np.array_equal(st, case["stencil"]), f"{case['name']}: stencil differs in {(st != case['stencil']).sum()} texels"


def test_batch_equals_single_calls_and_port(ctx, port):
    """A mixed batch (different sizes, orientations, per-glyph transforms, strided / flipped outputs) through
    b200msdf_generate_batch equals per-glyph calls and the oracle."""
    atlas = prep.prepare(synth_atlas(24, seed=4100, cubic_fraction=0.15))
    rng = np.random.default_rng(1)
    sizes = [(int(rng.integers(9, 40)), int(rng.integers(9, 40))) for _ in range(atlas.n_glyphs)]
    xfs = np.array([[w * rng.uniform(0.8, 1.1), h * rng.uniform(0.8, 1.1), rng.uniform(-0.05, 0.1), rng.uniform(-0.05, 0.1), -2.0 / w, 2.5 / w] for w, h in sizes])
    ydown = rng.integers(0, 2, atlas.n_glyphs)
    atlas.ydown[:] = rng.integers(0, 2, atlas.n_glyphs)
    for mode_name, mode in MODE.items():
        n = api.CHANNELS[mode]
        # outputs live inside one big "atlas page" with padding => non-packed row strides
        page = np.full((atlas.n_glyphs, 48, 48 * n + 5), np.nan, np.float32)
        outs = [page[g, 3:3 + h, 2:2 + w * n].reshape(h, w, n) for g, (w, h) in enumerate(sizes)]
        assert all(o.base is not None for o in outs)
        ctx.generate_batch(mode, outs, atlas, xfs, out_ydown=ydown)
        for g, (w, h) in enumerate(sizes):
            cb, pts, od, cl, yd = atlas.glyph(g)
            want = port.generate(mode_name, cb, pts, od, cl, xfs[g], w, h, shape_ydown=yd, out_ydown=bool(ydown[g]))
            assert_parity(outs[g], want, f"{mode_name} glyph {g}")
            single = np.empty((h, w, n), np.float32)
            ctx.generate(mode, single, FlatShapes.single(cb, pts, od, cl, ydown=yd), xfs[g], out_ydown=bool(ydown[g]))
            assert np.array_equal(single.view(np.uint32), outs[g].view(np.uint32))
        assert np.isnan(page[:, :3]).all() and np.isnan(page[:, :, :2]).all()      # nothing written outside the sections


def test_negative_row_stride(ctx, port):
    g = prep.prepare(synth_glyph(77))
    w = 20
    buf = np.full((w, w, 3), np.nan, np.float32)
    flipped = buf[::-1]          # BitmapSection with negative rowStride
    ctx.generate(api.MSDF, flipped, g, glyph_transform(w))
    cb, pts, od, cl, _ = g.glyph(0)
    want = port.generate("msdf", cb, pts, od, cl, glyph_transform(w), w, w)
    assert_parity(flipped, want, "negative stride")


def test_error_correction_entry_points(ctx, port):
    g = prep.prepare(synth_glyph(99, cubic_fraction=0.2))
    cb, pts, od, cl, _ = g.glyph(0)
    xf = glyph_transform(28)
    raw = np.empty((28, 28, 3), np.float32)
    ctx.generate(api.MSDF, raw, g, xf, api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED)))
    for n in (3, 4):
        field = raw if n == 3 else np.concatenate([raw, raw[..., :1]], axis=2).copy()
        for ecm in (1, 2, 3):
            for dcm in (0, 1, 2):
                cfg = api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=ecm, distanceCheckMode=dcm))
                want, wst = port.error_correction(field, cb, pts, od, cl, xf, _ocfg(cfg), want_stencil=True)
                got = field.copy()
                st = np.zeros((28, 28), np.uint8)
                ctx.error_correction([got], g, xf, cfg, ec_buffer=st)
                assert_parity(got, want, f"ec n={n} mode={ecm} dcm={dcm}")
                assert np.array_equal(st, wst)
        for edge in (False, True):
            want = port.fast_error_correction(field, xf, edge, 1.2)
            got = ctx.fast_error_correction(field.copy(), xf, edge, 1.2)
            assert_parity(got, want, f"fast ec edge={edge}")


def test_degenerate_inputs(ctx):
    empty = FlatShapes.single(np.zeros(1, np.uint32), np.zeros((0, 8)), np.zeros(0, np.uint8))
    for mode in (api.SDF, api.MSDF, api.MTSDF):
        out = np.zeros((6, 5, api.CHANNELS[mode]), np.float32)
        ctx.generate(mode, out, empty, glyph_transform(6))
        assert np.all(np.isneginf(out)), "empty shape must give -inf texels like the reference"
    # empty batch and zero-sized bitmaps are no-ops
    none = FlatShapes(np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros((0, 8)), np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    ctx.generate_batch(api.MSDF, np.zeros((0, 4, 4, 3), np.float32), none, glyph_transform(4))
    ctx.generate(api.MSDF, np.zeros((0, 7, 3), np.float32), prep.prepare(synth_glyph(1)), glyph_transform(7))
    # invalid arguments are reported, not computed
    with pytest.raises(Exception):
        ctx.generate(9, np.zeros((4, 4, 3), np.float32), empty, glyph_transform(4))


def test_big_shape_many_contours(ctx, port):
    """Shapes beyond the shared-memory record budget (global prepared records) and beyond the per-thread contour store
    (global scratch): a 12-contour, 360-segment shape."""
    big = prep.prepare(synth_big_shape(n_contours=12, segs_per_contour=30, seed=5))
    cb, pts, od, cl, _ = big.glyph(0)
    xf = glyph_transform(48, px_range=6.0)
    for mode_name in ("sdf", "psdf", "msdf", "mtsdf"):
        for ov in (True, False):
            cfg = api.MSDFGeneratorConfig(overlapSupport=ov)
            out = np.empty((48, 48, api.CHANNELS[MODE[mode_name]]), np.float32)
            ctx.generate(MODE[mode_name], out, big, xf, cfg)
            want = port.generate(mode_name, cb, pts, od, cl, xf, 48, 48, _ocfg(cfg))
            assert_parity(out, want, f"big {mode_name} overlap={ov}")


def test_big_shape_block_lists_vs_port(ctx, port):
    """Tile-level culling of big shapes (k1_big_blocks: 32x32 blocks with per-block record lists): every mode and combiner at
    sizes that are not multiples of the block, flipped output, a batch mixing big and small shapes, anisotropic projection."""
    big = prep.prepare(synth_big_shape(n_contours=30, segs_per_contour=40, seed=11))
    assert big.n_segments == 1200
    cb, pts, od, cl, _ = big.glyph(0)
    w, h = 100, 70
    xf = glyph_transform(w, h, px_range=5.0)
    for mode_name in ("sdf", "psdf", "msdf", "mtsdf"):
        for ov in (True, False):
            cfg = api.MSDFGeneratorConfig(overlapSupport=ov)
            for oyd in (False, True):
                out = np.full((h, w, api.CHANNELS[MODE[mode_name]]), np.nan, np.float32)
                ctx.generate(MODE[mode_name], out, big, xf, cfg, out_ydown=oyd)
                want = port.generate(mode_name, cb, pts, od, cl, xf, w, h, _ocfg(cfg), out_ydown=oyd)
                bits = assert_parity(out, want, f"blocks {mode_name} overlap={ov} ydown={oyd}")
                assert bits == 0, f"blocks {mode_name} overlap={ov}: {bits} differing floats"
    # zoomed in: the image shows a small part of the shape (most records are far from every block)
    zoom = np.array([400.0, 300.0, -0.35, -0.4, -0.01, 0.01])
    out = np.empty((64, 96, 3), np.float32)
    ctx.generate(api.MSDF, out, big, zoom)
    assert assert_parity(out, port.generate("msdf", cb, pts, od, cl, zoom, 96, 64, Config()), "blocks zoom") == 0
    # big and small shapes in one batch
    small = prep.prepare(synth_atlas(3, seed=5))
    mixed = FlatShapes.concat([small, big, small])
    outs = [np.empty((20, 20, 4), np.float32) for _ in range(3)] + [np.empty((45, 33, 4), np.float32)] + [np.empty((20, 20, 4), np.float32) for _ in range(3)]
    xfs = np.stack([glyph_transform(20)] * 3 + [glyph_transform(33, 45, px_range=3.0)] + [glyph_transform(20)] * 3)
    ctx.generate_batch(api.MTSDF, outs, mixed, xfs)
    for g in range(7):
        c2, p2, o2, l2, _ = mixed.glyph(g)
        hh, ww = outs[g].shape[:2]
        assert assert_parity(outs[g], port.generate("mtsdf", c2, p2, o2, l2, xfs[g], ww, hh, Config()), f"mixed batch glyph {g}") == 0


def test_resident_batch_equals_host_api(ctx):
    atlas = prep.prepare(synth_atlas(64, seed=8800))
    xf = glyph_transform(32)
    host = np.empty((64, 32, 32, 4), np.float32)
    ctx.generate_batch(api.MTSDF, host, atlas, xf)
    b = ctx.upload(api.MTSDF, atlas, xf, [32, 32])
    b.launch()
    flat = b.download()
    for g in range(64):
        assert np.array_equal(b.glyph(flat, g).view(np.uint32), host[g].view(np.uint32))
    b.free()


def test_full_size_atlas_properties(ctx, port):
    """BASELINE configs[1] at full size (4096 glyphs, msdf 32x32) — too much for the oracle, so: sampled glyphs against the
    oracle, batch determinism, and structural properties that hold for every glyph."""
    n, w = 4096, 32
    atlas = prep.prepare(synth_atlas(n))
    xf = glyph_transform(w)
    msdf = np.empty((n, w, w, 3), np.float32)
    ctx.generate_batch(api.MSDF, msdf, atlas, xf)
    again = np.empty_like(msdf)
    ctx.generate_batch(api.MSDF, again, atlas, xf)
    assert np.array_equal(msdf.view(np.uint32), again.view(np.uint32)), "non-deterministic"
    for g in list(range(0, n, 171)) + [n - 1]:
        cb, pts, od, cl, _ = atlas.glyph(g)
        assert_parity(msdf[g], port.generate("msdf", cb, pts, od, cl, xf, w, w), f"glyph {g}")
    off = api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED))
    raw = np.empty((n, w, w, 3), np.float32)
    mt = np.empty((n, w, w, 4), np.float32)
    sdf = np.empty((n, w, w, 1), np.float32)
    ctx.generate_batch(api.MSDF, raw, atlas, xf, off)
    ctx.generate_batch(api.MTSDF, mt, atlas, xf, off)
    ctx.generate_batch(api.SDF, sdf, atlas, xf)
    # MTSDF = MSDF colour channels + the true distance in alpha (edge-selectors.cpp:252-260)
    assert np.array_equal(mt[..., :3].view(np.uint32), raw.view(np.uint32))
    # with the simple combiner the alpha channel is exactly the SDF (every edge is coloured); the overlapping combiner
    # picks contours by the median of r,g,b for MTSDF and by the true distance for SDF, so there they may differ
    simple_off = api.MSDFGeneratorConfig(overlapSupport=False, errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED))
    mt_s, sdf_s = np.empty_like(mt), np.empty_like(sdf)
    ctx.generate_batch(api.MTSDF, mt_s, atlas, xf, simple_off)
    ctx.generate_batch(api.SDF, sdf_s, atlas, xf, simple_off)
    assert np.array_equal(mt_s[..., 3:].view(np.uint32), sdf_s.view(np.uint32))
    assert (mt[..., 3:] != sdf).mean() < 0.01
    # error correction only ever replaces a texel's three channels by their median, and never changes that median
    med = lambda a: np.sort(a, axis=-1)[..., 1]
    changed = (msdf.view(np.uint32) != raw.view(np.uint32)).any(axis=-1)
    assert np.array_equal(med(msdf), med(raw))
    assert (msdf[changed] == med(raw)[changed][:, None]).all()
    assert 0.005 < changed.mean() < 0.2
    # the sign of the median is the inside/outside classification of the true distance
    inside = sdf[..., 0] > 0.5
    m = med(raw)
    decided = np.abs(sdf[..., 0] - 0.5) > 0.02
    assert ((m > 0.5) == inside)[decided].mean() > 0.995
    # flipping the output orientation flips rows, nothing else
    flipped = np.empty_like(raw[:256])
    sub = atlas.slice(0, 256)
    ctx.generate_batch(api.MSDF, flipped, sub, xf, off, out_ydown=np.ones(256, np.uint8))
    assert np.array_equal(flipped[:, ::-1].view(np.uint32), raw[:256].view(np.uint32))


def test_mtsdf64_config3_sample(ctx, port):
    """BASELINE configs[2] (mtsdf 64x64 with error correction) on a sample of glyphs against the oracle."""
    atlas = prep.prepare(synth_atlas(48, seed=0xB200 + 4096))
    xf = glyph_transform(64)
    out = np.empty((48, 64, 64, 4), np.float32)
    ctx.generate_batch(api.MTSDF, out, atlas, xf)
    for g in range(0, 48, 5):
        cb, pts, od, cl, _ = atlas.glyph(g)
        assert_parity(out[g], port.generate("mtsdf", cb, pts, od, cl, xf, 64, 64), f"mtsdf64 glyph {g}")


def test_full_atlas_against_compiled_reference(ctx):
    """BASELINE configs[1] and [2] at FULL size against the unmodified reference (possible because the compiled reference
    runs glyph-parallel on the box's host cores in about a second): every one of the 12.6 M (msdf 32x32) and 67 M
    (mtsdf 64x64 = BASELINE configs[2] at full size, 268 M floats) floats within 1e-4, identical bytes; the number of differing bit patterns is printed."""
    from oracle import bindings
    if not bindings.ref_available(kind="reference_gp"):
        pytest.skip("compiled reference (oracle/_ref) did not travel to this box")
    ref = bindings.Oracle("reference_gp")
    n = 4096
    atlas = prep.prepare(synth_atlas(n))
    for mode_name, mode, w, count in (("msdf", api.MSDF, 32, n), ("mtsdf", api.MTSDF, 64, n), ("psdf", api.PSDF, 32, n)):
        sub = atlas.slice(0, count)
        xf = glyph_transform(w)
        got = np.empty((count, w, w, api.CHANNELS[mode]), np.float32)
        ctx.generate_batch(mode, got, sub, xf)
        want, _ = ref.generate_batch(mode_name, sub.gbegin, sub.cbegin, sub.points, sub.orders, sub.colors, xf, w, w, glyph_parallel=True)
        bits = assert_parity(got, want, f"full {mode_name} {w}x{w}")
        print(f"[parity] {mode_name} {w}x{w} x{count}: {bits} of {got.size} floats differ in bit pattern, max |delta| {compare(got, want)[1]:.3g}")


def test_config4_big_shape_tiles_and_oracle_samples(ctx):
    """BASELINE configs[3] (one 10k-segment shape, sdf) at reduced resolution: row bands generated as regions are bit-identical
    to the untiled image, and random texels match ShapeDistanceFinder::oneShotDistance of the oracle (the full generator is
    too slow on the CPU at this size — SURVEY.md App. B.9)."""
    from oracle import bindings
    from msdfgen_b200 import tiled
    shape = prep.prepare(synth_big_shape(100, 100))
    assert shape.n_segments == 10000
    size = 192
    xf = glyph_transform(size, px_range=8.0)
    full = np.empty((size, size, 1), np.float32)
    ctx.generate(api.SDF, full, shape, xf)
    rows = []
    for r in range(3):
        b, (r0, r1) = tiled.generate_band(ctx, api.SDF, shape, xf, size, size, r, 3)
        assert b.shape[0] == r1 - r0
        rows.append(b)
    assert np.array_equal(np.concatenate(rows).view(np.uint32), full.view(np.uint32))
    oracle = bindings.Oracle("reference" if bindings.ref_available() else "port")
    rng = np.random.default_rng(4)
    xs, ys = rng.integers(0, size, 400), rng.integers(0, size, 400)
    t = xf
    pts = np.stack([(xs + .5) / t[0] - t[2], (ys + .5) / t[1] - t[3]], axis=1)
    cb, p, od, cl, _ = shape.glyph(0)
    d = oracle.distance("sdf", True, cb, p, od, cl, pts)[:, 0]
    want = (np.float64(1.0) / (t[5] - t[4]) * (d + (-t[4]))).astype(np.float32)
    got = full[ys, xs, 0]
    assert_parity(got, want, "config-4 samples")
    # psdf too (perpendicular selector, many contours -> global contour store)
    fullp = np.empty((64, 64, 1), np.float32)
    xfp = glyph_transform(64, px_range=8.0)
    ctx.generate(api.PSDF, fullp, shape, xfp)
    xs, ys = rng.integers(0, 64, 200), rng.integers(0, 64, 200)
    pts = np.stack([(xs + .5) / xfp[0] - xfp[2], (ys + .5) / xfp[1] - xfp[3]], axis=1)
    d = oracle.distance("psdf", True, cb, p, od, cl, pts)[:, 0]
    want = (np.float64(1.0) / (xfp[5] - xfp[4]) * (d + (-xfp[4]))).astype(np.float32)
    assert_parity(fullp[ys, xs, 0], want, "config-4 psdf samples")


def test_config4_full_size_bands_and_oracle(ctx):
    """BASELINE configs[3] at FULL size (10k segments, sdf 4096x4096, default overlap support): the eight row bands of an 8-GPU run
    (generated here one after the other as regions) are bit-identical to the untiled image; 100 000 random texels per band-set
    equal ShapeDistanceFinder::oneShotDistance of the compiled reference (/root/reference/core/ShapeDistanceFinder.hpp:36-58,
    SURVEY.md App. B.9); and a full 1024x1024 image equals the reference's own OpenMP generateSDF texel for texel."""
    from oracle import bindings
    from msdfgen_b200 import tiled
    if not bindings.ref_available(kind="reference_gp"):
        pytest.skip("compiled reference (oracle/_ref) did not travel to this box")
    shape = prep.prepare(synth_big_shape(100, 100))
    size, world = 4096, 8
    xf = glyph_transform(size, px_range=8.0)
    batch = ctx.upload(api.SDF, shape, xf, [size, size])
    batch.launch(api.MSDFGeneratorConfig())
    full = batch.download()[: size * size].reshape(size, size)
    batch.free()
    for r in range(world):
        b, (r0, r1) = tiled.generate_band(ctx, api.SDF, shape, xf, size, size, r, world)
        assert np.array_equal(b[..., 0].view(np.uint32), full[r0:r1].view(np.uint32)), f"band {r} differs from the untiled image"
    gp = bindings.Oracle("reference_gp")
    rng = np.random.default_rng(44)
    n = 100000
    xs, ys = rng.integers(0, size, n), rng.integers(0, size, n)
    pts = np.stack([(xs + .5) / xf[0] - xf[2], (ys + .5) / xf[1] - xf[3]], axis=1)
    cb, p, od, cl, _ = shape.glyph(0)
    d = gp.distance("sdf", True, cb, p, od, cl, pts)[:, 0]
    want = (np.float64(1.0) / (xf[5] - xf[4]) * (d + (-xf[4]))).astype(np.float32)
    bits = assert_parity(full[ys, xs], want, "config-4 full-size samples")
    print(f"[parity] config 4, 4096x4096: {bits} of {n} sampled texels differ in bit pattern")
    # one whole image by the reference's own row-parallel OpenMP generator
    small = 1024
    xfs = glyph_transform(small, px_range=8.0)
    got = np.empty((small, small, 1), np.float32)
    ctx.generate(api.SDF, got, shape, xfs)
    want = bindings.Oracle("reference_omp").generate("sdf", cb, p, od, cl, xfs, small, small)
    bits = assert_parity(got, want, "config-4 1024x1024 full image")
    print(f"[parity] config 4, 1024x1024 full image: {bits} of {got.size} floats differ in bit pattern")


def test_config5_shard_of_the_65536_glyph_atlas(ctx):
    """BASELINE configs[4]: the 65536-glyph atlas sharded over 8 ranks exactly as bench.py does it (cost-balanced contiguous
    ranges from the glyphs' segment counts).  The ranges tile the atlas; one whole shard (rank 5 of 8, ~8192 glyphs) is generated
    and compared with the compiled reference float for float."""
    from oracle import bindings
    from msdfgen_b200 import shard
    from msdfgen_b200.synth import glyph_segment_counts, synth_atlas_range
    if not bindings.ref_available(kind="reference_gp"):
        pytest.skip("compiled reference (oracle/_ref) did not travel to this box")
    n, world, rank, w = 65536, 8, 5, 32
    costs = glyph_segment_counts(n, 0xB200) * w * w
    b = shard.partition(costs, world)
    assert b[0] == 0 and b[-1] == n and (np.diff(b) > 0).all()
    per = np.array([costs[b[r]:b[r + 1]].sum() for r in range(world)], dtype=np.float64)
    assert per.max() / per.mean() < 1.01          # balanced to 1 %
    g0, g1 = int(b[rank]), int(b[rank + 1])
    atlas = prep.prepare(synth_atlas_range(g0, g1, 0xB200))
    xf = glyph_transform(w)
    got = np.empty((g1 - g0, w, w, 3), np.float32)
    ctx.generate_batch(api.MSDF, got, atlas, xf)
    want, _ = bindings.Oracle("reference_gp").generate_batch("msdf", atlas.gbegin, atlas.cbegin, atlas.points, atlas.orders, atlas.colors, xf, w, w,
                                                             glyph_parallel=True)
    bits = assert_parity(got, want, f"config-5 shard {rank}/{world}")
    print(f"[parity] config 5 shard {rank}/{world} ({g1 - g0} glyphs): {bits} of {got.size} floats differ in bit pattern")


def test_misaligned_pinned_block_and_bad_orders(ctx):
    """A packed pinned block that does not start on a 16-byte boundary must not take the float4 zero-copy path;
    segment orders outside 1..3 are rejected on the host."""
    atlas = prep.prepare(synth_atlas(64, seed=99))
    xf = glyph_transform(32)
    want = np.empty((64, 32, 32, 3), np.float32)
    ctx.generate_batch(api.MSDF, want, atlas, xf)
    raw = ctx.pinned((64 * 32 * 32 * 3 + 4,), np.float32)
    for off in (1, 2, 3):
        view = raw[off:off + 64 * 32 * 32 * 3].reshape(64, 32, 32, 3)
        view[...] = np.nan
        ctx.generate_batch(api.MSDF, view, atlas, xf)
        assert np.array_equal(np.asarray(view).view(np.uint32), want.view(np.uint32)), off
    bad = FlatShapes(atlas.gbegin, atlas.cbegin, atlas.points, atlas.orders.copy(), atlas.colors, atlas.ydown)
    bad.orders[5] = 0
    with pytest.raises(Exception):
        ctx.generate_batch(api.MSDF, want, bad, xf)
    bad.orders[5] = 4
    with pytest.raises(Exception):
        ctx.generate_batch(api.MSDF, want, bad, xf)


def test_small_glyphs_with_many_contours(ctx, port):
    """Glyphs that fit the shared-memory record path (<= 96 segments) but have more contours than the per-thread contour store
    holds (4): the overlapping combiner's parked state goes through the global scratch slices.  Every mode, both combiners."""
    shapes = [prep.prepare(synth_big_shape(8, 8, seed=s)) for s in (1, 2)] + [prep.prepare(synth_big_shape(12, 7, seed=3))]
    for shape in shapes:
        assert shape.n_segments <= 96 and shape.n_contours > 4
        cb, p, od, cl, _ = shape.glyph(0)
        for mode_name, mode in MODE.items():
            for overlap in (True, False):
                xf = glyph_transform(40, px_range=6.0)
                cfg = api.MSDFGeneratorConfig(overlapSupport=overlap)
                got = np.full((40, 40, api.CHANNELS[mode]), np.nan, np.float32)
                ctx.generate(mode, got, shape, xf, cfg)
                assert_parity(got, port.generate(mode_name, cb, p, od, cl, xf, 40, 40, _ocfg(cfg)), f"{mode_name} overlap={overlap}")


def test_regions_reject_error_correction(ctx):
    g = prep.prepare(synth_glyph(3))
    b = ctx.upload(api.MSDF, g, glyph_transform(16), [16, 8], origins=[0, 4])
    with pytest.raises(Exception):
        b.launch(api.MSDFGeneratorConfig())
    b.launch(api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED)))
    part = b.download()[: 8 * 16 * 3].reshape(8, 16, 3)
    full = np.empty((16, 16, 3), np.float32)
    ctx.generate(api.MSDF, full, g, glyph_transform(16), api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED)))
    assert np.array_equal(part.view(np.uint32), full[4:12].view(np.uint32))
    b.free()


def test_zero_copy_pinned_output_equals_copied_output(ctx):
    """Packed pinned host output (b200msdf_host_alloc) takes the zero-copy path (kernels store straight into host memory);
    pageable output takes the staged D2H path. Same texels either way, for every mode and with error correction on."""
    atlas = prep.prepare(synth_atlas(300, seed=777))
    xf = glyph_transform(32)
    for mode in (api.SDF, api.MSDF, api.MTSDF):
        n = api.CHANNELS[mode]
        pageable = np.full((300, 32, 32, n), np.nan, np.float32)
        pinned = ctx.pinned((300, 32, 32, n), np.float32)
        pinned[...] = np.nan
        ctx.generate_batch(mode, pageable, atlas, xf)
        ctx.generate_batch(mode, pinned, atlas, xf)
        assert np.array_equal(pageable.view(np.uint32), np.asarray(pinned).view(np.uint32))


def test_u8_atlas_output_equals_quantised_float_output(ctx, port):
    """b200msdf_generate_batch_u8: fused pixelFloatToByte + placement into rectangles of an atlas page, for pageable pages
    (packed device buffer + host scatter) and pinned pages (kernel writes the rectangles directly)."""
    atlas = prep.prepare(synth_atlas(40, seed=4242))
    xf = glyph_transform(32)
    for mode in (api.SDF, api.MSDF, api.MTSDF):
        n = api.CHANNELS[mode]
        ref_f = np.empty((40, 32, 32, n), np.float32)
        ctx.generate_batch(mode, ref_f, atlas, xf)
        want = port.float_to_byte(ref_f)
        assert np.array_equal(want, to_bytes(ref_f))
        for pinned in (False, True):
            page = ctx.pinned((5 * 36, 8 * 36 * n + 7), np.uint8) if pinned else np.empty((5 * 36, 8 * 36 * n + 7), np.uint8)
            page[...] = 123
            rects = [page[(g // 8) * 36 + 2:(g // 8) * 36 + 34, (g % 8) * 36 * n + 3:(g % 8) * 36 * n + 3 + 32 * n].reshape(32, 32, n) for g in range(40)]
            ctx.generate_batch_u8(mode, rects, atlas, xf)
            for g in range(40):
                assert np.array_equal(rects[g], want[g]), (mode, pinned, g)
            assert (np.asarray(page)[0:2] == 123).all() and (np.asarray(page)[:, :3] == 123).all()
