"""Adversarial-shape parity (tests/util.py::fuzz_shape): degenerate curves, ties on the texel grid, zero-length edges, huge and
tiny coordinates, self-intersections.  The committed digests (tests/golden/fuzz_v1.json) are the compiled reference's outputs;
the CPU test pins the C restatement to them, the GPU test pins the CUDA path (through the C ABI) to them — bit for bit."""
import json
import os

import numpy as np
import pytest

from util import fuzz_shape, fuzz_configs, digest

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_v1.json")))
SEEDS = range(108)
HAIRPIN = 8      # fuzz family (quadratic cusps): the reference's texels at a cusp hinge on the last bit of libm's acos / cos / pow;
                 # the device's literal solver uses correctly rounded versions (csrc/cr_math.cuh) and is held to the same digests


def test_port_matches_reference_digests(port):
    bad = []
    for seed in SEEDS:
        fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
        for key, mode, cfg in fuzz_configs(seed):
            f = port.generate(mode, cb, pts, od, cl, xf, w, h, cfg)
            if digest(f) != GOLDEN[f"{seed}/{key}"]:
                bad.append((seed, fam, key))
            if key.endswith("ec0"):
                rule = seed % 4
                if digest(port.sign_correction(f, cb, pts, od, xf, 0.5, rule)) != GOLDEN[f"{seed}/{key}/sign{rule}"]:
                    bad.append((seed, fam, key, "sign"))
        if digest(port.rasterize(cb, pts, od, xf, w, h, seed % 4)) != GOLDEN[f"{seed}/raster"]:
            bad.append((seed, fam, "raster"))
    assert not bad, bad[:10]


def test_digests_are_current_reference(ref):
    for seed in list(SEEDS)[::7]:
        fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
        for key, mode, cfg in fuzz_configs(seed):
            assert digest(ref.generate(mode, cb, pts, od, cl, xf, w, h, cfg)) == GOLDEN[f"{seed}/{key}"]


@pytest.mark.gpu
def test_cuda_matches_reference_digests(ctx, port):
    from msdfgen_b200 import api
    from msdfgen_b200.flat import FlatShapes
    MODE = {"sdf": api.SDF, "psdf": api.PSDF, "msdf": api.MSDF, "mtsdf": api.MTSDF}
    bad = []
    for seed in SEEDS:
        fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
        shape = FlatShapes.single(cb, pts, od, cl)
        for key, mode, ocfg in fuzz_configs(seed):
            cfg = api.MSDFGeneratorConfig(overlapSupport=ocfg.overlap, errorCorrection=api.ErrorCorrectionConfig(mode=ocfg.ec_mode, distanceCheckMode=ocfg.ec_dcm))
            f = np.full((h, w, api.CHANNELS[MODE[mode]]), np.nan, np.float32)
            ctx.generate(MODE[mode], f, shape, xf, cfg)
            if digest(f) != GOLDEN[f"{seed}/{key}"]:
                want = port.generate(mode, cb, pts, od, cl, xf, w, h, ocfg)
                bad.append((seed, fam, key, int((f.view(np.uint32) != want.view(np.uint32)).sum()), float(np.nanmax(np.abs(f - want)))))
            if key.endswith("ec0"):
                rule = seed % 4
                g = f.copy()
                ctx.sign_correction([g], shape, xf[:4], 0.5, rule)
                if digest(g) != GOLDEN[f"{seed}/{key}/sign{rule}"]:
                    bad.append((seed, fam, key, "sign"))
        r = np.full((h, w, 1), np.nan, np.float32)
        ctx.rasterize([r], shape, xf[:4], seed % 4)         # (scanline crossings of a hairpin are exact: quadratic formula)
        if digest(r) != GOLDEN[f"{seed}/raster"]:
            bad.append((seed, fam, "raster"))
    assert not bad, (len(bad), bad[:12])


@pytest.mark.gpu
def test_literal_solver_is_rare_on_ordinary_glyphs_and_used_on_cusps(ctx):
    """The fast foot-point root path falls back to the literal (correctly rounded) solver only inside its guard band: a tiny
    fraction of the edge evaluations of ordinary glyphs.  Cusp segments always take the literal solver."""
    from msdfgen_b200 import api, prep
    from msdfgen_b200.flat import FlatShapes
    from msdfgen_b200.synth import synth_atlas, glyph_transform
    atlas = prep.prepare(synth_atlas(256, seed=31337))
    out = np.empty((256, 32, 32, 3), np.float32)
    ctx.exact_root_calls(reset=True)
    ctx.generate_batch(api.MSDF, out, atlas, glyph_transform(32))
    ordinary = ctx.exact_root_calls(reset=True)
    quadratic_evals_at_least = 256 * 32 * 32          # every texel evaluates at least one quadratic edge of these glyphs, typically several
    print(f"[guard band] literal-solver runs on 256 ordinary glyphs: {ordinary} ({ordinary / quadratic_evals_at_least:.2e} per texel)")
    assert ordinary < 0.01 * quadratic_evals_at_least
    fam, cb, pts, od, cl, xf, w, h = fuzz_shape(HAIRPIN)
    f = np.empty((h, w, 1), np.float32)
    ctx.generate(api.SDF, f, FlatShapes.single(cb, pts, od, cl), xf)
    assert ctx.exact_root_calls(reset=True) > 0
    # family 1: straight-line quadratics on the integer grid with texel centres on integer coordinates — foot points fall exactly
    # on end points, i.e. roots at 0 and 1: the guard band's home turf (these shapes carry no cusp segments, so every run of the
    # literal solver counted here comes from the recheck launch; their digests are checked by test_cuda_matches_reference_digests)
    grid = 0
    for seed in (1, 10, 19, 28):
        fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
        f = np.empty((h, w, 3), np.float32)
        ctx.generate(api.MSDF, f, FlatShapes.single(cb, pts, od, cl), xf)
        grid += ctx.exact_root_calls(reset=True)
    print(f"[guard band] literal-solver runs on four grid-aligned straight-quadratic shapes: {grid}")
