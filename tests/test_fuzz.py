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
HAIRPIN = 8      # fuzz family (cusps) whose result at the cusp is libm rounding noise in the reference (tests/util.py)


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
            if fam == HAIRPIN:
                # magnitudes only, and only where the sign is the whole story (true distance, no error correction)
                want = port.generate(mode, cb, pts, od, cl, xf, w, h, ocfg)
                if mode == "sdf" and not np.allclose(np.abs(f - 0.5), np.abs(want - 0.5), rtol=0, atol=1e-6):
                    bad.append((seed, fam, key, "magnitude"))
                # a channel without edges is -inf in both (error correction may then replace noise-dependent texels by the median)
                if key.endswith("ec0") and not np.array_equal(np.isfinite(f), np.isfinite(want)):
                    bad.append((seed, fam, key, "non-finite"))
                continue
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
