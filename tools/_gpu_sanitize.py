"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck); see profiles/sanitizer_r2.txt.
    compute-sanitizer --tool memcheck python tools/_gpu_sanitize.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from msdfgen_b200 import api, prep
from msdfgen_b200.flat import FlatShapes
from msdfgen_b200.synth import synth_atlas, synth_big_shape, glyph_transform
from util import fuzz_shape

ctx = api.Context(0)
atlas = prep.prepare(synth_atlas(24, seed=11, cubic_fraction=0.2))
for mode in (api.SDF, api.PSDF, api.MSDF, api.MTSDF):
    for ov in (True, False):
        out = np.empty((24, 20, 28, api.CHANNELS[mode]), np.float32)           # 28x20: strips, partial tiles
        ctx.generate_batch(mode, out, atlas, glyph_transform(24), api.MSDFGeneratorConfig(overlapSupport=ov, errorCorrection=api.ErrorCorrectionConfig(2, 2)))
        out = np.empty((24, 16, 16, api.CHANNELS[mode]), np.float32)           # 16x16: 8x4 blocks
        ctx.generate_batch(mode, out, atlas, glyph_transform(16), api.MSDFGeneratorConfig(overlapSupport=ov))
# many contours (global contour store), big shape (block lists, folded combiner), scanline pass
big = prep.prepare(synth_big_shape(9, 20))
out = np.empty((40, 40, 3), np.float32)
ctx.generate(api.MSDF, out, big, glyph_transform(40), api.MSDFGeneratorConfig(scanlinePass=True))
big2 = prep.prepare(synth_big_shape(12, 12))
out1 = np.empty((48, 48, 1), np.float32)
ctx.generate(api.SDF, out1, big2, glyph_transform(48))
# cusps: literal solver with the correctly rounded functions; guard-band recheck launch
fam, cb, pts, od, cl, xf, w, h = fuzz_shape(8)
f = np.empty((h, w, 3), np.float32)
ctx.generate(api.MSDF, f, FlatShapes.single(cb, pts, od, cl), xf)
# zero-copy pinned output, 8-bit output, sign correction, rasterize, render, error estimate
pin = ctx.pinned((24, 16, 16, 3), np.float32)
ctx.generate_batch(api.MSDF, pin, atlas, glyph_transform(16))
u8 = [np.empty((16, 16, 3), np.uint8) for _ in range(24)]
ctx.generate_batch_u8(api.MSDF, u8, atlas, glyph_transform(16))
one = atlas.slice(0, 1)
g = np.empty((16, 16, 3), np.float32)
ctx.generate(api.MSDF, g, one, glyph_transform(16))
ctx.sign_correction([g], one, glyph_transform(16)[:4])
r = np.empty((16, 16, 1), np.float32)
ctx.rasterize([r], one, glyph_transform(16)[:4])
big_r = np.empty((40, 40, 1), np.float32)
ctx.render_sdf(big_r, g, 2.0)
err = ctx.estimate_sdf_error([g], one, glyph_transform(16), 4)
print("done", float(np.nanmax(out)), float(err[0]))
ctx.close()
