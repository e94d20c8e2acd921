import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_atlas, synth_big_shape, glyph_transform
ctx = api.Context(0)
atlas = prep.prepare(synth_atlas(24, seed=11, cubic_fraction=0.2))
for mode in (api.SDF, api.PSDF, api.MSDF, api.MTSDF):
    for ov in (True, False):
        out = np.empty((24, 20, 28, api.CHANNELS[mode]), np.float32)
        ctx.generate_batch(mode, out, atlas, glyph_transform(24), api.MSDFGeneratorConfig(overlapSupport=ov, errorCorrection=api.ErrorCorrectionConfig(2, 2)))
big = prep.prepare(synth_big_shape(9, 20))
out = np.empty((40, 40, 3), np.float32)
ctx.generate(api.MSDF, out, big, glyph_transform(40))
pin = ctx.pinned((24, 16, 16, 3), np.float32)
ctx.generate_batch(api.MSDF, pin, atlas, glyph_transform(16))
u8 = [np.empty((16, 16, 3), np.uint8) for _ in range(24)]
ctx.generate_batch_u8(api.MSDF, u8, atlas, glyph_transform(16))
print('done', float(np.nanmax(out)))
