"""One launch of the big-shape distance kernel (for ncu): 10k-segment shape, overlapping combiner; MODE/SIZE from the env."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_big_shape, glyph_transform
mode = int(os.environ.get('MODE', 1)); size = int(os.environ.get('SIZE', 1024))
shape = prep.prepare(synth_big_shape(100, 100))
ctx = api.Context(0)
b = ctx.upload(mode, shape, glyph_transform(size, px_range=8.0), [size, size])
cfg = api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=0))
for _ in range(2):
    b.launch(cfg)
ctx.synchronize()
