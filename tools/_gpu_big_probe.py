"""BASELINE configs[3] timing (one 10k-segment shape, big images); run by hand on the GPU box, not a test."""
import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_big_shape, glyph_transform
shape = prep.prepare(synth_big_shape(100, 100))
ctx = api.Context(0)
torch.cuda.init(); s = torch.cuda.current_stream()
ctx.set_stream(s.cuda_stream)
for mode, size, ov in ((api.SDF, 2048, True), (api.SDF, 4096, True), (api.SDF, 2048, False), (api.PSDF, 2048, True), (api.MSDF, 1024, True), (api.MTSDF, 1024, True)):
    xf = glyph_transform(size, px_range=8.0)
    cfg = api.MSDFGeneratorConfig(overlapSupport=ov, errorCorrection=api.ErrorCorrectionConfig(mode=0))
    b = ctx.upload(mode, shape, xf, [size, size])
    b.launch(cfg); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); b.launch(cfg); e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"mode {mode} {size}^2 overlap={ov}: {ms:.2f} ms  {size*size/(ms*1e-3):.3e} px/s", flush=True)
    b.free()
