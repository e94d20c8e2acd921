import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from oracle.bindings import Oracle
from msdfgen_b200 import api
from msdfgen_b200.synth import synth_atlas, glyph_transform
from msdfgen_b200.flat import FlatShapes
NG = int(os.environ.get('NG', 4096)); W = int(os.environ.get('W', 32)); MODE = int(os.environ.get('MODE', 3))
ref = Oracle('reference')
t0 = time.time()
raw = synth_atlas(NG)
parts = []
for g in range(NG):
    cb, pts, od, cl, yd = raw.glyph(g)
    parts.append(FlatShapes.single(*ref.prepare(cb, pts, od)))
atlas = FlatShapes.concat(parts)
print('prep %.1fs  segs %d  contours %d' % (time.time()-t0, atlas.n_segments, atlas.n_contours), flush=True)
ctx = api.Context(0)
torch.cuda.init(); s = torch.cuda.current_stream()
ctx.set_stream(s.cuda_stream)
xf = glyph_transform(W)
batch = ctx.upload(MODE, atlas, xf, [W, W])
cfgs = {'default': api.MSDFGeneratorConfig(), 'ec_off': api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=0)),
        'simple_ec_off': api.MSDFGeneratorConfig(overlapSupport=False, errorCorrection=api.ErrorCorrectionConfig(mode=0))}
for name, cfg in cfgs.items():
    for _ in range(3): batch.launch(cfg)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record(s)
    for i in range(10):
        batch.launch(cfg); ev[i+1].record(s)
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i+1]) for i in range(10)]
    px = NG*W*W
    print(name, 'ms/launch median %.3f min %.3f  -> %.3e px/s' % (np.median(ts), min(ts), px/(np.median(ts)*1e-3)), flush=True)
# e2e
out = np.empty((NG, W, W, api.CHANNELS[MODE]), np.float32)
for _ in range(2): ctx.generate_batch(MODE, out, atlas, xf)
t0 = time.time(); 
for _ in range(5): ctx.generate_batch(MODE, out, atlas, xf)
dt = (time.time()-t0)/5
print('e2e pageable host buffers %.3f ms -> %.3e px/s' % (dt*1e3, NG*W*W/dt))
