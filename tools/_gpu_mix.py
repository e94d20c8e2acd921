import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_atlas, glyph_transform
ctx = api.Context(0); torch.cuda.init(); s = torch.cuda.current_stream(); ctx.set_stream(s.cuda_stream)
W=32
for name, kw in (('mixed', {}), ('all-linear', dict(linear_fraction=1.0)), ('all-quadratic', dict(linear_fraction=0.0)), ('all-cubic', dict(linear_fraction=0.0, cubic_fraction=1.0))):
    atlas = prep.prepare(synth_atlas(2048, **kw))
    for mode, mname in ((api.MSDF,'msdf'), (api.SDF,'sdf')):
        for ov in (True, False):
            batch = ctx.upload(mode, atlas, glyph_transform(W), [W, W])
            cfg = api.MSDFGeneratorConfig(overlapSupport=ov)
            batch.launch(cfg); torch.cuda.synchronize()
            ts=[]
            for i in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s); batch.launch_stages(2, cfg); b.record(s); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            t = np.median(ts); pairs = atlas.n_segments*W*W
            print(f'{name:14s} {mname:5s} overlap={int(ov)} k1 {t:.3f} ms  {t*1e6/pairs*1e3:.2f} ps per segment-texel', flush=True)
            batch.free()
