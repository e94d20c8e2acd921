import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_atlas, glyph_transform
NG = int(os.environ.get('NG', 4096)); W = 32
atlas = prep.prepare(synth_atlas(NG))
ctx = api.Context(0)
torch.cuda.init(); s = torch.cuda.current_stream(); ctx.set_stream(s.cuda_stream)
xf = glyph_transform(W)
batch = ctx.upload(api.MSDF, atlas, xf, [W, W])
E = api.ErrorCorrectionConfig
cfgs = {'default(2,1)': E(2,1), 'prio+nocheck(2,0)': E(2,0), 'indisc+nocheck(1,0)': E(1,0), 'edgeonly+nocheck(3,0)': E(3,0), 'prio+always(2,2)': E(2,2), 'edgeonly+atedge(3,1)': E(3,1)}
for name, ec in cfgs.items():
    cfg = api.MSDFGeneratorConfig(errorCorrection=ec)
    batch.launch(cfg); torch.cuda.synchronize()
    ts = []
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); batch.launch_stages(4, cfg); b.record(s); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print('%-24s k2 ms %.3f' % (name, np.median(ts)), flush=True)
