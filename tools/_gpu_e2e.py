import sys, os, time, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == '--sweep':
    for ch in (1, 2, 4, 8):
        out = subprocess.run([sys.executable, __file__], env=dict(os.environ, B200MSDF_CHUNKS=str(ch)), capture_output=True, text=True)
        print('chunks', ch, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
    sys.exit(0)
sys.path.insert(0, '/root/repo')
import numpy as np
from msdfgen_b200 import api, prep
from msdfgen_b200.flat import FlatShapes
from msdfgen_b200.synth import synth_atlas, glyph_transform
atlas = prep.prepare(synth_atlas(4096)); xf = glyph_transform(32)
ctx = api.Context(0)
def pin(a):
    b = ctx.pinned(a.shape, a.dtype); b[...] = a; return b
ha = FlatShapes(pin(atlas.gbegin), pin(atlas.cbegin), pin(atlas.points), pin(atlas.orders), pin(atlas.colors), pin(atlas.ydown))
out = ctx.pinned((4096, 32, 32, 3), np.float32)
for _ in range(3): ctx.generate_batch(api.MSDF, out, ha, xf)
ts = []
for _ in range(15):
    t0 = time.perf_counter(); ctx.generate_batch(api.MSDF, out, ha, xf); ts.append(time.perf_counter() - t0)
print(json.dumps({'e2e_ms_median': round(1e3*float(np.median(ts)), 3), 'min': round(1e3*min(ts), 3)}))
