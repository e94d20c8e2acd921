import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle.bindings import Oracle, Config
from msdfgen_b200 import api
from msdfgen_b200.synth import synth_glyph, glyph_transform
from msdfgen_b200.flat import FlatShapes
sys.path.insert(0, '/root/repo/tests'); from util import compare
ref = Oracle('reference')
ctx = api.Context(0)
for mode_name, mode in [('sdf',1),('psdf',2),('msdf',3),('mtsdf',4)]:
    for ov in [True, False]:
        nd = nb = n = 0; mx = 0.0; stm = 0
        for i in range(24):
            g = synth_glyph(2000+i, cubic_fraction=0.2 if i%3==0 else 0.0)
            cb, pts, od, cl = ref.prepare(g.cbegin, g.points, g.orders)
            fs = FlatShapes.single(cb, pts, od, cl)
            W = 32 if i%2==0 else 24
            xf = glyph_transform(W)
            a, sa = ref.generate(mode_name, cb, pts, od, cl, xf, W, W, Config(overlap=ov), want_stencil=True)
            out = np.full((W, W, api.CHANNELS[mode]), np.nan, np.float32)
            st = np.zeros((W,W), np.uint8)
            ctx.generate(mode, out, fs, xf, api.MSDFGeneratorConfig(overlapSupport=ov), ec_buffer=st if mode>=3 else None)
            db, m, over = compare(a, out)
            nd += db; nb += over; n += a.size; mx = max(mx, m)
            if mode >= 3: stm += int((sa != st).sum())
        print(mode_name, 'overlap', ov, 'diff-bits', nd, '/', n, 'max', mx, '>1e-4:', nb, 'stencil mism', stm, flush=True)
