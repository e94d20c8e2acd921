"""Lists every fuzz case where the CUDA path differs from the C restatement (run by hand on the GPU box)."""
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from util import fuzz_shape, fuzz_configs
from oracle.bindings import Oracle
from msdfgen_b200 import api
from msdfgen_b200.flat import FlatShapes
port = Oracle('port'); ctx = api.Context(0)
MODE = {"sdf": api.SDF, "psdf": api.PSDF, "msdf": api.MSDF, "mtsdf": api.MTSDF}
bad = []
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 96):
    fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
    shape = FlatShapes.single(cb, pts, od, cl)
    for key, mode, ocfg in fuzz_configs(seed):
        cfg = api.MSDFGeneratorConfig(overlapSupport=ocfg.overlap, errorCorrection=api.ErrorCorrectionConfig(mode=ocfg.ec_mode, distanceCheckMode=ocfg.ec_dcm))
        f = np.full((h, w, api.CHANNELS[MODE[mode]]), np.nan, np.float32)
        ctx.generate(MODE[mode], f, shape, xf, cfg)
        want = port.generate(mode, cb, pts, od, cl, xf, w, h, ocfg)
        ne = f.view(np.uint32) != want.view(np.uint32)
        if ne.any():
            ys, xs, cs = np.nonzero(ne)
            d = np.abs(f.astype(np.float64) - want.astype(np.float64))
            bad.append(dict(seed=seed, fam=fam, key=key, n=int(ne.sum()), maxabs=float(np.nanmax(np.where(ne, d, 0))),
                            first=[int(xs[0]), int(ys[0]), int(cs[0]), float(f[ys[0], xs[0], cs[0]]), float(want[ys[0], xs[0], cs[0]])]))
for b in bad:
    print(b)
print(len(bad), "differing cases")
json.dump(bad, open('/root/repo/gpurun_out/fuzz_bad.json', 'w'))
