"""Stage timing of the scanline pass on the headline atlas (run by hand on the GPU box; not a test)."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_atlas, synth_big_shape, glyph_transform
NG = int(os.environ.get('NG', 4096)); W = int(os.environ.get('W', 32))
atlas = prep.prepare(synth_atlas(NG))
ctx = api.Context(0)
torch.cuda.init(); s = torch.cuda.current_stream()
ctx.set_stream(s.cuda_stream)
xf = glyph_transform(W)


def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record(s)
    for i in range(n):
        fn(); ev[i + 1].record(s)
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(n)]))


for mode in (api.MSDF, api.MTSDF, api.SDF):
    batch = ctx.upload(mode, atlas, xf, [W, W])
    scan = api.MSDFGeneratorConfig(scanlinePass=True)
    plain = api.MSDFGeneratorConfig()
    batch.launch(scan)
    t_scan = timed(lambda: batch.launch_stages(8, scan))
    t_all = timed(lambda: batch.launch(scan))
    t_plain = timed(lambda: batch.launch(plain))
    ch = api.CHANNELS[mode]
    gb = NG * W * W * ch * 4 * 2 / 1e9
    print(f"mode {mode}: scanline stage {t_scan:.3f} ms ({gb / (t_scan * 1e-3):.0f} GB/s field r+w), launch with pass {t_all:.3f} ms, without {t_plain:.3f} ms", flush=True)
    batch.free()
big = prep.prepare(synth_big_shape(100, 100))
for size in (512, 2048):
    xfb = glyph_transform(size, px_range=8.0)
    batch = ctx.upload(api.SDF, big, xfb, [size, size])
    scan = api.MSDFGeneratorConfig(scanlinePass=True)
    batch.launch(scan)
    t_scan = timed(lambda: batch.launch_stages(8, scan), n=3)
    print(f"10k-segment shape sdf {size}^2: scanline stage {t_scan:.3f} ms", flush=True)
    batch.free()
