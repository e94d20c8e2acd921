"""A/B timing of kernel variants: python tests/_gpu_ab.py lib1.so lib2.so ... (each in a fresh process)."""
import sys, os, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] != '--child':
    for lib in sys.argv[1:]:
        env = dict(os.environ, B200MSDF_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, __file__, '--child'], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])
    sys.exit(0)
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from msdfgen_b200 import api, prep
from msdfgen_b200.synth import synth_atlas, glyph_transform
NG = int(os.environ.get('NG', 4096)); W = 32
atlas = prep.prepare(synth_atlas(NG))
ctx = api.Context(0)
torch.cuda.init(); s = torch.cuda.current_stream(); ctx.set_stream(s.cuda_stream)
batch = ctx.upload(api.MSDF, atlas, glyph_transform(W), [W, W])
cfg = api.MSDFGeneratorConfig()
batch.launch(cfg); torch.cuda.synchronize()
ref = batch.download().copy()
res = {}
for name, mask in (('k1', 2), ('k2k3', 4), ('all', 7)):
    ts = []
    for i in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); batch.launch_stages(mask, cfg); b.record(s); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    res[name] = round(float(np.median(ts)), 3)
res['checksum'] = int(ref.view(np.uint32).astype(np.uint64).sum())
print(json.dumps(res))
