"""Writes tests/golden/render_v1.npz: renderSDF / simulate8bit outputs of the UNMODIFIED reference (oracle/_ref) for every
channel combination the reference declares (core/render-sdf.h:11-16), hard threshold and ranged mapping, up- and down-scaling.
Run in the build container (needs /root/reference):   python tests/golden/make_render_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bindings import Oracle, Config  # noqa: E402
from msdfgen_b200.synth import synth_glyph, glyph_transform  # noqa: E402

ref = Oracle("reference")
cases, arrays = [], {}
g = synth_glyph(0xDE40, cubic_fraction=0.25)
cb, pts, od, cl = ref.prepare(g.cbegin, g.points, g.orders)
fields = {}
for mode, n in (("sdf", 1), ("msdf", 3), ("mtsdf", 4)):
    fields[n] = ref.generate(mode, cb, pts, od, cl, glyph_transform(24, 20), 24, 20, Config())
    arrays[f"field{n}"] = fields[n]
COMBOS = [(1, 1), (3, 1), (1, 3), (3, 3), (1, 4), (4, 4)]          # (out channels, sdf channels)
for out_n, sdf_n in COMBOS:
    for rng in ((0.0, 0.0), (-2.0, 2.0), (-1.0, 3.0)):
        for ow, oh in ((48, 40), (13, 27), (24, 20), (150, 3)):
            for th in (0.5, 0.42):
                i = len(cases)
                cases.append(dict(out_n=out_n, sdf_n=sdf_n, lower=rng[0], upper=rng[1], ow=ow, oh=oh, threshold=th))
                arrays[f"c{i}"] = ref.render_sdf(fields[sdf_n], ow, oh, out_n, rng, th)
for n in (1, 3, 4):
    arrays[f"snap{n}"] = ref.simulate_8bit(fields[n] * 1.6 - 0.3)
arrays["index"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
path = os.path.join(ROOT, "tests", "golden", "render_v1.npz")
np.savez_compressed(path, **arrays)
print(len(cases), "cases ->", path, os.path.getsize(path) // 1024, "KiB")
