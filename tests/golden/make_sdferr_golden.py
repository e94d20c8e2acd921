"""Writes tests/golden/sdferr_v1.npz: estimateSDFError of the UNMODIFIED reference (oracle/_ref) on fields it generated itself.
Run in the build container (needs /root/reference):   python tests/golden/make_sdferr_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import bindings  # noqa: E402
from msdfgen_b200.synth import synth_glyph, synth_big_shape  # noqa: E402

ref = bindings.Oracle("reference")
arrays, n = {}, 0
rng = np.random.default_rng(7)
for seed in range(900, 912):
    g = synth_glyph(seed, cubic_fraction=0.25) if seed % 4 else synth_big_shape(4, 40, seed=seed)
    cb, pts, od, _, _ = g.glyph(0)
    cb, pts, od, col = ref.prepare(cb, pts, od)
    mode = ("sdf", "msdf", "mtsdf")[seed % 3]
    w, h = int(rng.integers(12, 40)), int(rng.integers(12, 40))
    ydown = bool(seed % 2)
    # (no mirrored projections: with a negative x scale xTo < xFrom and the reference's Scanline::overlap never terminates)
    xf = np.array([w * (0.9 if seed % 5 else 0.6), h * 0.9, 0.05, 0.05 if seed % 7 else -0.2, -2.0 / w, 2.0 / w])
    field = ref.generate(mode, cb, pts, od, col, xf, w, h, shape_ydown=ydown)
    if seed % 3 == 1:
        field = field + rng.normal(0, 0.02, field.shape).astype(np.float32)       # a damaged field: larger errors, more crossings
    for spr in (1, 4, 10):
        rule = int(rng.integers(0, 4))
        arrays[f"cb{n}"], arrays[f"pts{n}"], arrays[f"od{n}"] = cb, pts, od
        arrays[f"xf{n}"], arrays[f"field{n}"] = xf, field
        arrays[f"meta{n}"] = np.array([spr, rule, int(ydown)], np.int64)
        arrays[f"err{n}"] = np.float64(ref.estimate_sdf_error(field, cb, pts, od, xf, spr, rule, ydown))
        n += 1
arrays["count"] = np.int64(n)
path = os.path.join(ROOT, "tests", "golden", "sdferr_v1.npz")
np.savez_compressed(path, **arrays)
print(path, os.path.getsize(path) // 1024, "KiB", n, "cases; errors", [float(arrays[f"err{i}"]) for i in range(0, n, 7)])
