"""Writes tests/golden/dropin_v1.npz: outputs of tests/dropin/caller.cpp built against the UNMODIFIED reference.
Run in the build container (needs /root/reference):   python tests/golden/make_dropin_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_dropin  # noqa: E402
from msdfgen_b200.synth import synth_glyph  # noqa: E402

lib = test_dropin._build_ref()
seeds = np.array([501, 502, 503, 504, 505, 506], dtype=np.int64)
arrays = {"seeds": seeds}
for i, seed in enumerate(seeds):
    r = test_dropin._run(lib, synth_glyph(int(seed), cubic_fraction=0.2))
    arrays[f"edges{i}"] = np.int64(r["edges"])
    for key in ("msdf", "modes", "ec", "stencil", "scanline", "render", "sdferr"):
        arrays[f"{key}{i}"] = r[key]
# caller_threads: four threads, sizes (24+t) x (20+t), colouring seed t — the reference's images for the re-entrancy test
for t, img in enumerate(test_dropin._threads(lib, synth_glyph(321, cubic_fraction=0.2), 4, 1)):
    arrays[f"threads{t}"] = img
path = os.path.join(ROOT, "tests", "golden", "dropin_v1.npz")
np.savez_compressed(path, **arrays)
print(path, os.path.getsize(path) // 1024, "KiB")
