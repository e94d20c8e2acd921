"""Writes tests/golden/fuzz_v1.json: SHA-256 digests of the UNMODIFIED reference's output (oracle/_ref) on the adversarial
shape families of tests/util.py::fuzz_shape — generators (all modes, both combiners, error correction on/off), the scanline
sign correction and rasterize.  Digests pin bit-exact parity without storing 2000 bitmaps.
Run in the build container (needs /root/reference):   python tests/golden/make_fuzz_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.bindings import Oracle, Config  # noqa: E402
from util import fuzz_shape, fuzz_configs, digest  # noqa: E402

ref = Oracle("reference")
out = {}
for seed in range(108):
    fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
    for key, mode, cfg in fuzz_configs(seed):
        f = ref.generate(mode, cb, pts, od, cl, xf, w, h, cfg)
        out[f"{seed}/{key}"] = digest(f)
        if key.endswith("ec0"):
            rule = seed % 4
            out[f"{seed}/{key}/sign{rule}"] = digest(ref.sign_correction(f, cb, pts, od, xf, 0.5, rule))
    out[f"{seed}/raster"] = digest(ref.rasterize(cb, pts, od, xf, w, h, seed % 4))
path = os.path.join(ROOT, "tests", "golden", "fuzz_v1.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=0, sort_keys=True)
print(len(out), "digests ->", path)
