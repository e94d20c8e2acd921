"""Writes tests/golden/hostscan_v1.npz: everything tests/dropin/host_caller.cpp computes when built against the UNMODIFIED
reference (Scanline, scanlineIntersections, Shape::scanline / orientContours, scanlineSDF, estimateSDFError, solvers).
Run in the build container (needs /root/reference):   python tests/golden/make_hostscan_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hostscan  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402

fields = hostscan.fields_from_oracle(Oracle("reference"))
res = hostscan.run(hostscan.build("ref"), fields)
for n, f in fields.items():
    res[f"field{n}"] = f
path = os.path.join(ROOT, "tests", "golden", "hostscan_v1.npz")
np.savez_compressed(path, **res)
print(len(res), "arrays ->", path, os.path.getsize(path) // 1024, "KiB")
