#!/usr/bin/env python
"""Long random campaigns through the host build of the kernels' texel path (tests/devcore) against the compiled reference — the
open-ended version of tests/test_devcore.py::test_random_shape_soup.

    python tools/devcore_fuzz.py [--seed 1] [--shapes 1500] [--kinds plain,grid,mixed,repeated,sliver]

Every shape goes through k1a's filter + k1b's walk (FAST_GUARDED) and the big-shape single-pass form (FOLDED) in six mode / combiner
combinations on a 24x24 frame plus 100 random points; values are compared as the floats a generate* call would write.  Prints the
totals per kind and the first mismatching cases.  Needs /root/reference or a built oracle/_ref, and g++."""
import argparse
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--shapes", type=int, default=1500)
    ap.add_argument("--kinds", default="plain,grid,mixed,repeated,sliver")
    args = ap.parse_args()
    import test_devcore as T
    from oracle.bindings import Oracle
    os.makedirs(T.OUT, exist_ok=True)
    lib = os.path.join(T.OUT, "libdevcore.so")
    subprocess.run([T.CXX, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", f"-I{T.SHIM}", f"-I{T.CSRC}",
                    "-o", lib, T.SRC], check=True)
    dev, ref = C.CDLL(lib), Oracle("reference")
    rng = np.random.default_rng(args.seed)
    kinds = args.kinds.split(",")
    totals = {k: [0, 0] for k in kinds}
    cases = []
    t0 = time.time()
    for it in range(args.shapes):
        kind = kinds[it % len(kinds)]
        g = T._soup(rng, kind)
        if not np.all(np.ptp(T._control_points(g), 0) > 0):
            continue
        xf = T._frame(g, 24, 24)
        pts = T._texels(xf, 24, 24, rng, extra=100)
        for mode, overlap in ((1, True), (2, False), (3, True), (4, False), (3, False), (4, True)):
            want = ref.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts)
            for policy in (T.FAST_GUARDED, T.FOLDED):
                got, _ = T._filtered(dev, mode, overlap, policy, it % 3 - 1, g, xf, 24, 24, pts)
                d = (T._texel_bits(got, xf) != T._texel_bits(want, xf)) & ~(np.isnan(got) & np.isnan(want))
                totals[kind][0] += d.size
                totals[kind][1] += int(d.sum())
                if d.any() and len(cases) < 20:
                    cases.append((it, kind, mode, overlap, policy, int(d.sum()), g.orders.tolist()))
    for k, (n, bad) in totals.items():
        print(f"{k:9s} {bad:7d} of {n} values differ")
    for c in cases:
        print("  ", c)
    print(f"{time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
