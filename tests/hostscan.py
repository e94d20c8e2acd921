"""Runs tests/dropin/host_caller.cpp (msdfgen's HOST-side scanline interface) through one of its two builds and returns
everything it computes for a fixed set of inputs as a dict of numpy arrays — the same dict from the reference build and from
the mirror build is the parity statement (tests/test_host_scanline.py, tests/golden/make_hostscan_golden.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

from util import fuzz_shape
from msdfgen_b200.synth import synth_glyph, glyph_transform

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "dropin", "host_caller.cpp")
OUT = os.path.join(ROOT, "tests", "dropin", "_build")
REF = "/root/reference"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(kind):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, f"libhost_caller_{kind}.so")
    host = os.path.join(ROOT, "msdfgen_b200")
    if kind == "ref":
        deps = [SRC]
        core = sorted(os.path.join(REF, "core", f) for f in os.listdir(os.path.join(REF, "core")) if f.endswith(".cpp"))
        cmd = [CXX, "-O2", "-ffp-contract=off", "-std=c++11", "-fPIC", "-shared", "-DMSDFGEN_PUBLIC=", "-DMSDFGEN_USE_CPP11", f"-I{REF}", "-o", lib, SRC] + core
    else:
        deps = [SRC, os.path.join(host, "host", "msdfgen.h"), os.path.join(host, "libmsdfgen_b200_host.so")]
        cmd = [CXX, "-O2", "-ffp-contract=off", "-std=c++11", "-fPIC", "-shared", f"-I{host}/host", "-o", lib, SRC,
               f"-L{host}", "-lmsdfgen_b200_host", "-lmsdfgen_b200", f"-Wl,-rpath,{host}"]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(cmd, check=True, capture_output=True)
    lib = C.CDLL(lib)
    lib.hc_overlap.restype = C.c_double
    lib.hc_estimate_error.restype = C.c_double
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def shapes():
    """(name, cbegin, points, orders): the fuzz families plus synthetic glyphs with cubics."""
    out = []
    for seed in range(36):
        fam, cb, pts, od, cl, xf, w, h = fuzz_shape(seed)
        if fam in (4, 5):
            continue        # 1e6 / 1e-6 coordinates: same code, and y sampling below assumes a 16-unit box
        out.append((f"fuzz{seed}", cb, pts, od))
    for k in range(6):
        g = synth_glyph(0x51A0 + k, cubic_fraction=0.4 if k % 2 else 0.0)
        gp = g.glyph(0)
        out.append((f"glyph{k}", gp[0], gp[1] * 16.0, gp[2]))       # scaled to the same 16-unit box
    return out


def run(lib, fields):
    """fields: {N: float32 [h,w,N]} distance fields of shape `glyph0` at 24x24 (from the oracle), used by the SDF functions."""
    res = {}
    rng = np.random.default_rng(77)
    # solvers
    coef = np.concatenate([rng.normal(0, 3, (300, 4)), rng.integers(-3, 4, (300, 4)).astype(np.float64)])
    coef[::7, 0] = 0
    coef[::11, 1] = 0
    nq, nc = np.zeros(len(coef), np.int32), np.zeros(len(coef), np.int32)
    rq, rc = np.zeros((len(coef), 2)), np.zeros((len(coef), 3))
    for i, (a, b, c, d) in enumerate(coef):
        x2, x3 = np.zeros(2), np.zeros(3)
        nq[i] = lib.hc_solve_quadratic(_p(x2, C.c_double), C.c_double(a), C.c_double(b), C.c_double(c))
        nc[i] = lib.hc_solve_cubic(_p(x3, C.c_double), C.c_double(a), C.c_double(b), C.c_double(c), C.c_double(d))
        rq[i, :max(nq[i], 0)] = x2[:max(nq[i], 0)]
        rc[i, :max(nc[i], 0)] = x3[:max(nc[i], 0)]
    res["solve_nq"], res["solve_nc"], res["solve_rq"], res["solve_rc"] = nq, nc, rq, rc
    ys = np.concatenate([np.arange(0, 17, 0.5), rng.uniform(0, 16, 40)])
    xs = np.concatenate([np.arange(-1, 18, 0.5), rng.uniform(-1, 17, 30)])
    for name, cb, pts, od in shapes():
        cb = np.ascontiguousarray(cb, np.uint32); pts = np.ascontiguousarray(pts, np.float64); od = np.ascontiguousarray(od, np.uint8)
        n = len(od)
        # per-segment intersections
        cnt = np.zeros((n, len(ys)), np.int32); ix = np.zeros((n, len(ys), 3)); idy = np.zeros((n, len(ys), 3), np.int32)
        for s in range(n):
            lib.hc_intersections(C.c_int(int(od[s])), _p(pts[s], C.c_double), C.c_int(len(ys)), _p(ys, C.c_double), _p(cnt[s], C.c_int), _p(ix[s], C.c_double), _p(idy[s], C.c_int))
        res[f"{name}_cnt"], res[f"{name}_ix"], res[f"{name}_idy"] = cnt, ix, idy
        # shape scanlines
        q = np.zeros((len(ys), 4, 3, len(xs)), np.int32)
        for i, y in enumerate(ys):
            for rule in range(4):
                lib.hc_scanline(C.c_int(len(cb) - 1), _p(cb, C.c_uint32), _p(pts, C.c_double), _p(od, C.c_uint8), C.c_double(y), C.c_int(len(xs)), _p(xs, C.c_double),
                                C.c_int(rule), _p(q[i, rule, 0], C.c_int), _p(q[i, rule, 1], C.c_int), _p(q[i, rule, 2], C.c_int))
        res[f"{name}_scan"] = q
        # orientation
        op, oo = np.zeros_like(pts), np.zeros_like(od)
        lib.hc_orient(C.c_int(len(cb) - 1), _p(cb, C.c_uint32), _p(pts, C.c_double), _p(od, C.c_uint8), _p(op, C.c_double), _p(oo, C.c_uint8))
        res[f"{name}_orient_pts"], res[f"{name}_orient_orders"] = op, oo
    # Scanline::overlap on explicit lists (ties, empty lists, intersections outside the window)
    ov = []
    for k in range(200):
        na, nb = int(rng.integers(0, 7)), int(rng.integers(0, 7))
        ax = np.round(rng.uniform(-2, 12, na) * 2) / 2 if k % 2 else rng.uniform(-2, 12, na)
        bx = np.round(rng.uniform(-2, 12, nb) * 2) / 2 if k % 2 else rng.uniform(-2, 12, nb)
        ad, bd = rng.choice([-1, 1], na).astype(np.int32), rng.choice([-1, 1], nb).astype(np.int32)
        ax, bx = np.ascontiguousarray(ax), np.ascontiguousarray(bx)
        ov.append(lib.hc_overlap(C.c_int(na), _p(ax, C.c_double), _p(ad, C.c_int), C.c_int(nb), _p(bx, C.c_double), _p(bd, C.c_int),
                                 C.c_double(0.0), C.c_double(10.0), C.c_int(k % 4)))
    res["overlap"] = np.array(ov)
    # SDF scanlines and the error estimate on oracle-made fields of glyph0 (24x24, Projection(24/16, 0))
    name, cb, pts, od = [s for s in shapes() if s[0] == "glyph0"][0]
    cb = np.ascontiguousarray(cb, np.uint32); pts = np.ascontiguousarray(pts, np.float64); od = np.ascontiguousarray(od, np.uint8)
    xf = np.array([1.5, 1.5, 0.0, 0.0])
    for n, f in sorted(fields.items()):
        f = np.ascontiguousarray(f, np.float32)
        h, w = f.shape[:2]
        sums = np.zeros((len(ys), 2, len(xs)), np.int32)
        for i, y in enumerate(ys):
            for yd in (0, 1):
                lib.hc_scanline_sdf(C.c_int(n), _p(f, C.c_float), C.c_int(w), C.c_int(h), _p(xf, C.c_double), C.c_double(y), C.c_int(yd), C.c_int(len(xs)), _p(xs, C.c_double), _p(sums[i, yd], C.c_int))
        res[f"sdfscan{n}"] = sums
        est = []
        for spr in (1, 3, 7):
            for rule in range(4):
                for yd in (0, 1):
                    est.append(lib.hc_estimate_error(C.c_int(n), _p(f, C.c_float), C.c_int(w), C.c_int(h), C.c_int(len(cb) - 1), _p(cb, C.c_uint32), _p(pts, C.c_double), _p(od, C.c_uint8),
                                                     C.c_int(yd), _p(xf, C.c_double), C.c_int(spr), C.c_int(rule)))
        res[f"estimate{n}"] = np.array(est)
    return res


def fields_from_oracle(oracle):
    """Distance fields of glyph0 for run(): from the given oracle (reference when making the golden file, else the port)."""
    from oracle.bindings import Config
    name, cb, pts, od = [s for s in shapes() if s[0] == "glyph0"][0]
    cl = np.tile(np.array([3, 5, 6], np.uint8), len(od))[:len(od)]
    xf = np.array([1.5, 1.5, 0.0, 0.0, -2.0, 2.0])
    return {n: oracle.generate(m, cb, pts, od, cl, xf, 24, 24, Config(ec_mode=0)) for m, n in (("sdf", 1), ("msdf", 3), ("mtsdf", 4))}
