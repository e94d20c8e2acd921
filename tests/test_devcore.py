"""The texel semantics of the CUDA path, run on the CPU: msdfgen_b200/csrc/distance_core.cuh (segment records, the three signed-
distance kernels with fast / guarded / literal root solvers, selectors, Simple and Overlapping combiners in parked and folded
form) is per-thread code, so the same source compiles with g++ behind a shim of <cuda_runtime.h> (tests/devcore).  These tests
hold it to the oracle without a GPU: identical floats after the distance mapping (the product's contract), fp64 agreement far
inside north_star's 1e-4, and bit-level fp64 agreement when every foot point takes the literal solver.  The fp32 filter in front
of it (distance_filter.cuh: k1a_filter's bounds and masks, k1b_exact's walk over what survived, the big-shape kernel's single-pass
form) has two warp votes that reduce to the identity for a warp of one lane, and compiles the same way; it is run with the
stand-in for sqrt.approx biased by twice the instruction's error in either direction, on shapes moved far from the origin, scaled
by 1e-3 .. 1e3, stretched and rotated.  The GPU tier checks the kernels themselves (tests/test_gpu_parity.py, test_fuzz.py); this
tier guards the arithmetic they are built from.  The error-correction pass is covered the same way: error_correction.cuh's
per-texel classification (protectCorners / protectEdges / both findErrors passes with the exact-distance check evaluated in place —
the sink k3 uses for overflowed texels) runs on the host over the reference's uncorrected fields and must reproduce the
reference's corrected floats and stencil bytes in all nine mode x distance-check combinations."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import fuzz_shape
from msdfgen_b200 import prep
from msdfgen_b200.flat import FlatShapes
from msdfgen_b200.synth import synth_glyph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "devcore", "devcore_host.cpp")
SHIM = os.path.join(ROOT, "tests", "devcore", "shim")
CSRC = os.path.join(ROOT, "msdfgen_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "devcore", "_build")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
CHANNELS = {1: 1, 2: 1, 3: 3, 4: 4}
FAST_GUARDED, LITERAL, FOLDED = 1, 2, 3
FOLDED_UNGUARDED = 4        # the big-shape kernel as it runs today: texels that hit the guard band are NOT redone with the literal solver


@pytest.fixture(scope="module")
def devcore():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libdevcore.so")
    deps = [SRC, os.path.join(SHIM, "cuda_runtime.h")] + [os.path.join(CSRC, f) for f in ("distance_core.cuh", "distance_filter.cuh", "cr_math.cuh",
                                                                                          "kernels.cuh", "error_correction.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        # -ffp-contract=off is the host's -fmad=false; explicit fma() calls stay fused on both sides
        subprocess.run([CXX, "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", f"-I{SHIM}", f"-I{CSRC}",
                        "-o", lib, SRC], check=True, capture_output=True)
    h = C.CDLL(lib)
    h.devcore_exact_root_calls.restype = C.c_ulonglong
    return h


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _distance(lib, mode, overlap, policy, g, pts):
    out = np.full((len(pts), CHANNELS[mode]), np.nan)
    hits = lib.devcore_distance(C.c_int(mode), C.c_int(int(overlap)), C.c_int(policy), C.c_int(g.n_contours), _p(g.cbegin, C.c_uint32),
                                _p(g.points, C.c_double), _p(g.orders, C.c_uint8), _p(g.colors, C.c_uint8), C.c_int(len(pts)),
                                _p(pts, C.c_double), _p(out, C.c_double))
    assert hits >= 0
    return out, hits


def _sample_points(g, rng, n_random=240, n_grid=80):
    """Uniform points around the shape plus points on a half-unit lattice (ties between edges, foot points exactly on end points,
    texel centres on edges — what integer-coordinate fuzz shapes are built to provoke)."""
    lo, hi = g.points[:, :2].min(0) - 1.5, g.points[:, :2].max(0) + 1.5
    lattice = np.round(rng.uniform(lo, hi, size=(n_grid, 2)) * 2) / 2
    return np.ascontiguousarray(np.concatenate([rng.uniform(lo, hi, size=(n_random, 2)), lattice]))


def _mapped(d):
    """DistanceMapping(Range(-2, 2)) and the conversion to float, as the kernels' epilogue does it."""
    with np.errstate(over="ignore"):
        return (0.25 * (d + 2.0)).astype(np.float32).view(np.uint32)


def _shapes():
    shapes = [prep.prepare(synth_glyph(100 + i, cubic_fraction=0.3 if i % 2 else 0.0)) for i in range(16)]
    shapes += [FlatShapes.single(*fuzz_shape(seed)[1:5]) for seed in range(108)]
    return shapes


def _check(lib, oracle, shapes, seed):
    rng = np.random.default_rng(seed)
    values = fp64_diffs_literal = guard_hits = 0
    for g in shapes:
        pts = _sample_points(g, rng)
        for mode in (1, 2, 3, 4):
            for overlap in (False, True):
                want = oracle.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts)
                for policy in (FAST_GUARDED, LITERAL) + ((FOLDED,) if overlap else ()):
                    got, hits = _distance(lib, mode, overlap, policy, g, pts)
                    # the contract of the product: the float the reference would have written
                    assert np.array_equal(_mapped(got), _mapped(want)), (mode, overlap, policy)
                    finite = np.isfinite(want)
                    assert np.array_equal(finite, np.isfinite(got))
                    assert np.abs(got[finite] - want[finite]).max(initial=0.0) <= 1e-8          # north_star: 1e-4
                    if policy == LITERAL:
                        fp64_diffs_literal += int((got.view(np.uint64) != want.view(np.uint64)).sum())
                    if policy == FAST_GUARDED:
                        guard_hits += hits
                    values += got.size
    return values, fp64_diffs_literal, guard_hits


def test_device_texel_semantics_match_the_reference_on_the_cpu(devcore, ref):
    values, fp64_diffs_literal, guard_hits = _check(devcore, ref, _shapes(), seed=1)
    # with the literal solver everywhere even the doubles agree, up to a handful of last-bit differences from hoisted constants
    assert fp64_diffs_literal <= 1e-4 * values
    # the guard band fires on lattice points of the integer-coordinate fuzz shapes, and rarely
    assert 0 < guard_hits <= 2e-3 * values
    print(f"[devcore] {values} values: floats identical; literal path: {fp64_diffs_literal} doubles differ in the last bits; "
          f"guard band hit by {guard_hits} texel evaluations")


def test_device_texel_semantics_match_the_c_port(devcore, port):
    """The same check against the plain-C restatement (always available, also where /root/reference and oracle/_ref are not)."""
    _check(devcore, port, _shapes()[:8] + _shapes()[16::9], seed=2)


def test_device_windings_match_the_host_model(devcore):
    for g in _shapes():
        out = np.zeros(g.n_contours, np.int32)
        devcore.devcore_windings(C.c_int(g.n_contours), _p(g.cbegin, C.c_uint32), _p(g.points, C.c_double), _p(g.orders, C.c_uint8), _p(out, C.c_int))
        assert np.array_equal(out, prep.windings(g))


def test_literal_solver_only_runs_where_it_is_asked_for(devcore):
    """Ordinary glyphs never reach the closed-form cubic solver on the fast path (its calls are counted); cusp shapes (fuzz family
    8: SEG_EXACT_ROOTS segments) always do."""
    rng = np.random.default_rng(3)
    devcore.devcore_exact_root_calls()
    g = prep.prepare(synth_glyph(5))
    _distance(devcore, 3, True, 0, g, _sample_points(g, rng, n_grid=0))
    assert devcore.devcore_exact_root_calls() == 0
    cusps = [s for s in range(108) if fuzz_shape(s)[0] == 8]
    assert cusps
    g = FlatShapes.single(*fuzz_shape(cusps[0])[1:5])
    _distance(devcore, 3, True, 0, g, _sample_points(g, rng))
    assert devcore.devcore_exact_root_calls() > 0


# ---- the filtered path (k1a_filter + k1b_exact, and the big-shape kernel's single-pass form) ---------------------------------------

def _filtered(lib, mode, overlap, policy, sqrt_bias, g, xf, w, h, pts):
    out = np.full((len(pts), CHANNELS[mode]), np.nan)
    survivors = C.c_double(0)
    frame = np.asarray(xf[:4], np.float64)
    hits = lib.devcore_distance_filtered(C.c_int(mode), C.c_int(int(overlap)), C.c_int(policy), C.c_int(sqrt_bias), C.c_int(g.n_contours),
                                         _p(g.cbegin, C.c_uint32), _p(g.points, C.c_double), _p(g.orders, C.c_uint8), _p(g.colors, C.c_uint8),
                                         _p(frame, C.c_double), C.c_int(w), C.c_int(h), C.c_int(len(pts)), _p(pts, C.c_double),
                                         _p(out, C.c_double), C.byref(survivors))
    assert hits >= 0
    return out, survivors.value


def _control_points(g):
    return np.concatenate([g.points[g.orders >= k, 2 * k:2 * k + 2] for k in range(4)])


def _frame(g, w, h):
    """A projection that puts the shape's control-point hull into 80 % of a w x h bitmap: (sx, sy, tx, ty)."""
    cp = _control_points(g)
    lo, hi = cp.min(0), cp.max(0)
    s = 0.8 * min(w / (hi[0] - lo[0]), h / (hi[1] - lo[1]))
    return s, s, -lo[0] + 0.5 * (w / s - (hi[0] - lo[0])), -lo[1] + 0.5 * (h / s - (hi[1] - lo[1]))


def _texels(xf, w, h, rng, extra=0):
    """Unprojected texel centres of the bitmap (what the kernels evaluate) plus random points inside it."""
    xs, ys = (np.arange(w) + 0.5) / xf[0] - xf[2], (np.arange(h) + 0.5) / xf[1] - xf[3]
    pts = np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2)
    if extra:
        pts = np.concatenate([pts, np.stack([rng.uniform(0, w, extra) / xf[0] - xf[2], rng.uniform(0, h, extra) / xf[1] - xf[3]], 1)])
    return np.ascontiguousarray(pts)


def _texel_bits(d, xf):
    """The float a generate* call with a 4-pixel range writes for raw distance d: DistanceMapping(Range(-2/s, 2/s)) as the
    reference and the kernels' epilogue evaluate it, float(scale*(d+translate)) with scale = 1/(upper-lower), translate = -lower."""
    lower, upper = -2.0 / xf[0], 2.0 / xf[0]
    with np.errstate(over="ignore", invalid="ignore"):
        return ((1.0 / (upper - lower)) * (d + -lower)).astype(np.float32).view(np.uint32)


def _affine(g, scale, angle, offset):
    pts = g.points.copy()
    c, s = np.cos(angle), np.sin(angle)
    for k in range(4):
        sel = g.orders >= k
        x, y = g.points[sel, 2 * k] * scale[0], g.points[sel, 2 * k + 1] * scale[1]
        pts[sel, 2 * k], pts[sel, 2 * k + 1] = c * x - s * y + offset[0], s * x + c * y + offset[1]
    return FlatShapes.single(g.cbegin, pts, g.orders, g.colors)


def test_filtered_path_keeps_every_texel(devcore, ref):
    """k1a_filter's masks + k1b_exact's walk over the survivors, and the big-shape kernel's single-pass form, give the reference's
    floats whichever way the approximate square root errs; the filter leaves a handful of exact evaluations per texel."""
    rng = np.random.default_rng(4)
    values = 0
    survivors = []
    for i, g in enumerate(_shapes()[:12] + _shapes()[16::3]):
        w, h = ((32, 32), (64, 32), (24, 40))[i % 3]
        xf = _frame(g, w, h)
        pts = _texels(xf, w, h, rng)
        for mode in (1, 2, 3, 4):
            for overlap in (False, True):
                want = _texel_bits(ref.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts), xf)
                for policy in (FAST_GUARDED, FOLDED):
                    for bias in ((-1, 0, 1) if mode == 3 else (i % 3 - 1,)):
                        got, per_texel = _filtered(devcore, mode, overlap, policy, bias, g, xf, w, h, pts)
                        assert np.array_equal(_texel_bits(got, xf), want), (i, mode, overlap, policy, bias)
                        values += got.size
                        if policy == FAST_GUARDED and mode == 3 and overlap and i < 12:
                            survivors.append(per_texel)
    # ordinary glyphs of 8..64 segments: a few exact edge evaluations per texel survive (all of them: 8..64)
    assert 1.0 <= np.mean(survivors) <= 8.0
    print(f"[devcore] filtered path: {values} values identical; {np.mean(survivors):.2f} exact edge evaluations per texel on glyphs")


AFFINE = [((1, 1), 0.0, (1e4, -3e3)), ((1, 1), 0.0, (1e6, 7e5)), ((1e-3, 1e-3), 0.0, (0, 0)), ((1e3, 1e3), 0.0, (5e5, 5e5)),
          ((1, 0.25), 0.0, (10, 20)), ((1, 1), 0.0, (2.0**20, 2.0**20)), ((0.37, 0.37), 0.3, (0, 0)), ((7.3, 7.3), 1.0, (123.456, -77.7)),
          ((1, 1), np.pi / 4, (1e5, 1e5))]


def test_filtered_path_far_from_the_origin_scaled_and_rotated(devcore, ref):
    """The filter's slack (2^-16 M + 2^-44 |O|) and the exact path under affine maps of the test shapes, cusps excepted (next
    test): coordinates around 1e6, scales 1e-3 .. 1e3, anisotropy, rotations that turn exact collinearities into near ones."""
    rng = np.random.default_rng(5)
    shapes = [(fuzz_shape(s)[0], FlatShapes.single(*fuzz_shape(s)[1:5])) for s in range(0, 108, 2)] + [(-1, g) for g in _shapes()[:6]]
    values = 0
    for n, (family, g0) in enumerate(shapes):
        if family == 8:
            continue
        for m, (scale, angle, offset) in enumerate(AFFINE):
            if (n + m) % 3:
                continue
            g = _affine(g0, scale, angle, offset)
            w, h = (64, 32) if n % 2 else (24, 40)
            xf = _frame(g, w, h)
            pts = _texels(xf, w, h, rng, extra=300)
            for mode, overlap in ((1, False), (3, True), (4, True), (2, False)):
                want = _texel_bits(ref.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts), xf)
                for policy in (FAST_GUARDED, FOLDED):
                    got, _ = _filtered(devcore, mode, overlap, policy, n % 3 - 1, g, xf, w, h, pts)
                    assert np.array_equal(_texel_bits(got, xf), want), (n, scale, angle, offset, mode, overlap, policy)
                    values += got.size
    print(f"[devcore] affine stress: {values} values identical")


def test_cusps_under_affine_maps(devcore, ref):
    """Quadratic cusps (control point on an end point; collinear and doubling back) stay cusps under a similarity, but only up to
    rounding: cross(ab, bc) is 1e-17 instead of 0.  prepare_record recognises them by the vanishing tangent (relative criterion)
    and sends them to the literal solver like exact ones.  What is left is the reference's own dependence on the host libm: at a
    cusp the foot-point cubic has a root at exactly 0 (or 1), the closed form returns +-1e-16 there, and which of the two decides
    the texel; glibc's pow is not correctly rounded for about one argument in a few hundred (0.5004 ulp observed), cr_pow_third
    is.  Rate of such texels on this set: 2.6e-5 (asserted below 1e-4); all other values are identical."""
    rng = np.random.default_rng(9)
    cusps = [s for s in range(108) if fuzz_shape(s)[0] == 8]
    values = different = 0
    for n, seed in enumerate(cusps):
        g0 = FlatShapes.single(*fuzz_shape(seed)[1:5])
        for m, (scale, angle, offset) in enumerate([((1, 1), 0.0, (0, 0)), ((0.1, 0.1), 0.0, (0, 0)), ((1e-3, 1e-3), 0.0, (0, 0)),
                                                    ((0.37, 0.37), 0.3, (0, 0)), ((7.3, 7.3), 1.0, (123.456, -77.7)), ((1, 1), 2.5, (123.456, -77.7))]):
            g = _affine(g0, scale, angle, offset)
            xf = _frame(g, 32, 32)
            pts = _texels(xf, 32, 32, rng, extra=200)
            for mode, overlap in ((1, False), (3, True), (4, False)) if (n + m) % 2 else ((2, True), (3, False), (4, True)):
                want = ref.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts)
                got, _ = _filtered(devcore, mode, overlap, FAST_GUARDED, 0, g, xf, 32, 32, pts)
                d = _texel_bits(got, xf) != _texel_bits(want, xf)
                if scale == (1, 1) and angle == 0.0:
                    assert not d.any(), (seed, mode, overlap)           # the exact cusps of the fuzz set: every value
                different += int(d.sum())
                values += d.size
    assert different <= 1e-4 * values, (different, values)
    print(f"[devcore] cusps under affine maps: {different} of {values} values differ (libm-dependent texels of the reference)")


# ---- random shape soup -----------------------------------------------------------------------------------------------------------

def _soup(rng, kind):
    """1-4 closed contours of 2-8 random edges (self-intersecting at will).  kind: plain | grid (integer coordinates: exact ties)
    | mixed (vertex clouds of size 1e-6, 1 and 1e3 in one shape) | repeated (an edge listed twice) | sliver (control points on or
    within 1e-12 .. 1e-3 of the chord's line, also beyond its ends: near-straight edges, doubled-back edges, zero-area contours)."""
    rows, orders, cbegin = [], [], [0]
    for _ in range(rng.integers(1, 5)):
        n = rng.integers(2, 9)
        if kind == "grid":
            v = rng.integers(-8, 9, size=(n, 2)).astype(float)
        elif kind == "mixed":
            v = rng.normal(size=(n, 2)) * rng.choice([1e-6, 1.0, 1e3], size=(n, 1))
        else:
            v = rng.normal(size=(n, 2)) * 4
        for i in range(n):
            a, b = v[i], v[(i + 1) % n]
            order = int(rng.integers(1, 4))
            if kind == "sliver":
                t = rng.uniform(-0.3, 1.3, 2)
                eps = rng.choice([0, 1e-12, 1e-7, 1e-3])
                normal = np.array([a[1] - b[1], b[0] - a[0]])
                c1, c2 = a + (b - a) * t[0] + normal * eps * rng.normal(), a + (b - a) * t[1] + normal * eps * rng.normal()
            elif kind == "grid":
                c1, c2 = rng.integers(-8, 9, 2).astype(float), rng.integers(-8, 9, 2).astype(float)
            else:
                c1, c2 = (a + b) / 2 + rng.normal(size=2) * 2, (a + b) / 2 + rng.normal(size=2) * 2
            if kind == "repeated" and i and rng.random() < 0.3:
                rows.append(rows[-1])
                orders.append(orders[-1])
                continue
            row = np.zeros(8)
            row[:2 * (order + 1)] = np.concatenate([a, b] if order == 1 else [a, c1, b] if order == 2 else [a, c1, c2, b])
            rows.append(row)
            orders.append(order)
        cbegin.append(len(orders))
    colors = np.array([(6, 5, 3)[i % 3] for i in range(len(orders))], np.uint8)
    return FlatShapes.single(np.array(cbegin, np.uint32), np.array(rows), np.array(orders, np.uint8), colors)


def _soup_run(lib, oracle, kind, seed, count):
    rng = np.random.default_rng(seed)
    values = different = 0
    for it in range(count):
        g = _soup(rng, kind)
        if not np.all(np.ptp(_control_points(g), 0) > 0):
            continue
        xf = _frame(g, 24, 24)
        pts = _texels(xf, 24, 24, rng, extra=100)
        for mode, overlap in ((1, True), (2, False), (3, True), (4, False), (3, False), (4, True)):
            want = oracle.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts)
            for policy in (FAST_GUARDED, FOLDED):
                got, _ = _filtered(lib, mode, overlap, policy, it % 3 - 1, g, xf, 24, 24, pts)
                d = (_texel_bits(got, xf) != _texel_bits(want, xf)) & ~(np.isnan(got) & np.isnan(want))
                different += int(d.sum())
                values += d.size
    return values, different


@pytest.mark.parametrize("kind", ["plain", "mixed", "repeated"])
def test_random_shape_soup(devcore, ref, kind):
    values, different = _soup_run(devcore, ref, kind, seed=7, count=120)
    assert different == 0, (different, values)


def test_random_slivers_and_lattice_soup(devcore, ref):
    """Constructions aimed at what the fast root path does not guard by itself: a NEAR-TIE BETWEEN TWO DIFFERENT EDGES one of whose
    distances comes from a polished root.  quad_root's guard band covers ties between roots of ONE segment; an edge retraced
    backwards by its neighbour (a two-edge sliver contour, or the shared edge of two abutting contours) puts whole regions of
    texels at equal distance from two edges of opposite orientation, and there the reference's winner is decided in the last
    bit of its closed-form root.  prepare_record's chord-twin rule sends such quadratics to the literal solver, which reproduces
    the reference's bits (before that rule: 980 of 4.3 M values on the sliver soup).  What remains is counted and bounded."""
    values_s, different_s = _soup_run(devcore, ref, "sliver", seed=7, count=200)
    values_g, different_g = _soup_run(devcore, ref, "grid", seed=7, count=200)
    print(f"[devcore] cross-edge near-ties: slivers {different_s} of {values_s} values, lattice soup {different_g} of {values_g}")
    assert different_s <= 2e-5 * values_s and different_g <= 2e-5 * values_g


def _cubic_soup(rng, kind):
    """1-3 closed contours of 2-6 cubic edges.  kind: cusp (control points on the end points, both on one end, swapped, or doubling
    back along the chord) | collinear (all four points on one line, controls also beyond the ends) | loop (self-intersecting
    edges) | lattice (integer coordinates)."""
    rows, orders, cbegin = [], [], [0]
    for _ in range(rng.integers(1, 4)):
        n = rng.integers(2, 7)
        v = rng.integers(-8, 9, size=(n, 2)).astype(float) if kind == "lattice" else rng.normal(size=(n, 2)) * 4
        for i in range(n):
            a, b = v[i], v[(i + 1) % n]
            if kind == "cusp":
                c1, c2 = [(a, b), (a, a), (b, b), (a + (b - a) * 1.4, a - (b - a) * 0.3), (b, a)][rng.integers(0, 5)]
            elif kind == "collinear":
                t = np.sort(rng.uniform(-0.5, 1.5, 2))
                c1, c2 = a + (b - a) * t[0], a + (b - a) * t[1]
            elif kind == "loop":
                normal = np.array([a[1] - b[1], b[0] - a[0]])
                c1, c2 = b + normal * 1.5, a + normal * 1.5
            else:
                c1, c2 = rng.integers(-8, 9, 2).astype(float), rng.integers(-8, 9, 2).astype(float)
            rows.append(np.concatenate([a, c1, c2, b]))
            orders.append(3)
        cbegin.append(len(orders))
    colors = np.array([(6, 5, 3)[i % 3] for i in range(len(orders))], np.uint8)
    return FlatShapes.single(np.array(cbegin, np.uint32), np.array(rows), np.array(orders, np.uint8), colors)


@pytest.mark.parametrize("kind", ["cusp", "collinear", "loop", "lattice"])
def test_degenerate_cubics(devcore, ref, kind):
    """sd_cubic follows the reference's Newton search step for step, so cubic cusps, straight cubics, loops and lattice ties need no
    special path: every value identical."""
    rng = np.random.default_rng(5)
    for it in range(60):
        g = _cubic_soup(rng, kind)
        if not np.all(np.ptp(_control_points(g), 0) > 0):
            continue
        xf = _frame(g, 24, 24)
        pts = _texels(xf, 24, 24, rng, extra=100)
        for mode, overlap in ((1, True), (2, False), (3, True), (4, False)):
            want = ref.distance(mode, overlap, g.cbegin, g.points, g.orders, g.colors, pts)
            got, _ = _filtered(devcore, mode, overlap, FAST_GUARDED, it % 3 - 1, g, xf, 24, 24, pts)
            same = (_texel_bits(got, xf) == _texel_bits(want, xf)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), (kind, it, mode, overlap, int((~same).sum()))


# ---- error correction ------------------------------------------------------------------------------------------------------------

def _error_correction(lib, field, g, xf, overlap, mode, dcm):
    out = np.ascontiguousarray(field, np.float32).copy()
    h, w, n = out.shape
    stencil = np.zeros((h, w), np.uint8)
    x = np.asarray(xf, np.float64)
    ratio = C.c_double(1.11111111111111111)
    rc = lib.devcore_error_correction(C.c_int(n), C.c_int(int(overlap)), C.c_int(mode), C.c_int(dcm), ratio, ratio, C.c_int(g.n_contours),
                                      _p(g.cbegin, C.c_uint32), _p(g.points, C.c_double), _p(g.orders, C.c_uint8), _p(g.colors, C.c_uint8),
                                      _p(x, C.c_double), C.c_int(w), C.c_int(h), _p(out, C.c_float), _p(stencil, C.c_uint8))
    assert rc == 0
    return out, stencil


def test_error_correction_classification_on_the_cpu(devcore, ref):
    from oracle.bindings import Config, EC_DISABLED
    shapes = _shapes()[:10] + _shapes()[16::4]
    floats = 0
    for i, g in enumerate(shapes):
        w, h = (32, 32) if i % 2 else (24, 30)
        sx, sy, tx, ty = _frame(g, w, h)
        xf = (sx, sy, tx, ty, -2.0 / sx, 2.0 / sx)
        for name in ("msdf", "mtsdf"):
            for overlap in (True, False):
                field = ref.generate(name, g.cbegin, g.points, g.orders, g.colors, xf, w, h, Config(overlap=overlap, ec_mode=EC_DISABLED))
                for mode in (1, 2, 3):
                    for dcm in (0, 1, 2):
                        cfg = Config(overlap=overlap, ec_mode=mode, ec_dcm=dcm)
                        want, want_stencil = ref.error_correction(field, g.cbegin, g.points, g.orders, g.colors, xf, cfg, want_stencil=True)
                        got, stencil = _error_correction(devcore, field, g, xf, overlap, mode, dcm)
                        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (i, name, overlap, mode, dcm)
                        assert np.array_equal(stencil, want_stencil), (i, name, overlap, mode, dcm)
                        floats += want.size
    print(f"[devcore] error correction: {floats} floats and their stencils identical over 9 mode x distance-check combinations")
