"""Host-side scanline interface of the C++ mirror (msdfgen_b200/host/scanline_host.cpp): Scanline, EdgeSegment::
scanlineIntersections, Shape::scanline / orientContours, scanlineSDF, estimateSDFError and the equation solvers.  One user
program (tests/dropin/host_caller.cpp) is built against the reference and against the mirror; the mirror's results must equal
the committed reference results (tests/golden/hostscan_v1.npz) exactly.  Host code only: no GPU needed."""
import os

import numpy as np
import pytest

import hostscan

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hostscan_v1.npz"))
FIELDS = {n: GOLDEN[f"field{n}"] for n in (1, 3, 4)}


def _check(res):
    bad = []
    for k, v in res.items():
        want = GOLDEN[k]
        same = np.array_equal(v.view(np.uint64), want.view(np.uint64)) if v.dtype == np.float64 else np.array_equal(v, want)
        if not same:
            bad.append((k, int(np.sum(v != want))))
    assert not bad, bad[:10]


def test_mirror_matches_reference_results():
    _check(hostscan.run(hostscan.build("b200"), FIELDS))


def test_golden_is_current_reference():
    if not os.path.isdir(hostscan.REF):
        pytest.skip("/root/reference is not present on this box")
    _check(hostscan.run(hostscan.build("ref"), FIELDS))


def test_orientation_makes_nonzero_fill_equal_even_odd_fill():
    """After orientContours() the shape filled by the non-zero rule is the shape its even-odd rule describes
    (Shape.cpp:144-199) — checked on nested rings wound the wrong way round."""
    import ctypes as C
    lib = hostscan.build("b200")
    sq = lambda x0, y0, x1, y1, ccw: [(x0, y0), (x1, y0), (x1, y1), (x0, y1)] if ccw else [(x0, y0), (x0, y1), (x1, y1), (x1, y0)]
    rings = [sq(1, 1, 15, 15, True), sq(3, 3, 13, 13, True), sq(5, 5, 11, 11, False)]     # the middle one should be a hole
    pts, cb = [], [0]
    for ring in rings:
        for i in range(4):
            a, b = ring[i], ring[(i + 1) % 4]
            pts.append([a[0], a[1], b[0], b[1], 0, 0, 0, 0])
        cb.append(len(pts))
    pts, cb, od = np.array(pts, np.float64), np.array(cb, np.uint32), np.ones(len(pts), np.uint8)
    op, oo = np.zeros_like(pts), np.zeros_like(od)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    lib.hc_orient(C.c_int(3), P(cb, C.c_uint32), P(pts, C.c_double), P(od, C.c_uint8), P(op, C.c_double), P(oo, C.c_uint8))
    xs = np.arange(0.25, 16, 0.5)
    for y in (2.2, 4.1, 8.3):
        out = {}
        for label, p, rule in (("odd_before", pts, 1), ("nonzero_after", op, 0)):
            cnt, s, f = (np.zeros(len(xs), np.int32) for _ in range(3))
            lib.hc_scanline(C.c_int(3), P(cb, C.c_uint32), P(p, C.c_double), P(od, C.c_uint8), C.c_double(y), C.c_int(len(xs)), P(xs, C.c_double), C.c_int(rule),
                            P(cnt, C.c_int), P(s, C.c_int), P(f, C.c_int))
            out[label] = f.copy()
        assert np.array_equal(out["odd_before"], out["nonzero_after"])
