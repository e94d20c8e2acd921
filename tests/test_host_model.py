"""CPU tests of the host-side C++ mirror of the reference's data model (msdfgen_b200/host): Shape::normalize, edge
colouring, Contour::winding and bounds must reproduce the compiled reference exactly — they produce the INPUT of the
device path."""
import numpy as np
import pytest

from msdfgen_b200 import prep
from msdfgen_b200.flat import FlatShapes
from msdfgen_b200.synth import synth_glyph, synth_atlas


def _same(a: FlatShapes, cb, pts, od, cl):
    assert np.array_equal(a.cbegin, cb)
    assert np.array_equal(a.orders, od)
    assert np.array_equal(a.colors, cl)
    assert np.array_equal(a.points.view(np.uint64), np.ascontiguousarray(pts).view(np.uint64))


@pytest.mark.parametrize("coloring,code", [("simple", 1), ("inktrap", 2)])
def test_prepare_matches_reference_on_synthetic_glyphs(ref, coloring, code):
    for i in range(40):
        g = synth_glyph(31000 + i, cubic_fraction=0.3 if i % 2 else 0.0)
        mine = prep.prepare(g, coloring=coloring, seed=i % 5)
        _same(mine, *ref.prepare(g.cbegin, g.points, g.orders, coloring=code, seed=i % 5))


def _cusp_shapes():
    """Contours whose consecutive edges meet at 180-degree cusps (Shape::normalize's de-convergence), tear-drops,
    one- and two-edge contours, smooth loops."""
    shapes = []
    z = [0, 0]
    # quadratic + cubic leaving and re-entering a point along the same line (convergent edges)
    shapes.append(([0, 2], [[0, 0, 1, 0, 1, 1, *z], [1, 1, 1, 0, 0.5, 0, 0, 0]], [2, 3]))
    shapes.append(([0, 3], [[0, 0, 2, 0, 2, 1, *z], [2, 1, 2, 0.2, 1, 0.1, 1, 0], [1, 0, 0.5, 0, *z, *z][:8]], [2, 3, 1]))
    shapes.append(([0, 2], [[0, 0, 1, 1, 2, 0, *z], [2, 0, 1, 1e-9, 0, 0, *z]], [2, 2]))
    # single cubic tear-drop (/root/reference/README.md:218) and a single quadratic / linear degenerate loop
    shapes.append(([0, 1], [[0, 1, 1.6, -0.8, -1.6, -0.8, 0, 1]], [3]))
    shapes.append(([0, 1], [[0, 0, 1, 2, 0, 0, *z]], [2]))
    # two-edge lens
    shapes.append(([0, 2], [[0, 0, 1, 1, 2, 0, *z], [2, 0, 1, -1, 0, 0, *z]], [2, 2]))
    # smooth 4-arc circle-ish contour (no corners) followed by a square (4 corners)
    k = 0.5
    arcs = [[1, 0, 1, k, k, 1, 0, 1], [0, 1, -k, 1, -1, k, -1, 0], [-1, 0, -1, -k, -k, -1, 0, -1], [0, -1, k, -1, 1, -k, 1, 0]]
    sq = [[3, 0, 4, 0, *z, *z][:8], [4, 0, 4, 1, *z, *z][:8], [4, 1, 3, 1, *z, *z][:8], [3, 1, 3, 0, *z, *z][:8]]
    shapes.append(([0, 4, 8], arcs + sq, [3, 3, 3, 3, 1, 1, 1, 1]))
    out = []
    for cb, pts, od in shapes:
        pts = np.array([p + [0] * (8 - len(p)) for p in pts], dtype=np.float64)
        out.append(FlatShapes.single(np.array(cb, np.uint32), pts, np.array(od, np.uint8)))
    return out


def test_prepare_matches_reference_on_cusps_and_small_contours(ref):
    for s in _cusp_shapes():
        for coloring, code in (("simple", 1), ("inktrap", 2)):
            for seed in (0, 1, 7):
                mine = prep.prepare(s, coloring=coloring, seed=seed)
                _same(mine, *ref.prepare(s.cbegin, s.points, s.orders, coloring=code, seed=seed))


def test_windings_and_validate_match_reference(ref):
    for s in _cusp_shapes() + [synth_glyph(5), synth_glyph(6, cubic_fraction=0.5)]:
        assert np.array_equal(prep.windings(s), ref.windings(s.cbegin, s.points, s.orders))
        assert prep.validate(s) == ref.validate(s.cbegin, s.points, s.orders)
    bad = synth_glyph(7)
    bad.points[0, 0] += 0.125
    assert prep.validate(bad) == ref.validate(bad.cbegin, bad.points, bad.orders) == False  # noqa: E712


def test_batch_prepare_equals_per_glyph(ref):
    atlas = synth_atlas(24, seed=90)
    mine = prep.prepare(atlas)
    for g in range(atlas.n_glyphs):
        cb, pts, od, cl, _ = atlas.glyph(g)
        mcb, mpts, mod, mcl, _ = mine.glyph(g)
        rcb, rpts, rod, rcl = ref.prepare(cb, pts, od)
        assert np.array_equal(mcb, rcb) and np.array_equal(mod, rod) and np.array_equal(mcl, rcl)
        assert np.array_equal(mpts.view(np.uint64), rpts.view(np.uint64))


def test_bounds():
    s = synth_glyph(11, cubic_fraction=0.4)
    l, b, r, t = prep.bounds(s)
    # control-point hull contains the tight bounds, which contain all on-curve endpoints
    ends = np.concatenate([s.points[:, 0:2], np.array([s.points[i, 2 * o:2 * o + 2] for i, o in enumerate(s.orders)])])
    assert l <= ends[:, 0].min() + 1e-12 and r >= ends[:, 0].max() - 1e-12 and b <= ends[:, 1].min() + 1e-12 and t >= ends[:, 1].max() - 1e-12
    lb, bb, rb, tb = prep.bounds(s, border=0.1)
    assert lb <= l - 0.1 + 1e-12 and rb >= r + 0.1 - 1e-12


def test_flat_preparation_path_equals_shape_path():
    """libmsdfgen_b200_host prepares a batch straight on the flat arrays (host/flat_prep.h: no Shape objects; no allocation either
    for glyphs whose geometry stays as it is).  B200MSDF_HOST_FLAT=0 sends every glyph through the msdfgen-compatible Shape objects
    (Shape::normalize, edgeColoring*) instead; both entry paths must produce identical batches."""
    import os
    from util import fuzz_shape
    from msdfgen_b200.flat import FlatShapes
    from msdfgen_b200.synth import synth_atlas

    def both(shapes, **kw):
        os.environ["B200MSDF_HOST_FLAT"] = "0"
        try:
            general = prep.prepare(shapes, **kw)
        finally:
            os.environ.pop("B200MSDF_HOST_FLAT")
        return general, prep.prepare(shapes, **kw)

    def same(a, b):
        return (all(np.array_equal(getattr(a, k), getattr(b, k)) for k in ("gbegin", "cbegin", "orders", "colors"))
                and np.array_equal(a.points.view(np.uint64), b.points.view(np.uint64)))

    fuzz = FlatShapes.concat([FlatShapes.single(*fuzz_shape(seed)[1:5]) for seed in range(108)])
    for shapes in (synth_atlas(600, seed=78, cubic_fraction=0.3), synth_atlas(600, seed=79), fuzz):
        for coloring in ("simple", "inktrap", None):
            for normalize in (True, False):
                for seed in (0, 5, 12345678901):
                    a, b = both(shapes, coloring=coloring, normalize=normalize, seed=seed)
                    assert same(a, b), (coloring, normalize, seed)
