"""estimateSDFError on the device (b200msdf_estimate_sdf_error) against the compiled reference's values
(tests/golden/sdferr_v1.npz, made by tests/golden/make_sdferr_golden.py): every scanline follows the reference's operations, the
host adds them in its order, so the doubles are expected to be identical."""
import os

import numpy as np
import pytest

from msdfgen_b200.flat import FlatShapes

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sdferr_v1.npz"))


def _case(i):
    spr, rule, ydown = (int(v) for v in Z[f"meta{i}"])
    shape = FlatShapes.single(Z[f"cb{i}"], Z[f"pts{i}"], Z[f"od{i}"], None, ydown=bool(ydown))
    return shape, Z[f"xf{i}"], Z[f"field{i}"], spr, rule, float(Z[f"err{i}"])


def test_golden_values_are_the_current_reference(ref):
    for i in range(0, int(Z["count"]), 5):
        shape, xf, field, spr, rule, want = _case(i)
        cb, pts, od, _, yd = shape.glyph(0)
        assert ref.estimate_sdf_error(field, cb, pts, od, xf, spr, rule, yd) == want


@pytest.mark.gpu
def test_device_estimate_matches_reference(ctx):
    worst = 0.0
    for i in range(int(Z["count"])):
        shape, xf, field, spr, rule, want = _case(i)
        got = float(ctx.estimate_sdf_error([field], shape, xf, spr, rule)[0])
        worst = max(worst, abs(got - want))
        assert got == want or abs(got - want) <= 1e-15, (i, got, want)
    print(f"[parity] estimateSDFError: {int(Z['count'])} cases, max |delta| {worst:.3g}")


@pytest.mark.gpu
def test_device_estimate_batch_and_degenerate(ctx):
    """A batch call equals the per-bitmap calls; bitmaps narrower / lower than two texels give 0 like the reference."""
    first = _case(0)
    cases = [c for c in (_case(i) for i in range(int(Z["count"]))) if c[3:5] == first[3:5] and c[2].shape[2] == first[2].shape[2]][:4]
    shapes = FlatShapes.concat([c[0] for c in cases])
    xfs = np.stack([c[1] for c in cases])
    got = ctx.estimate_sdf_error([c[2] for c in cases], shapes, xfs, first[3], first[4])
    assert [float(v) for v in got] == [c[5] for c in cases]
    shape, xf, field, spr, rule, _ = first
    assert float(ctx.estimate_sdf_error([np.ascontiguousarray(field[:1])], shape, xf, spr, rule)[0]) == 0.0
