"""renderSDF / simulate8bit (core/render-sdf.cpp): the C restatement (CPU) and the CUDA path (GPU, through the C ABI) against
the compiled reference's outputs in tests/golden/render_v1.npz — bit for bit."""
import json
import os

import numpy as np
import pytest

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_v1.npz"))
CASES = json.loads(bytes(Z["index"]).decode())


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_port_matches_render_golden(port):
    for i, c in enumerate(CASES):
        got = port.render_sdf(Z[f"field{c['sdf_n']}"], c["ow"], c["oh"], c["out_n"], (c["lower"], c["upper"]), c["threshold"])
        assert same(got, Z[f"c{i}"]), c
    for n in (1, 3, 4):
        assert same(port.simulate_8bit(Z[f"field{n}"] * 1.6 - 0.3), Z[f"snap{n}"])
    with pytest.raises(RuntimeError):
        port.render_sdf(Z["field3"], 8, 8, 4)           # the reference has no 4 <- 3 overload


def test_render_golden_is_current_reference(ref):
    for i, c in list(enumerate(CASES))[::9]:
        got = ref.render_sdf(Z[f"field{c['sdf_n']}"], c["ow"], c["oh"], c["out_n"], (c["lower"], c["upper"]), c["threshold"])
        assert same(got, Z[f"c{i}"])


@pytest.mark.gpu
def test_cuda_matches_render_golden(ctx):
    for i, c in enumerate(CASES):
        out = np.full((c["oh"], c["ow"], c["out_n"]), np.nan, np.float32)
        ctx.render_sdf(out, np.ascontiguousarray(Z[f"field{c['sdf_n']}"]), (c["lower"], c["upper"]), c["threshold"])
        assert same(out, Z[f"c{i}"]), (c, int((out.view(np.uint32) != Z[f"c{i}"].view(np.uint32)).sum()))
    for n in (1, 3, 4):
        b = np.ascontiguousarray(Z[f"field{n}"] * 1.6 - 0.3).astype(np.float32)
        ctx.simulate_8bit(b)
        assert same(b, Z[f"snap{n}"])
    # strided output / input sections and the symmetric-width spelling of the range
    field = np.ascontiguousarray(Z["field3"])
    big = np.full((40, 60, 3), 7.0, np.float32)
    view = big[5:35, 10:50]
    ctx.render_sdf(view, field, 4.0, 0.5)
    want = np.empty((30, 40, 3), np.float32)
    ctx.render_sdf(want, field, (-2.0, 2.0), 0.5)
    assert same(view, want) and (big[:5] == 7.0).all() and (big[:, :10] == 7.0).all()
    with pytest.raises(Exception):
        ctx.render_sdf(np.zeros((8, 8, 4), np.float32), field)      # 4 <- 3: no such overload
