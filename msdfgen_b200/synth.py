"""Deterministic synthetic glyph workloads (BASELINE.json configs).

Glyphs live in the unit em box [0,1]^2: a clockwise (Y-up) outer contour with jittered radius, one inner counter of
opposite winding, and (with probability 0.3) a third contour of the outer's winding that overlaps it, which exercises the
OverlappingContourCombiner path.  Edges are linear or quadratic with equal probability (optionally some cubics).
Shapes come out RAW (uncoloured, un-normalised): callers run Shape::normalize + edgeColoringSimple on them —
msdfgen_b200.host for the product path, the compiled reference in the oracle tests.
"""
from __future__ import annotations

import numpy as np

from .flat import FlatShapes, WHITE


def _ring(rng, cx, cy, radius, n, clockwise, cubic_fraction, linear_fraction=0.5):
    """One closed contour of n edges around (cx, cy). Returns (points[n,8], orders[n])."""
    ang = np.sort(rng.uniform(0.0, 2.0 * np.pi, n))
    # keep vertices from collapsing onto each other
    ang = ang + np.linspace(0.0, 2.0 * np.pi, n, endpoint=False) * 0.5
    ang = np.sort(np.mod(ang, 2.0 * np.pi))
    if clockwise:
        ang = ang[::-1]
    r = radius * rng.uniform(0.75, 1.0, n)
    vx, vy = cx + r * np.cos(ang), cy + r * np.sin(ang)
    nx, ny = np.roll(vx, -1), np.roll(vy, -1)
    pts = np.zeros((n, 8))
    orders = np.ones(n, dtype=np.uint8)
    kind = rng.uniform(size=n)
    dx, dy = nx - vx, ny - vy
    mx, my = 0.5 * (vx + nx), 0.5 * (vy + ny)
    bulge = rng.uniform(-0.3, 0.3, n)
    bulge2 = rng.uniform(-0.3, 0.3, n)
    # (vectorised; element for element the same IEEE expressions as a per-edge loop)
    lin = kind < linear_fraction
    quad = ~lin & (kind < 1.0 - cubic_fraction)
    cub = ~lin & ~quad
    pts[:, 0], pts[:, 1] = vx, vy
    pts[lin, 2], pts[lin, 3] = nx[lin], ny[lin]
    orders[quad] = 2
    pts[quad, 2], pts[quad, 3] = (mx - bulge * dy)[quad], (my + bulge * dx)[quad]
    pts[quad, 4], pts[quad, 5] = nx[quad], ny[quad]
    if cub.any():
        orders[cub] = 3
        pts[cub, 2], pts[cub, 3] = (vx + dx / 3 - bulge * dy)[cub], (vy + dy / 3 + bulge * dx)[cub]
        pts[cub, 4], pts[cub, 5] = (vx + 2 * dx / 3 - bulge2 * dy)[cub], (vy + 2 * dy / 3 + bulge2 * dx)[cub]
        pts[cub, 6], pts[cub, 7] = nx[cub], ny[cub]
    return pts, orders


def synth_glyph(seed, min_segments=8, max_segments=64, cubic_fraction=0.0, third_contour_prob=0.3, linear_fraction=0.5):
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(rng.integers(min_segments, max_segments + 1))
    third = n >= 11 and rng.uniform() < third_contour_prob
    if third:
        n_outer = max(3, int(round(n * 0.45)))
        n_inner = max(3, int(round(n * 0.30)))
        n_third = max(3, n - n_outer - n_inner)
    else:
        n_outer = max(3, int(round(n * 0.58)))
        n_inner = max(3, n - n_outer)
        n_third = 0
    parts = [_ring(rng, 0.5, 0.5, 0.45, n_outer, True, cubic_fraction, linear_fraction),
             _ring(rng, 0.5 + rng.uniform(-0.04, 0.04), 0.5 + rng.uniform(-0.04, 0.04), 0.22, n_inner, False, cubic_fraction, linear_fraction)]
    if third:
        a = rng.uniform(0, 2 * np.pi)
        parts.append(_ring(rng, 0.5 + 0.33 * np.cos(a), 0.5 + 0.33 * np.sin(a), 0.16, n_third, True, cubic_fraction, linear_fraction))
    cbegin = np.cumsum([0] + [len(o) for _, o in parts]).astype(np.uint32)
    points = np.concatenate([p for p, _ in parts])
    orders = np.concatenate([o for _, o in parts])
    return FlatShapes.single(cbegin, points, orders, np.full(len(orders), WHITE, np.uint8))


def synth_atlas(n_glyphs, seed=0xB200, **kw):
    """n_glyphs raw glyphs; glyph i uses the stream seed+i (so any sub-range can be regenerated independently)."""
    return FlatShapes.concat([synth_glyph(seed + i, **kw) for i in range(n_glyphs)])


def synth_big_shape(n_contours=100, segs_per_contour=100, seed=0xC4):
    """BASELINE config 4: one SVG-style shape of n_contours x segs_per_contour mixed line/quadratic edges scattered in [0,1]^2."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    for _ in range(n_contours):
        cx, cy = rng.uniform(0.08, 0.92, 2)
        parts.append(_ring(rng, cx, cy, rng.uniform(0.02, 0.07), segs_per_contour, bool(rng.uniform() < 0.7), 0.0))
    cbegin = np.cumsum([0] + [len(o) for _, o in parts]).astype(np.uint32)
    points = np.concatenate([p for p, _ in parts])
    orders = np.concatenate([o for _, o in parts])
    return FlatShapes.single(cbegin, points, orders, np.full(len(orders), WHITE, np.uint8))


def glyph_transform(w, h=None, px_range=4.0):
    """SDFTransformation(Projection(Vector2(W,H), 0), Range(px_range/W)) as (sx, sy, tx, ty, range_lower, range_upper)."""
    h = w if h is None else h
    rng = px_range / w
    return np.array([float(w), float(h), 0.0, 0.0, -0.5 * rng, 0.5 * rng], dtype=np.float64)


def glyph_segment_counts(n_glyphs, seed=0xB200, min_segments=8, max_segments=64):
    """Segment count of every glyph of synth_atlas(n_glyphs, seed) without building it (the first draw of each glyph's stream) —
    the cost estimate used to shard a large atlas before any rank has synthesised its part."""
    return np.array([int(np.random.Generator(np.random.PCG64(seed + i)).integers(min_segments, max_segments + 1)) for i in range(n_glyphs)],
                    dtype=np.int64)


def _synth_range(args):
    g0, g1, seed, kw = args
    return FlatShapes.concat([synth_glyph(seed + i, **kw) for i in range(g0, g1)])


def synth_atlas_range(g0, g1, seed=0xB200, workers=1, **kw):
    """Glyphs [g0, g1) of synth_atlas(*, seed): identical arrays, optionally built by `workers` forked processes (call before CUDA
    is initialised in this process)."""
    n = g1 - g0
    if workers <= 1 or n < 2048:
        return _synth_range((g0, g1, seed, kw))
    import multiprocessing as mp
    step = max(512, -(-n // (workers * 4)))
    jobs = [(a, min(a + step, g1), seed, kw) for a in range(g0, g1, step)]
    with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
        return FlatShapes.concat(pool.map(_synth_range, jobs))
