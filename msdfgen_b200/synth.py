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
    for i in range(n):
        if kind[i] < linear_fraction:
            pts[i, 0:4] = (vx[i], vy[i], nx[i], ny[i])
        elif kind[i] < 1.0 - cubic_fraction:
            orders[i] = 2
            pts[i, 0:6] = (vx[i], vy[i], mx[i] - bulge[i] * dy[i], my[i] + bulge[i] * dx[i], nx[i], ny[i])
        else:
            orders[i] = 3
            pts[i] = (vx[i], vy[i],
                      vx[i] + dx[i] / 3 - bulge[i] * dy[i], vy[i] + dy[i] / 3 + bulge[i] * dx[i],
                      vx[i] + 2 * dx[i] / 3 - bulge2[i] * dy[i], vy[i] + 2 * dy[i] / 3 + bulge2[i] * dx[i],
                      nx[i], ny[i])
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
