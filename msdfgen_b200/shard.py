"""Sharding of a glyph atlas across the GPUs of one box (one process per GPU).

Glyph batches are embarrassingly parallel (every glyph reads only its own segments and writes only its own tile —
SURVEY.md §8e), so ranks take contiguous glyph ranges balanced by cost = sum(segments x texels) and exchange nothing on
the data path; torch.distributed is used only for start-up barriers and for reducing timings / gathering results.
"""
from __future__ import annotations

import numpy as np

from .flat import FlatShapes


def glyph_costs(shapes: FlatShapes, dims):
    """Per-glyph work estimate: segments x texels (the distance kernel walks every segment for every texel)."""
    d = np.asarray(dims, dtype=np.int64).reshape(-1, 2)
    texels = (d[:, 0] * d[:, 1]) if len(d) > 1 else np.full(shapes.n_glyphs, d[0, 0] * d[0, 1], np.int64)
    segs = shapes.cbegin[shapes.gbegin[1:]].astype(np.int64) - shapes.cbegin[shapes.gbegin[:-1]].astype(np.int64)
    return np.maximum(segs, 1) * np.maximum(texels, 1)


def partition(costs, world):
    """Contiguous ranges [b[r], b[r+1]) with near-equal total cost; every glyph belongs to exactly one rank."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if world <= 1 or n == 0:
        return np.array([0] + [n] * max(world, 1), dtype=np.int64)
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    b = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def shard(shapes: FlatShapes, dims, rank, world):
    """The rank's slice of the atlas plus its global glyph range."""
    b = partition(glyph_costs(shapes, dims), world)
    g0, g1 = int(b[rank]), int(b[rank + 1])
    return shapes.slice(g0, g1), (g0, g1)
