"""Flat (struct-of-arrays) shape batches — the data that crosses the C ABI (include/msdfgen_b200.h: b200msdf_shapes).

A batch of glyphs is the reference's ``Shape -> Contour -> EdgeHolder -> EdgeSegment`` object tree
(/root/reference/core/Shape.h:15-58, Contour.h:10-32, edge-segments.h:15-144) flattened into five arrays:

    gbegin  uint32[nGlyphs+1]    contour index range of every glyph (Shape::contours)
    cbegin  uint32[nContours+1]  segment index range of every contour (Contour::edges, in edges order)
    points  float64[nSeg, 8]     control points p0..p3 as x,y pairs (EdgeSegment::controlPoints); unused slots are zero
    orders  uint8[nSeg]          EdgeSegment::type(): 1 linear, 2 quadratic, 3 cubic
    colors  uint8[nSeg]          EdgeColor bitmask (EdgeColor.h:9-18)
    ydown   uint8[nGlyphs]       Shape::inverseYAxis
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

BLACK, RED, GREEN, YELLOW, BLUE, MAGENTA, CYAN, WHITE = range(8)


@dataclass
class FlatShapes:
    gbegin: np.ndarray
    cbegin: np.ndarray
    points: np.ndarray
    orders: np.ndarray
    colors: np.ndarray
    ydown: np.ndarray

    def __post_init__(self):
        self.gbegin = np.ascontiguousarray(self.gbegin, dtype=np.uint32)
        self.cbegin = np.ascontiguousarray(self.cbegin, dtype=np.uint32)
        self.points = np.ascontiguousarray(self.points, dtype=np.float64).reshape(-1, 8)
        self.orders = np.ascontiguousarray(self.orders, dtype=np.uint8)
        self.colors = np.ascontiguousarray(self.colors, dtype=np.uint8)
        self.ydown = np.ascontiguousarray(self.ydown, dtype=np.uint8)
        assert self.gbegin[-1] == len(self.cbegin) - 1 and self.cbegin[-1] == len(self.orders) == len(self.points) == len(self.colors)
        assert len(self.ydown) == self.n_glyphs

    @property
    def n_glyphs(self):
        return len(self.gbegin) - 1

    @property
    def n_contours(self):
        return len(self.cbegin) - 1

    @property
    def n_segments(self):
        return len(self.orders)

    @staticmethod
    def single(cbegin, points, orders, colors=None, ydown=False):
        orders = np.asarray(orders, dtype=np.uint8)
        colors = np.full(len(orders), WHITE, dtype=np.uint8) if colors is None else colors
        return FlatShapes(np.array([0, len(cbegin) - 1], dtype=np.uint32), cbegin, points, orders, colors, np.array([int(ydown)], dtype=np.uint8))

    def glyph(self, g):
        """Arrays of one glyph as (cbegin, points, orders, colors, ydown) with local offsets."""
        c0, c1 = int(self.gbegin[g]), int(self.gbegin[g + 1])
        s0, s1 = int(self.cbegin[c0]), int(self.cbegin[c1])
        return (self.cbegin[c0:c1 + 1] - np.uint32(s0), self.points[s0:s1], self.orders[s0:s1], self.colors[s0:s1], bool(self.ydown[g]))

    def slice(self, g0, g1):
        """Contiguous glyph range [g0, g1) as an independent batch (used to shard an atlas across GPUs)."""
        c0, c1 = int(self.gbegin[g0]), int(self.gbegin[g1])
        s0, s1 = int(self.cbegin[c0]), int(self.cbegin[c1])
        return FlatShapes(self.gbegin[g0:g1 + 1] - np.uint32(c0), self.cbegin[c0:c1 + 1] - np.uint32(s0), self.points[s0:s1],
                          self.orders[s0:s1], self.colors[s0:s1], self.ydown[g0:g1])

    @staticmethod
    def concat(parts):
        gb, cb, pts, od, cl, yd = [np.zeros(1, np.uint32)], [np.zeros(1, np.uint32)], [], [], [], []
        nc = ns = 0
        for p in parts:
            gb.append(p.gbegin[1:] + np.uint32(nc))
            cb.append(p.cbegin[1:] + np.uint32(ns))
            pts.append(p.points); od.append(p.orders); cl.append(p.colors); yd.append(p.ydown)
            nc += p.n_contours
            ns += p.n_segments
        return FlatShapes(np.concatenate(gb), np.concatenate(cb), np.concatenate(pts) if pts else np.zeros((0, 8)),
                          np.concatenate(od) if od else np.zeros(0, np.uint8), np.concatenate(cl) if cl else np.zeros(0, np.uint8),
                          np.concatenate(yd) if yd else np.zeros(0, np.uint8))
