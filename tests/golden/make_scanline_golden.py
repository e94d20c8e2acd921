"""Generates tests/golden/scanline_v1.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference):
distanceSignCorrection (core/rasterization.cpp:19-108) and rasterize (:9-17).

Run here (the build container), never on the GPU box:   python tests/golden/make_scanline_golden.py
Per case: the flat shape, the projection, the INPUT field (a reference-generated distance field, possibly inverted or with
texels forced onto the zero value), zero value, fill rule, and the reference's output.  kind "raster" cases hold the
rasterize output instead.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle.bindings import Oracle, Config  # noqa: E402
from msdfgen_b200.synth import synth_glyph, glyph_transform  # noqa: E402
from make_golden import SHAPE_A, SHAPE_A3, SHAPE_TEARDROP, SHAPE_YDOWN, frame  # noqa: E402

# vertices and curve end points on integer coordinates; with scale 1 and translate 0.5 the texel centres ARE the integers, so
# scanlines run through vertices, along horizontal edges and through curve extrema (every special case of
# edge-segments.cpp:289-403)
SHAPE_GRID = ("{ 2,2; 10,2; 10,6; (12,8); 10,10; (6,14); 2,10; (0,8; 0,4); # }"
              "{ 4,4; 4,8; 8,8; 8,4; # }"
              "{ 5,5; (7,5); 7,7; (5,7; 5,6); # }")
# figure eight: the two lobes wind in opposite directions (NONZERO / ODD / POSITIVE / NEGATIVE all differ somewhere)
SHAPE_EIGHT = "{ 0,0; 10,10; 0,10; 10,0; # }{ 2,4; 8,4; 8,6; 2,6; # }"


def main():
    ref = Oracle("reference")
    cases, arrays = [], {}

    def described(text):
        cb, pts, od, cl, yd, cs = ref.read_shape_description(text)
        cb, pts, od, cl = ref.prepare(cb, pts, od, cl, normalize=False, coloring=1, ydown=yd)
        return (cb, pts, od, cl), yd

    def bounds(shape):
        pts, od = shape[1], shape[2]
        xs = np.concatenate([pts[i, 0:2 * (od[i] + 1):2] for i in range(len(od))])
        ys = np.concatenate([pts[i, 1:2 * (od[i] + 1):2] for i in range(len(od))])
        return xs.min(), ys.min(), xs.max(), ys.max()

    def put(meta, shape, xf, **arr):
        i = len(cases)
        cases.append(meta)
        cb, pts, od, cl = shape
        arrays[f"c{i}_cbegin"], arrays[f"c{i}_points"], arrays[f"c{i}_orders"], arrays[f"c{i}_colors"] = cb, pts, od, cl
        arrays[f"c{i}_xf"] = np.asarray(xf, dtype=np.float64)
        for k, v in arr.items():
            arrays[f"c{i}_{k}"] = v

    def add(name, shape, xf, w, h, mode, rule, zero=0.5, shape_ydown=False, sdf_ydown=False, tweak=None):
        cb, pts, od, cl = shape
        field = ref.generate(mode, cb, pts, od, cl, xf, w, h, Config(ec_mode=0), shape_ydown=shape_ydown, out_ydown=sdf_ydown)
        if tweak is not None:
            field = tweak(field)
        out = ref.sign_correction(field, cb, pts, od, xf, zero, rule, shape_ydown=shape_ydown, sdf_ydown=sdf_ydown)
        put(dict(name=name, kind="sign", mode=mode, w=w, h=h, rule=rule, zero=zero, shape_ydown=bool(shape_ydown), sdf_ydown=bool(sdf_ydown)),
            shape, xf, field=field, out=out)

    def add_raster(name, shape, xf, w, h, rule, shape_ydown=False, out_ydown=False):
        cb, pts, od, cl = shape
        out = ref.rasterize(cb, pts, od, xf, w, h, rule, shape_ydown=shape_ydown, out_ydown=out_ydown)
        put(dict(name=name, kind="raster", mode="sdf", w=w, h=h, rule=rule, zero=0.5, shape_ydown=bool(shape_ydown), sdf_ydown=bool(out_ydown)),
            shape, xf, out=out)

    invert = lambda f: (1 - f).astype(np.float32)

    def on_zero(f):
        # force a block of texels exactly onto the zero value (ambiguous medians) and break the sign of their surroundings
        f = f.copy()
        f[6:12, 5:11, :3] = 0.5
        f[4:9, 3:8, :3] = 1 - f[4:9, 3:8, :3]
        f[7, 7, :3] = 0.5
        f[0, 0, :3] = 0.5
        f[-1, -1, :3] = 0.5
        return f

    shp, yd = described(SHAPE_A)
    for m in ("sdf", "psdf", "msdf", "mtsdf"):
        for rule in range(4):
            add(f"A_{m}_r{rule}", shp, frame(32, 32, bounds(shp), 4.0), 32, 32, m, rule, shape_ydown=yd)
    add("A_msdf_inverted", shp, frame(32, 32, bounds(shp), 4.0), 32, 32, "msdf", 0, shape_ydown=yd, tweak=invert)
    add("A_mtsdf_onzero", shp, frame(32, 32, bounds(shp), 4.0), 32, 32, "mtsdf", 0, shape_ydown=yd, tweak=on_zero)
    add("A_msdf_onzero_odd", shp, frame(32, 32, bounds(shp), 4.0), 32, 32, "msdf", 1, shape_ydown=yd, tweak=on_zero)
    shp, yd = described(SHAPE_A3)
    for rule in range(4):
        add(f"A3_mtsdf_r{rule}", shp, frame(40, 36, bounds(shp), 4.0), 40, 36, "mtsdf", rule, shape_ydown=yd)
        add_raster(f"A3_raster_r{rule}", shp, frame(40, 36, bounds(shp), 4.0), 40, 36, rule, shape_ydown=yd)
    shp, yd = described(SHAPE_EIGHT)
    for rule in range(4):
        add(f"eight_msdf_r{rule}", shp, frame(24, 24, bounds(shp), 3.0), 24, 24, "msdf", rule, shape_ydown=yd)
        add_raster(f"eight_raster_r{rule}", shp, np.array([1.0, 1.0, 0.5, 0.5]), 12, 12, rule)
    shp, yd = described(SHAPE_GRID)
    grid_xf = np.array([1.0, 1.0, 0.5, 0.5, -2.0, 2.0])
    for rule in range(4):
        add(f"grid_mtsdf_r{rule}", shp, grid_xf, 16, 16, "mtsdf", rule, shape_ydown=yd)
        add_raster(f"grid_raster_r{rule}", shp, grid_xf, 16, 16, rule)
    add("grid_sdf_half", shp, np.array([2.0, 2.0, 0.25, 0.25, -2.0, 2.0]), 30, 30, "sdf", 0)
    add("grid_msdf_mirror", shp, np.array([-1.0, 1.0, -15.5, 0.5, -2.0, 2.0]), 16, 16, "msdf", 0)      # negative x scale
    add("grid_msdf_flipy", shp, np.array([1.0, -1.0, 0.5, -15.5, -2.0, 2.0]), 16, 16, "msdf", 1)       # negative y scale
    shp, yd = described(SHAPE_TEARDROP)
    add("teardrop_sdf_zero0", shp, np.array([8.0, 8.0, 1.5, 1.2, -1.0, 1.0]), 24, 20, "sdf", 0, zero=0.25)
    add("teardrop_msdf_aniso", shp, np.array([9.0, 6.0, 1.3, 1.6, -0.4, 0.6]), 24, 20, "msdf", 0, zero=0.4)
    shp, yd = described(SHAPE_YDOWN)
    for oyd in (False, True):
        add(f"ydown_mtsdf_out{int(oyd)}", shp, frame(28, 28, bounds(shp), 4.0), 28, 28, "mtsdf", 0, shape_ydown=yd, sdf_ydown=oyd, tweak=invert)
        add_raster(f"ydown_raster_out{int(oyd)}", shp, frame(28, 28, bounds(shp), 4.0), 28, 28, 0, shape_ydown=yd, out_ydown=oyd)
    for k in range(6):
        g = synth_glyph(0x5CA0 + k, cubic_fraction=0.3 if k % 2 else 0.0)
        shape = ref.prepare(g.cbegin, g.points, g.orders)
        add(f"synth{k}_msdf", shape, glyph_transform(32), 32, 32, "msdf", k % 4, tweak=invert if k % 3 == 0 else None)
        add(f"synth{k}_sdf_w1", shape, glyph_transform(32), 1, 32, "sdf", 0)
    empty = (np.zeros(1, np.uint32), np.zeros((0, 8)), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    add("empty_msdf", empty, glyph_transform(8), 8, 8, "msdf", 0)
    add_raster("empty_raster", empty, glyph_transform(8), 8, 8, 0)

    arrays["index"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scanline_v1.npz")
    np.savez_compressed(path, **arrays)
    print(f"{len(cases)} cases -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
