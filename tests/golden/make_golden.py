"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run here (the build container), never on the GPU box:   python tests/golden/make_golden.py
The fixture holds, per case, the prepared flat shape (Shape::normalize + edgeColoringSimple by the reference's own code,
or the colours given in the shape description), the transformation, the config, and the reference's output floats and
error-correction stencil.  tests/ check the C restatement (CPU) and the CUDA path (GPU) against it bit-for-bit / within 1e-4.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.bindings import Oracle, Config  # noqa: E402
from msdfgen_b200.synth import synth_glyph, glyph_transform  # noqa: E402

# /root/reference/build-release.bat:46 ('A', 2 contours) + a crossbar contour overlapping both stems (3-contour config-1 shape)
SHAPE_A = "{ 1471,0; 1149,0; 1021,333; 435,333; 314,0; 0,0; 571,1466; 884,1466; # }{ 926,580; 724,1124; 526,580; # }"
SHAPE_A3 = SHAPE_A + "{ 380,300; 380,470; 1080,470; 1080,300; # }"
# /root/reference/README.md:214,218
SHAPE_SQUARE = "{ -1, -1; m; -1, +1; y; +1, +1; m; +1, -1; y; # }"
SHAPE_TEARDROP = "{ 0, 1; (+1.6, -0.8; -1.6, -0.8); # }"
SHAPE_YDOWN = "@y-down { 100,100; 900,120; (950,500); 880,900; 130,870; (60,480); # }{ 400,400; 420,600; 610,590; 600,380; # }"


def frame(w, h, bounds, px_range):
    """CLI-style autoframe (main.cpp:1153-1183): fit bounds into w x h with px_range pixels of distance range."""
    l, b, r, t = bounds
    l -= 0.02 * (r - l); r += 0.02 * (r - l); b -= 0.02 * (t - b); t += 0.02 * (t - b)
    s = min(w / (r - l), h / (t - b))
    tx = -l + 0.5 * (w / s - (r - l))
    ty = -b + 0.5 * (h / s - (t - b))
    rng = px_range / s
    return np.array([s, s, tx, ty, -0.5 * rng, 0.5 * rng])


def main():
    ref = Oracle("reference")
    cases, arrays = [], {}

    def add(name, shape, xf, w, h, mode, cfg=Config(), shape_ydown=False, out_ydown=False):
        cb, pts, od, cl = shape
        out, st = ref.generate(mode, cb, pts, od, cl, xf, w, h, cfg, shape_ydown=shape_ydown, out_ydown=out_ydown, want_stencil=True)
        i = len(cases)
        cases.append(dict(name=name, mode=mode, w=w, h=h, overlap=bool(cfg.overlap), ec_mode=cfg.ec_mode, ec_dcm=cfg.ec_dcm,
                          min_dev=cfg.min_deviation_ratio, min_imp=cfg.min_improve_ratio, shape_ydown=bool(shape_ydown), out_ydown=bool(out_ydown)))
        arrays[f"c{i}_cbegin"] = cb
        arrays[f"c{i}_points"] = pts
        arrays[f"c{i}_orders"] = od
        arrays[f"c{i}_colors"] = cl
        arrays[f"c{i}_xf"] = np.asarray(xf, dtype=np.float64)
        arrays[f"c{i}_out"] = out
        arrays[f"c{i}_stencil"] = st

    def described(text, coloring=True):
        cb, pts, od, cl, yd, cs = ref.read_shape_description(text)
        cb, pts, od, cl = ref.prepare(cb, pts, od, cl, normalize=True, coloring=0 if cs or not coloring else 1, ydown=yd)
        return (cb, pts, od, cl), yd

    def bounds(shape):
        pts, od = shape[1], shape[2]
        xs = np.concatenate([pts[i, 0:2 * (od[i] + 1):2] for i in range(len(od))])
        ys = np.concatenate([pts[i, 1:2 * (od[i] + 1):2] for i in range(len(od))])
        return xs.min(), ys.min(), xs.max(), ys.max()

    modes = ["sdf", "psdf", "msdf", "mtsdf"]
    # config 1: 'A' — every mode, both combiners, default error correction
    for text, nm in [(SHAPE_A, "A2"), (SHAPE_A3, "A3")]:
        shp, yd = described(text)
        for m in modes:
            for ov in (True, False):
                add(f"{nm}_{m}_ov{int(ov)}", shp, frame(32, 32, bounds(shp), 4.0), 32, 32, m, Config(overlap=ov), shape_ydown=yd)
    # README shapes
    for text, nm in [(SHAPE_SQUARE, "square"), (SHAPE_TEARDROP, "teardrop")]:
        shp, yd = described(text)
        for m in ("msdf", "mtsdf", "sdf"):
            add(f"{nm}_{m}", shp, frame(24, 20, bounds(shp), 3.0), 24, 20, m, shape_ydown=yd)
    # y-down shape into y-up and y-down bitmaps
    shp, yd = described(SHAPE_YDOWN)
    for oyd in (False, True):
        for m in ("msdf", "psdf"):
            add(f"ydown_{m}_out{int(oyd)}", shp, frame(28, 28, bounds(shp), 4.0), 28, 28, m, shape_ydown=yd, out_ydown=oyd)
    # synthetic glyphs (BASELINE configs 2/3 in small), incl. cubics, all error-correction mode combinations
    for k in range(10):
        g = synth_glyph(0xB200 + k, cubic_fraction=0.25 if k % 2 else 0.0)
        shp = ref.prepare(g.cbegin, g.points, g.orders)
        add(f"synth{k}_msdf32", shp, glyph_transform(32), 32, 32, "msdf")
        if k < 4:
            add(f"synth{k}_mtsdf24", shp, glyph_transform(24), 24, 24, "mtsdf")
            add(f"synth{k}_sdf16", shp, glyph_transform(16), 16, 16, "sdf")
            add(f"synth{k}_psdf16_simple", shp, glyph_transform(16), 16, 16, "psdf", Config(overlap=False))
    g = synth_glyph(0xB200 + 3)
    shp = ref.prepare(g.cbegin, g.points, g.orders)
    for ecm in (0, 1, 2, 3):
        for dcm in (0, 1, 2):
            add(f"ec{ecm}{dcm}_msdf", shp, glyph_transform(24), 24, 24, "msdf", Config(ec_mode=ecm, ec_dcm=dcm))
    add("ec_ratios_mtsdf", shp, glyph_transform(24), 24, 24, "mtsdf", Config(min_deviation_ratio=1.5, min_improve_ratio=1.01, ec_dcm=2))
    # anisotropic projection, asymmetric range, translation
    add("aniso_msdf", shp, np.array([40.0, 24.0, 0.1, -0.05, -0.02, 0.11]), 36, 22, "msdf")
    # edge cases: empty shape, an empty contour between real ones, un-normalised 1- and 2-edge contours, degenerate zero-length edge
    empty = (np.zeros(1, np.uint32), np.zeros((0, 8)), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    add("empty_msdf", empty, glyph_transform(8), 8, 8, "msdf")
    add("empty_sdf", empty, glyph_transform(8), 8, 8, "sdf")
    cb, pts, od, cl = shp
    withempty = (np.concatenate([cb[:1], cb[:1], cb[1:2], cb[1:2], cb[2:]]).astype(np.uint32), pts, od, cl)
    add("emptycontour_msdf", withempty, glyph_transform(20), 20, 20, "msdf")
    one = np.zeros((1, 8)); one[0] = (0.2, 0.2, 1.3, 0.1, -0.3, 0.9, 0.2, 0.2)
    add("oneedge_cubic_mtsdf", (np.array([0, 1], np.uint32), one, np.array([3], np.uint8), np.array([7], np.uint8)), glyph_transform(16), 16, 16, "mtsdf")
    two = np.zeros((2, 8)); two[0, :6] = (0.2, 0.3, 0.5, 0.9, 0.8, 0.3); two[1, :4] = (0.8, 0.3, 0.2, 0.3)
    add("twoedge_msdf", (np.array([0, 2], np.uint32), two, np.array([2, 1], np.uint8), np.array([6, 5], np.uint8)), glyph_transform(16), 16, 16, "msdf")
    deg = np.zeros((4, 8)); deg[0, :4] = (0.2, 0.2, 0.8, 0.2); deg[1, :4] = (0.8, 0.2, 0.8, 0.2); deg[2, :4] = (0.8, 0.2, 0.5, 0.8); deg[3, :4] = (0.5, 0.8, 0.2, 0.2)
    add("zerolen_msdf", (np.array([0, 4], np.uint32), deg, np.ones(4, np.uint8), np.array([6, 5, 3, 6], np.uint8)), glyph_transform(16), 16, 16, "msdf")

    arrays["index"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **arrays)
    print(f"{len(cases)} cases -> {path} ({os.path.getsize(path)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
