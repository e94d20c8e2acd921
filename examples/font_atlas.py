#!/usr/bin/env python
"""A TrueType font to an 8-bit MTSDF atlas page on one B200 — msdf-atlas-gen's inner loop without FreeType and without one
Shape object:

    font file -> every outline as ONE flat batch (host library: glyf reader with FreeType's decomposition rules)
              -> Shape::normalize + edgeColoringInkTrap on the batch
              -> autoframe per glyph, generateMTSDF + scanline sign correction + error correction in one device batch,
                 8-bit texels stored straight into the rectangles of a pinned atlas page

    python examples/font_atlas.py [font.ttf] [--size 32] [--range 4]

Without a font argument the script uses the icon font inside the CUDA toolkit's HTML documentation, if the image has it.
Prints one JSON line.  Needs the built libraries (python -c "import __graft_entry__ as g; g.build()")."""
import argparse
import glob
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("font", nargs="?")
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--range", type=float, default=4.0, help="distance range in output pixels")
    args = ap.parse_args()
    path = args.font or next(iter(sorted(glob.glob("/usr/local/cuda*/compute-sanitizer/docs/_static/css/fonts/*.ttf"))), None)
    if not path:
        sys.exit("no font given and none found in the image")
    from msdfgen_b200 import api, prep
    from msdfgen_b200.font import Font
    from msdfgen_b200.flat import FlatShapes

    s = args.size
    t0 = time.perf_counter()
    font = Font(path)
    outlines, advances, _ = font.load_glyphs(np.arange(font.n_glyphs), "em")
    t_ingest = time.perf_counter() - t0
    keep = np.flatnonzero(np.diff(outlines.gbegin) > 0)                 # glyphs with an outline
    outlines = FlatShapes.concat([FlatShapes.single(*outlines.glyph(int(g))[:4]) for g in keep])
    t0 = time.perf_counter()
    shapes = prep.prepare(outlines, normalize=True, coloring="inktrap", angle_threshold=3.0)
    t_prep = time.perf_counter() - t0

    # autoframe (main.cpp's -autoframe): the glyph's bounds, grown by half the range, centred in the cell
    xfs = np.zeros((shapes.n_glyphs, 6))
    for g in range(shapes.n_glyphs):
        l, b, r, t = prep.bounds(FlatShapes.single(*shapes.glyph(g)[:4]))
        scale = (s - args.range) / max(r - l, t - b)
        xfs[g] = (scale, scale, -l + 0.5 * (s / scale - (r - l)), -b + 0.5 * (s / scale - (t - b)), -0.5 * args.range / scale, 0.5 * args.range / scale)

    ctx = api.Context(0)
    n = shapes.n_glyphs
    cols = int(np.ceil(np.sqrt(n)))
    rows = (n + cols - 1) // cols
    page = ctx.pinned((rows * s, cols * s * 4), np.uint8)
    page[...] = 0
    rects = [page[(g // cols) * s:(g // cols + 1) * s, (g % cols) * s * 4:(g % cols + 1) * s * 4].reshape(s, s, 4) for g in range(n)]
    cfg = api.MSDFGeneratorConfig(scanlinePass=True, fillRule=api.FILL_NONZERO)
    ctx.generate_batch_u8(api.MTSDF, rects, shapes, xfs, cfg)          # warm-up
    t0 = time.perf_counter()
    ctx.generate_batch_u8(api.MTSDF, rects, shapes, xfs, cfg)
    t_gen = time.perf_counter() - t0
    inside = float((page.reshape(rows * s, cols * s, 4)[..., 3] > 127).mean())
    print(json.dumps({"font": os.path.basename(path), "glyphs": int(n), "segments": int(shapes.n_segments), "largest_glyph_segments":
                      int(np.diff(shapes.cbegin[shapes.gbegin]).max()), "page": list(page.shape), "ingest_ms": round(t_ingest * 1e3, 2),
                      "prepare_ms": round(t_prep * 1e3, 2), "generate_u8_ms": round(t_gen * 1e3, 2), "px_per_s": round(n * s * s / t_gen),
                      "inside_fraction": round(inside, 4)}))


if __name__ == "__main__":
    main()
