#!/usr/bin/env python
"""An atlas builder's loop on one B200, end to end through the public Python binding:

    flat glyph outlines -> Shape::normalize + edgeColoringInkTrap (host library, threaded, allocation-free)
                        -> generateMTSDF + scanline sign correction + error correction, one device batch,
                           8-bit texels written straight into the rectangles of a pinned atlas page
                        -> renderSDF preview of one glyph from the page

    python examples/atlas_pipeline.py [--glyphs 1024] [--size 32]

Prints one JSON line with the stage times.  Needs the built libraries (python -c "import __graft_entry__ as g; g.build()")."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--glyphs", type=int, default=1024)
    ap.add_argument("--size", type=int, default=32)
    args = ap.parse_args()
    from msdfgen_b200 import api, prep
    from msdfgen_b200.synth import synth_atlas, glyph_transform

    n, s = args.glyphs, args.size
    outlines = synth_atlas(n, seed=0xA71A5)
    t0 = time.perf_counter()
    shapes = prep.prepare(outlines, normalize=True, coloring="inktrap", angle_threshold=3.0)
    t_prep = time.perf_counter() - t0

    ctx = api.Context(0)
    cols = int(np.ceil(np.sqrt(n)))
    rows = (n + cols - 1) // cols
    pad = 2
    cell = s + 2 * pad
    page = ctx.pinned((rows * cell, cols * cell * 4), np.uint8)       # RGBA page, device-visible: kernels store into it directly
    page[...] = 0
    rects = [page[(g // cols) * cell + pad:(g // cols) * cell + pad + s, ((g % cols) * cell + pad) * 4:((g % cols) * cell + pad + s) * 4].reshape(s, s, 4)
             for g in range(n)]
    cfg = api.MSDFGeneratorConfig(scanlinePass=True, fillRule=api.FILL_NONZERO)
    xf = glyph_transform(s)
    ctx.generate_batch_u8(api.MTSDF, rects, shapes, xf, cfg)           # warm-up (allocations, first launch)
    t0 = time.perf_counter()
    ctx.generate_batch_u8(api.MTSDF, rects, shapes, xf, cfg)
    t_gen = time.perf_counter() - t0

    # preview of glyph 0 at 4x from its 8-bit texels (what a text renderer's shader does with the page)
    field = (np.asarray(rects[0], dtype=np.float32) / np.float32(255)).copy()
    preview = np.empty((4 * s, 4 * s, 1), np.float32)
    ctx.render_sdf(preview, field, sdf_px_range=4.0)
    print(json.dumps({"glyphs": n, "size": s, "page": list(page.shape), "prepare_ms": round(t_prep * 1e3, 2), "generate_u8_ms": round(t_gen * 1e3, 2),
                      "px_per_s": round(n * s * s / t_gen), "preview_coverage": float(preview.mean())}))


if __name__ == "__main__":
    main()
