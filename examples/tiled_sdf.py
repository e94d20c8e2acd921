#!/usr/bin/env python
"""BASELINE.json config 4: one 10k-segment shape, sdf 4096x4096, tiled by rows across the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 examples/tiled_sdf.py [--size 4096]

Rank 0 builds the shape and broadcasts the segment pool (NCCL); every rank generates its row band; rank 0 gathers and checks
the seams against a single-GPU run of the rows around each seam.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--contours", type=int, default=100)
    ap.add_argument("--segments", type=int, default=100)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from msdfgen_b200 import api, prep, tiled
    from msdfgen_b200.synth import synth_big_shape, glyph_transform

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = prep.prepare(synth_big_shape(args.contours, args.segments)) if rank == 0 else None
    if world > 1:
        shape, _ = tiled.broadcast_shape(shape, 0, torch.device("cuda", local))
    xf = glyph_transform(args.size, px_range=8.0)
    ctx = api.Context(local)
    tiled.generate_band(ctx, api.SDF, shape, xf, args.size, args.size, rank, world)     # warm-up: module load, buffer growth
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine, (r0, r1) = tiled.generate_band(ctx, api.SDF, shape, xf, args.size, args.size, rank, world)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # seam check: the last row of every band boundary, regenerated as part of a band that spans the seam
    ok = True
    if rank > 0 and r1 > r0:
        batch = ctx.upload(api.SDF, shape, xf, [args.size, 2], origins=[0, r0 - 1])
        batch.launch(api.MSDFGeneratorConfig())
        two = batch.download()[: 2 * args.size].reshape(2, args.size, 1)
        batch.free()
        ok = bool(np.array_equal(two[1].view(np.uint32), mine[0].view(np.uint32)))
    flags = [ok]
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, ok)
    if rank == 0:
        print(json.dumps({"workload": f"1 shape, {shape.n_segments} segments in {shape.n_contours} contours, sdf {args.size}x{args.size}",
                          "n_gpus": world, "seconds": float(dt.item()), "px_per_s": args.size * args.size / float(dt.item()),
                          "seams_bit_identical": all(flags), "collective": "1 broadcast of the segment pool" if world > 1 else "none"}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
