"""Row-band tiling of ONE large image of one shape across ranks (BASELINE.json config 4: a 10k-segment SVG-style shape,
sdf 4096x4096 over 8 GPUs).  Pixels shard trivially, but every rank needs the whole segment pool: rank 0 broadcasts the flat
shape arrays once (the only collective on this path), then each rank generates its band with full-image pixel coordinates
(`b200msdf_batch_upload_regions`), so the bands are bit-identical to the corresponding rows of the untiled image.  Only the
modes without an error-correction pass tile this way (SDF, PSDF, or MSDF/MTSDF with correction disabled)."""
from __future__ import annotations

import numpy as np

from . import api
from .flat import FlatShapes


def band(height, rank, world):
    """Rows [r0, r1) of rank `rank`: contiguous, differing by at most one row."""
    base, extra = divmod(height, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def broadcast_shape(shape: FlatShapes | None, src=0, device=None):
    """torch.distributed broadcast of the flat shape arrays from rank `src` (NCCL when `device` is a CUDA device, else gloo)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    names = ("gbegin", "cbegin", "points", "orders", "colors", "ydown")
    meta = [None]
    if rank == src:
        meta = [[(getattr(shape, n).shape, str(getattr(shape, n).dtype)) for n in names]]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for n, (shp, dt) in zip(names, meta[0]):
        if rank == src:
            a = np.ascontiguousarray(getattr(shape, n))
            t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy())
        else:
            t = torch.empty(int(np.prod(shp)) * np.dtype(dt).itemsize, dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=src)
        out[n] = t.cpu().numpy().view(np.dtype(dt)).reshape(shp)
    return FlatShapes(**out)


def generate_band(ctx: api.Context, mode, shape: FlatShapes, transformation, width, height, rank, world, config=None):
    """This rank's rows of the width x height image -> (float32[rows, width, N], (r0, r1))."""
    r0, r1 = band(height, rank, world)
    n = api.CHANNELS[mode]
    if r1 <= r0:
        return np.empty((0, width, n), np.float32), (r0, r1)
    batch = ctx.upload(mode, shape, transformation, [width, r1 - r0], origins=[0, r0])
    try:
        batch.launch(config or api.MSDFGeneratorConfig(errorCorrection=api.ErrorCorrectionConfig(mode=api.EC_DISABLED)))
        flat = batch.download()
    finally:
        batch.free()
    return flat[: (r1 - r0) * width * n].reshape(r1 - r0, width, n), (r0, r1)
