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


_NAMES = ("gbegin", "cbegin", "points", "orders", "colors", "ydown")
_DTYPES = (np.uint32, np.uint32, np.float64, np.uint8, np.uint8, np.uint8)


def pack_shape(shape: FlatShapes) -> np.ndarray:
    """The flat shape arrays as ONE byte buffer: int64[8] header (element counts of the six arrays), then the arrays back to
    back, each padded to 16 bytes.  This is what crosses NVLink — a single broadcast of the shared segment pool."""
    arrays = [np.ascontiguousarray(getattr(shape, n), dtype=dt).reshape(-1) for n, dt in zip(_NAMES, _DTYPES)]
    header = np.zeros(8, np.int64)
    header[:6] = [a.size for a in arrays]
    chunks = [header.view(np.uint8)]
    for a in arrays:
        raw = a.view(np.uint8)
        chunks += [raw, np.zeros((-raw.size) % 16, np.uint8)]
    return np.concatenate(chunks)


def packed_size(header) -> int:
    return 64 + sum(int(-(-int(n) * np.dtype(dt).itemsize // 16) * 16) for n, dt in zip(header[:6], _DTYPES))


def unpack_shape(buf: np.ndarray) -> FlatShapes:
    header = buf[:64].view(np.int64)
    out, off = {}, 64
    for n, dt, count in zip(_NAMES, _DTYPES, header[:6]):
        nbytes = int(count) * np.dtype(dt).itemsize
        a = buf[off:off + nbytes].view(dt)
        out[n] = a.reshape(-1, 8) if n == "points" else a
        off += -(-nbytes // 16) * 16
    return FlatShapes(**out)


def broadcast_shape(shape: FlatShapes | None, src=0, device=None):
    """Rank `src` sends the packed shape to every rank: an 8-word size header, then ONE broadcast of the byte buffer (NCCL over
    NVLink when `device` is a CUDA device, gloo otherwise).  Returns (FlatShapes, payload bytes)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = device if device is not None else "cpu"
    if rank == src:
        packed = pack_shape(shape)
        header = torch.from_numpy(packed[:64].view(np.int64).copy()).to(dev)
    else:
        header = torch.zeros(8, dtype=torch.int64, device=dev)
    dist.broadcast(header, src=src)
    size = packed_size(header.cpu().numpy())
    if rank == src:
        payload = torch.from_numpy(packed).to(dev)
    else:
        payload = torch.empty(size, dtype=torch.uint8, device=dev)
    dist.broadcast(payload, src=src)
    return unpack_shape(payload.cpu().numpy()), size


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
