"""Multi-process (world_size 2, gloo, CPU) test of the glyph-sharding logic used for N > 1: ranks partition one atlas by
cost without overlap, and gathering their (stand-in) per-glyph results reproduces the single-process order."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from msdfgen_b200 import shard
from msdfgen_b200.synth import synth_atlas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_glyphs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    atlas = synth_atlas(n_glyphs, seed=321)
    mine, (g0, g1) = shard.shard(atlas, [32, 32], rank, world)
    assert mine.n_glyphs == g1 - g0
    # stand-in for the per-glyph device result: something every rank can compute from its slice only
    local = [(g0 + i, int(mine.cbegin[mine.gbegin[i + 1]] - mine.cbegin[mine.gbegin[i]]), float(mine.glyph(i)[1].sum())) for i in range(mine.n_glyphs)]
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    cost = float(shard.glyph_costs(mine, [32, 32]).sum())
    costs = [None] * world
    dist.all_gather_object(costs, cost)
    if rank == 0:
        q.put((gathered, costs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_atlas_once():
    world, n = 2, 97
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, costs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = [x for part in gathered for x in part]
    assert [g for g, _, _ in flat] == list(range(n)), "every glyph exactly once, in order"
    atlas = synth_atlas(n, seed=321)
    for g, nseg, psum in flat:
        cb, pts, od, cl, _ = atlas.glyph(g)
        assert nseg == len(od) and psum == float(pts.sum())
    assert max(costs) / (sum(costs) / world) < 1.15, f"unbalanced shards: {costs}"


def test_partition_edge_cases():
    assert list(shard.partition([], 4)) == [0, 0, 0, 0, 0]
    assert list(shard.partition([5, 5, 5, 5], 1)) == [0, 4]
    b = shard.partition(np.ones(10), 4)
    assert b[0] == 0 and b[-1] == 10 and (np.diff(b) >= 0).all()
    b = shard.partition([100, 1, 1, 1], 4)
    assert b[0] == 0 and b[-1] == 4 and (np.diff(b) >= 0).all()


def test_band_partition():
    from msdfgen_b200 import tiled
    for h in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            bands = [tiled.band(h, r, world) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == h
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in bands) - min(b[1] - b[0] for b in bands) <= 1


def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msdfgen_b200 import tiled
    from msdfgen_b200.synth import synth_big_shape
    shape = synth_big_shape(5, 7) if rank == 0 else None
    got, nbytes = tiled.broadcast_shape(shape, 0, None)
    assert nbytes == tiled.pack_shape(got).size
    want = synth_big_shape(5, 7)
    ok = all(np.array_equal(getattr(got, n), getattr(want, n)) for n in ("gbegin", "cbegin", "points", "orders", "colors", "ydown"))
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if rank == 0:
        q.put(oks)
    dist.barrier()
    dist.destroy_process_group()


def test_segment_pool_broadcast_two_ranks():
    """The one collective of the tiled single-shape path: rank 0's flat shape arrays arrive intact on every rank."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    oks = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert oks == [True, True]
