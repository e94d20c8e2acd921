#!/usr/bin/env python
"""bench.py — MSDF pixels/sec on the BASELINE.json workload (configs[1]: 4096 synthetic glyphs, 8-64 mixed line/quadratic
segments each, msdf 32x32, default MSDFGeneratorConfig i.e. error correction ON), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU implementation on the host cores

A step = one pass of the hot path (contour windings -> distance field -> error correction) over one atlas.
  value  : whole-job px/s with shapes already resident in HBM (CUDA events around every step, L2 flushed between steps,
           max over ranks).  Weak scaling: every rank owns its own 4096-glyph atlas (glyph batches shard embarrassingly,
           no data-path collective — SURVEY.md §8e).
  e2e    : the same metric through the C-ABI call a user makes (b200msdf_generate_batch) with pinned HOST buffers:
           host descriptor build + H2D of the shapes + kernels + D2H of the atlas inside the timed region.
  roofline / cpu_baseline : see DESIGN.md §Measurement.
The oracle (oracle/) is used only by the cpu_baseline leg and by --impl reference, as the thing timed on the CPU.
"""
import argparse
import atexit
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "msdf_pixels_per_sec"
UNIT = "px/s"


def workload(n_glyphs, size, rank):
    """The rank's atlas: raw synthetic glyphs (stream seed 0xB200 + global glyph index) prepared by the host library
    (Shape::normalize + edgeColoringSimple(shape, 3.0, 0)), and the README transformation Projection(W, 0), Range(4/W)."""
    from msdfgen_b200 import prep
    from msdfgen_b200.synth import synth_atlas, glyph_transform
    raw = synth_atlas(n_glyphs, seed=0xB200 + rank * n_glyphs)
    return prep.prepare(raw), glyph_transform(size)


def algorithmic_bytes(atlas, size, channels):
    """out W*H*N*4 per glyph + in 16*(order+1)+1 B per segment + 4 B per contour offset + 1 B winding +
    48 B transform per glyph."""
    out_b = atlas.n_glyphs * size * size * channels * 4
    in_b = int((16 * (atlas.orders.astype(np.int64) + 1) + 1).sum()) + 5 * atlas.n_contours + 48 * atlas.n_glyphs
    return out_b + in_b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            atexit.register(self._kill)         # never leave the sampler behind, whatever happens to the run
        except Exception:
            self.proc = None

    def _kill(self):
        if self.proc and self.proc.poll() is None:
            self.proc.kill()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        # samples inside the timed region; the region is ~0.1 s, so if the 50 ms sampling grid missed it take the samples within
        # 0.25 s around it (the GPU is under the same load during warm-up and the stage timing that follow immediately)
        inside = [t for t, _ in self.rows if t0 - 0.05 <= t <= t1 + 0.05]
        pad = 0.05 if inside else 0.25
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 - pad <= t <= t1 + pad:
                    sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            if t0 - pad <= t <= t1 + pad:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not mx:
            return None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference(atlas, xf, size, sample_glyphs, mode="msdf"):
    """Times the reference's CPU implementation of the path on the host cores over the first `sample_glyphs` glyphs.
    Returns the cpu_baseline dict. Prefers oracle/_ref (the unmodified reference), else the C restatement."""
    from oracle import bindings
    sub = atlas.slice(0, min(sample_glyphs, atlas.n_glyphs))
    px = sub.n_glyphs * size * size
    nproc = os.cpu_count() or 1
    if bindings.ref_available(kind="reference_gp") and bindings.ref_available(kind="reference_omp"):
        os.environ.setdefault("OMP_NUM_THREADS", str(nproc))
        args = (mode, sub.gbegin, sub.cbegin, sub.points, sub.orders, sub.colors, xf, size, size)
        # (a) outer OpenMP loop over glyphs around the unmodified single-threaded generateMSDF — how an atlas builder
        #     (msdf-atlas-gen) drives the reference, and the fastest way to use it on a many-core host
        gp = bindings.Oracle("reference_gp")
        warm = atlas.slice(0, min(64, atlas.n_glyphs))      # spin up the OpenMP team and per-thread malloc arenas (untimed)
        gp.generate_batch(mode, warm.gbegin, warm.cbegin, warm.points, warm.orders, warm.colors, xf, size, size, glyph_parallel=True)
        _, t_glyph = gp.generate_batch(*args, glyph_parallel=True)
        # (b) the reference's own parallelisation: OpenMP over the rows of one bitmap (core/msdfgen.cpp:56-75)
        few = atlas.slice(0, min(max(sample_glyphs // 64, 8), 32, atlas.n_glyphs))
        fargs = (mode, few.gbegin, few.cbegin, few.points, few.orders, few.colors, xf, size, size)
        _, t_rows = bindings.Oracle("reference_omp").generate_batch(*fargs, glyph_parallel=False)
        rows_rate = few.n_glyphs * size * size / t_rows
        # (c) single thread, non-OpenMP build
        _, t_single = bindings.Oracle("reference").generate_batch(*fargs)
        single_rate = few.n_glyphs * size * size / t_single
        return {"value": px / t_glyph, "unit": UNIT, "cores": nproc, "kind": "reference",
                "sample": f"first {sub.n_glyphs} glyphs of the workload; unmodified generateMSDF (default config) under an outer OpenMP loop "
                          f"over glyphs on {nproc} host threads; for comparison the reference's own row-parallel OpenMP build: "
                          f"{rows_rate:.3e} px/s, single thread: {single_rate:.3e} px/s (both on {few.n_glyphs} glyphs)",
                "row_parallel_value": rows_rate, "single_thread_value": single_rate, "seconds": t_glyph}
    port = bindings.Oracle("port")
    sub = atlas.slice(0, min(max(sample_glyphs // 8, 8), atlas.n_glyphs))
    _, t = port.generate_batch(mode, sub.gbegin, sub.cbegin, sub.points, sub.orders, sub.colors, xf, size, size)
    return {"value": sub.n_glyphs * size * size / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {sub.n_glyphs} glyphs of the workload, C restatement oracle/msdf_oracle.c, 1 thread", "seconds": t}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    atlas, xf = workload(args.glyphs, args.size, 0)
    sample = min(args.ref_sample, args.glyphs)
    for _ in range(min(args.warmup, 1)):
        cpu_reference(atlas, xf, args.size, max(sample // 8, 8), args.mode)
    t_total, res = 0.0, None
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        res = cpu_reference(atlas, xf, args.size, sample, args.mode)
        t_total += res["seconds"]
    px = min(sample, atlas.n_glyphs) * args.size * args.size
    value = px * steps / t_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.glyphs} synthetic glyphs (8-64 mixed line/quadratic segments), {args.mode} {args.size}x{args.size}, default MSDFGeneratorConfig (error correction on)",
                       "timed_sample_glyphs": min(sample, atlas.n_glyphs)},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    line["cpu_baseline"]["value"] = value
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--glyphs", type=int, default=4096)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--mode", default="msdf", choices=["sdf", "psdf", "msdf", "mtsdf"],
                    help="generator (default msdf = the BASELINE metric; mtsdf --size 64 is BASELINE configs[2])")
    ap.add_argument("--ref-sample", type=int, default=4096, help="glyphs per reference-arm step")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="glyphs for the cpu_baseline leg of the b200 arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from msdfgen_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: msdfgen_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    mode = {"sdf": api.SDF, "psdf": api.PSDF, "msdf": api.MSDF, "mtsdf": api.MTSDF}[args.mode]
    channels = api.CHANNELS[mode]
    # nvidia-smi takes a few hundred ms to start streaming: launch it before the host-side workload preparation (seconds), so
    # that it is sampling every 50 ms by the time the ~0.1 s timed region runs
    sampler = ClockSampler(local_rank) if rank == 0 else None
    atlas, xf = workload(args.glyphs, args.size, rank)
    px_rank = atlas.n_glyphs * args.size * args.size
    cfg = api.MSDFGeneratorConfig()

    ctx = api.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    batch = ctx.upload(mode, atlas, xf, [args.size, args.size])
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step_events():
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        batch.launch(cfg)
        e.record(stream)
        return s, e

    # ---- value: shapes resident in HBM ----------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_events()
    barrier()
    launches0 = ctx.launch_count
    t_wall0 = time.perf_counter()
    evs = [step_events() for _ in range(args.steps)]
    barrier()
    t_wall1 = time.perf_counter()
    launches = ctx.launch_count - launches0
    ms = [s.elapsed_time(e) for s, e in evs]
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    t_dev = max_over_ranks(sum(ms) * 1e-3)
    value = world * px_rank * args.steps / t_dev

    # ---- per-kernel timing for the roofline (same events, stages launched alone) --------------------------------------------
    def time_stage(mask, reps=10):
        ts = []
        for _ in range(reps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(stream)
            batch.launch_stages(mask, cfg)
            e.record(stream)
            ts.append((s, e))
        torch.cuda.synchronize()
        return float(np.mean([s.elapsed_time(e) for s, e in ts])) * 1e-3

    batch.launch(cfg)
    torch.cuda.synchronize()
    t_k1 = time_stage(2)
    t_k2 = time_stage(4)
    t_k0 = time_stage(1)
    alg_k1 = algorithmic_bytes(atlas, args.size, channels)                       # distance field: shapes in, texels out
    alg_k2 = 2 * atlas.n_glyphs * args.size * args.size * channels * 4 + (alg_k1 - atlas.n_glyphs * args.size * args.size * channels * 4)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dominant = "k1_distance_field" if t_k1 >= t_k2 else "k2_error_correction"
    t_dom, alg_dom = (t_k1, alg_k1) if t_k1 >= t_k2 else (t_k2, alg_k2)
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get(dominant)
    achieved = alg_dom / t_dom / 1e9
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_dom,
                "kernel_ms": {"k0_windings": 1e3 * t_k0, "k1_distance_field": 1e3 * t_k1, "k2_error_correction": 1e3 * t_k2},
                "note": "fp64-ALU/latency bound by construction (about 1e4 fp64 ops per texel, SURVEY.md §8d); see profiles/ for fp64-pipe utilisation"}

    # ---- e2e: the C-ABI call with pinned host buffers -----------------------------------------------------------------------
    from msdfgen_b200.flat import FlatShapes
    def pin(a):
        b = ctx.pinned(a.shape, a.dtype)
        b[...] = a
        return b
    host_atlas = FlatShapes(pin(atlas.gbegin), pin(atlas.cbegin), pin(atlas.points), pin(atlas.orders), pin(atlas.colors), pin(atlas.ydown))
    host_out = ctx.pinned((atlas.n_glyphs, args.size, args.size, channels), np.float32)
    ctx.set_stream(0, own=True)
    for _ in range(args.warmup):
        ctx.generate_batch(mode, host_out, host_atlas, xf, cfg)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.generate_batch(mode, host_out, host_atlas, xf, cfg)
    t_e2e = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    e2e_value = world * px_rank * args.steps / t_e2e
    h2d = int(atlas.points.nbytes + atlas.orders.nbytes + atlas.colors.nbytes + atlas.cbegin.nbytes + 120 * atlas.n_glyphs + 8 * atlas.n_glyphs)
    d2h = int(host_out.nbytes)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.glyphs} synthetic glyphs per GPU (8-64 mixed line/quadratic segments), {args.mode} {args.size}x{args.size}, "
                                   "default MSDFGeneratorConfig (overlap support + error correction EDGE_PRIORITY/CHECK_DISTANCE_AT_EDGE)",
                       "glyphs_per_gpu": args.glyphs, "segments_per_gpu": int(atlas.n_segments), "l2": "flushed between timed steps (512 MB memset)",
                       "parallelism": f"glyph sharding x{world}, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
                    "api": "b200msdf_generate_batch (pinned host buffers in/out)"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "wall_s": t_wall1 - t_wall0}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cb = cpu_reference(atlas, xf, args.size, args.cpu_sample, args.mode)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    batch.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
