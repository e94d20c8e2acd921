#!/usr/bin/env python
"""bench.py — MSDF pixels/sec on the BASELINE.json workloads, one process per GPU.

    python bench.py --gpus N --steps K --warmup W                  # this repo's CUDA path; headline = configs[1] (c2)
    python bench.py --config c3|c4|c5 ...                          # another BASELINE config as the line's workload
    python bench.py --impl reference [--config ...] --gpus N ...   # the reference's CPU implementation on the host cores

BASELINE.json configs:  c2 = 4096 synthetic glyphs, msdf 32x32 (the metric's workload)        c3 = 4096 glyphs, mtsdf 64x64 + EC
                        c4 = one 10k-segment shape, sdf 4096x4096 tiled by rows over the GPUs   c5 = 65536-glyph atlas, msdf 32x32, sharded
A step = one pass of the hot path (contour windings -> distance field -> error correction) over the rank's part of the workload.
  value    whole-job px/s with shapes already resident in HBM (CUDA events around every step on the launching stream, L2
           flushed between steps, max over ranks).  c2/c3 scale WEAK (every rank owns its own 4096-glyph atlas: glyph batches
           shard embarrassingly, no data-path collective — SURVEY.md §8e); c4/c5 and the `c2_strong` entry scale STRONG (one
           fixed workload split over the ranks).  c4 is the only path with a collective: ONE NCCL broadcast of the packed
           segment pool (timed separately as `broadcast_ms`), then every rank generates its row band.
  e2e      the same metric through the C-ABI call a user makes with pinned HOST buffers: descriptor build + H2D of the shapes +
           kernels + D2H of the texels inside the timed region (wall clock between barriers, max over ranks).
  configs  (default run only) short measurements of the other BASELINE configs, so that the driver's records carry them.
  roofline / cpu_baseline   see DESIGN.md §Measurement.
The oracle (oracle/) is used only by the cpu_baseline leg and by --impl reference, as the thing timed on the CPU; the reference
arm prepares its shapes with the reference's own Shape::normalize + edgeColoringSimple and loads none of this repo's libraries.
"""
import argparse
import atexit
import json
import math
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
SEED = 0xB200

CONFIGS = {
    "c2": dict(kind="atlas", glyphs=4096, size=32, mode="msdf", scaling="weak"),
    "c3": dict(kind="atlas", glyphs=4096, size=64, mode="mtsdf", scaling="weak"),
    "c4": dict(kind="tiled", size=4096, mode="sdf", contours=100, segments=100, scaling="strong"),
    "c5": dict(kind="atlas", glyphs=65536, size=32, mode="msdf", scaling="strong"),
}


def workload_name(name, cfg, scaling=None):
    """One string per BASELINE config, identical in both arms."""
    if cfg["kind"] == "tiled":
        return (f"{name}: one synthetic SVG-style shape, {cfg['contours']} contours x {cfg['segments']} mixed line/quadratic segments, "
                f"{cfg['mode']} {cfg['size']}x{cfg['size']}, default GeneratorConfig (overlap support), tiled by rows over the GPUs")
    per = " per GPU" if (scaling or cfg["scaling"]) == "weak" else " in total (sharded over the GPUs)"
    return (f"{name}: {cfg['glyphs']} synthetic glyphs{per} (8-64 mixed line/quadratic segments each), {cfg['mode']} {cfg['size']}x{cfg['size']}, "
            "default MSDFGeneratorConfig (overlap support + error correction EDGE_PRIORITY/CHECK_DISTANCE_AT_EDGE)")


def host_threads():
    """Threads this process may really use: scheduler affinity capped by the cgroup CPU quota (os.cpu_count() reports the
    machine, not the lease)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, math.ceil(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    n = min(n, max(1, math.ceil(q / int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read()))))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(n, 1)


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
        sm, mx, pw, reasons = [], [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 - pad <= t <= t1 + pad:
                    sm.append(float(f[0]))
                    pw.append(float(f[2]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            if t0 - pad <= t <= t1 + pad:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not mx:
            return None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons),
                "power_w": float(np.median(pw)) if pw else None}


# ======================================================================================================================
# CPU legs: the reference's own implementation on the host cores (oracle/_ref = the unmodified reference, compiled)
# ======================================================================================================================

def reference_prepared(raw):
    """Shape::normalize() + edgeColoringSimple(shape, 3.0, 0) by the REFERENCE's own code (oracle/_ref), glyph by glyph."""
    from oracle import bindings
    from msdfgen_b200.flat import FlatShapes
    ref = bindings.Oracle("reference")
    parts = []
    for g in range(raw.n_glyphs):
        cb, pts, od, _, yd = raw.glyph(g)
        cb2, pts2, od2, col2 = ref.prepare(cb, pts, od, None, normalize=True, coloring=1, angle=3.0, seed=0, ydown=yd)
        parts.append(FlatShapes(np.array([0, len(cb2) - 1], np.uint32), cb2, pts2, od2, col2, np.array([int(yd)], np.uint8)))
    return FlatShapes.concat(parts)


def cpu_atlas_rates(atlas, xf, size, mode, sample_glyphs, threads, small=64):
    """The reference generator over the first `sample_glyphs` glyphs of a prepared atlas:
       value                outer OpenMP loop over glyphs around the unmodified single-threaded generator (how an atlas builder
                            such as msdf-atlas-gen drives it; the fastest CPU use of the reference), `threads` threads
       row_parallel_value   the reference's own parallelisation: OpenMP over the rows of one bitmap (core/msdfgen.cpp:56-75)
       single_thread_value  the non-OpenMP build."""
    from oracle import bindings
    if not (bindings.ref_available(kind="reference_gp") and bindings.ref_available(kind="reference_omp")):
        port = bindings.Oracle("port")
        sub = atlas.slice(0, min(max(sample_glyphs // 8, 8), atlas.n_glyphs))
        _, t = port.generate_batch(mode, sub.gbegin, sub.cbegin, sub.points, sub.orders, sub.colors, xf, size, size)
        v = sub.n_glyphs * size * size / t
        return {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "threads_used": 1, "single_thread_value": v, "row_parallel_value": None,
                "sample": f"first {sub.n_glyphs} glyphs of the workload, C restatement oracle/msdf_oracle.c, 1 thread", "seconds": t, "glyphs": sub.n_glyphs}
    gp, rows, single = bindings.Oracle("reference_gp"), bindings.Oracle("reference_omp"), bindings.Oracle("reference")
    used = gp.set_threads(threads)
    rows.set_threads(threads)
    sub = atlas.slice(0, min(sample_glyphs, atlas.n_glyphs))
    args = (mode, sub.gbegin, sub.cbegin, sub.points, sub.orders, sub.colors, xf, size, size)
    warm = atlas.slice(0, min(2 * used, atlas.n_glyphs))      # spin up the OpenMP team and the per-thread malloc arenas (untimed)
    gp.generate_batch(mode, warm.gbegin, warm.cbegin, warm.points, warm.orders, warm.colors, xf, size, size, glyph_parallel=True)
    _, t_glyph = gp.generate_batch(*args, glyph_parallel=True)
    few = atlas.slice(0, min(small, atlas.n_glyphs))
    fargs = (mode, few.gbegin, few.cbegin, few.points, few.orders, few.colors, xf, size, size)
    rows.generate_batch(*fargs, glyph_parallel=False)
    _, t_rows = rows.generate_batch(*fargs, glyph_parallel=False)
    _, t_single = single.generate_batch(*fargs)
    px = sub.n_glyphs * size * size
    fpx = few.n_glyphs * size * size
    return {"value": px / t_glyph, "unit": UNIT, "cores": used, "kind": "reference", "threads_used": used,
            "single_thread_value": fpx / t_single, "row_parallel_value": fpx / t_rows,
            "sample": f"first {sub.n_glyphs} glyphs of the workload: unmodified reference generator (default config) under an outer OpenMP loop over "
                      f"glyphs, {used} host threads (omp_get_max_threads); row_parallel_value = the reference's own OpenMP build (rows of one "
                      f"bitmap, {used} threads) and single_thread_value = the non-OpenMP build, both on the first {few.n_glyphs} glyphs",
            "seconds": t_glyph, "glyphs": sub.n_glyphs}


def cpu_tiled_rates(shape, xf, size, rows, threads):
    """c4 on the CPU: a `rows`-row window from the middle of the size x size image by the reference's generateSDF — its own
    row-parallel OpenMP build on all threads, and the single-threaded build on a 16x narrower window."""
    from oracle import bindings
    omp, single = bindings.Oracle("reference_omp"), bindings.Oracle("reference")
    used = omp.set_threads(threads)
    y0 = (size - rows) // 2
    wxf = np.array(xf, dtype=np.float64)
    wxf[3] = xf[3] - y0 / xf[1]                 # window origin folded into Projection's translate
    cb, pts, od, col, _ = shape.glyph(0)
    omp.generate("sdf", cb, pts, od, col, wxf, size, min(rows, 8))        # untimed: team start-up
    t0 = time.perf_counter()
    omp.generate("sdf", cb, pts, od, col, wxf, size, rows)
    t_rows = time.perf_counter() - t0
    narrow = max(size // 16, 1)
    t0 = time.perf_counter()
    single.generate("sdf", cb, pts, od, col, wxf, narrow, min(rows, 4))
    t_single = time.perf_counter() - t0
    return {"value": size * rows / t_rows, "unit": UNIT, "cores": used, "kind": "reference", "threads_used": used,
            "single_thread_value": narrow * min(rows, 4) / t_single, "row_parallel_value": size * rows / t_rows,
            "sample": f"rows {y0}..{y0 + rows} of the {size}x{size} image ({size * rows} texels) by the reference's generateSDF, its own OpenMP build "
                      f"(rows in parallel, {used} threads); single_thread_value on a {narrow}x{min(rows, 4)} window",
            "seconds": t_rows, "glyphs": 1}


def run_reference_arm(args, rank):
    """--impl reference: rank 0 only; no repo library is loaded (synthesis is numpy, preparation is the reference's own)."""
    if rank != 0:
        return
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)        # unconditionally: torchrun exports OMP_NUM_THREADS=1
    os.environ["OMP_PROC_BIND"] = "false"
    os.environ["OMP_DYNAMIC"] = "false"
    from msdfgen_b200.synth import synth_atlas, synth_big_shape, glyph_transform
    name = args.config or "c2"
    cfg = CONFIGS[name]
    steps = max(1, args.steps)
    t_wall = time.perf_counter()
    budget = 150.0 / (steps + 1)                        # seconds of CPU work per step
    if cfg["kind"] == "tiled":
        shape = reference_prepared(synth_big_shape(cfg["contours"], cfg["segments"]))
        xf = glyph_transform(cfg["size"], px_range=8.0)
        probe = cpu_tiled_rates(shape, xf, cfg["size"], min(threads, cfg["size"]), threads)     # one row per thread
        rows = probe["value"] * min(budget, 10.0) / cfg["size"]
        rows = int(min(cfg["size"], max(1, round(rows / threads)) * threads))                   # whole rounds of the row-parallel loop
        runs = [cpu_tiled_rates(shape, xf, cfg["size"], rows, threads) for _ in range(steps)]
        px_step = cfg["size"] * rows
        sample_desc = {"timed_sample_texels": px_step}
    else:
        n_sample = min(args.ref_sample, cfg["glyphs"])
        atlas = reference_prepared(synth_atlas(n_sample, seed=SEED))
        xf = glyph_transform(cfg["size"])
        probe = cpu_atlas_rates(atlas, xf, cfg["size"], cfg["mode"], min(256, n_sample), threads, small=16)
        n_step = int(min(n_sample, max(64, probe["value"] * min(budget, 10.0) / (cfg["size"] ** 2))))
        runs = [cpu_atlas_rates(atlas, xf, cfg["size"], cfg["mode"], n_step, threads) for _ in range(steps)]
        px_step = runs[-1]["glyphs"] * cfg["size"] ** 2
        sample_desc = {"timed_sample_glyphs": runs[-1]["glyphs"]}
    t_total = sum(r["seconds"] for r in runs)
    value = px_step * steps / t_total
    last = runs[-1]
    cb = {k: last[k] for k in ("value", "unit", "cores", "kind", "sample", "threads_used", "single_thread_value", "row_parallel_value")}
    cb["value"] = value
    cb["single_thread_value"] = float(np.median([r["single_thread_value"] for r in runs]))
    if last["row_parallel_value"] is not None:
        cb["row_parallel_value"] = float(np.median([r["row_parallel_value"] for r in runs]))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 1,
            "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict({"workload": workload_name(name, cfg), "host_threads": threads,
                                                 "shape_preparation": "reference Shape::normalize + edgeColoringSimple (oracle/_ref)"}, **sample_desc),
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall}
    print(json.dumps(line))


# ======================================================================================================================
# GPU arm
# ======================================================================================================================

class Rig:
    """torch / torch.distributed plumbing shared by the workloads of one run."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        from msdfgen_b200 import api
        self.torch, self.dist, self.api = torch, dist, api
        self.rank, self.local, self.world = rank, local_rank, world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: msdfgen_b200 has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.ctx = api.Context(local_rank)
        # one process per GPU: keep this rank's host thread and the pinned buffers it allocates on the GPU's NUMA node
        self.numa_cpus = None
        self.affinity0 = os.sched_getaffinity(0)
        if os.environ.get("B200MSDF_NO_NUMA_BIND") is None:
            try:
                self.numa_cpus = self.ctx.bind_host_thread()
            except Exception:
                self.numa_cpus = None
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        """[obj of rank 0, obj of rank 1, ...] on every rank."""
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed_steps(self, fn, steps, warmup):
        """W untimed + K timed calls of fn() on the launching stream, one CUDA-event pair per step, L2 flushed before each;
        returns (device seconds of the K steps: max over ranks, wall t0, wall t1)."""
        torch = self.torch
        self.ctx.set_stream(self.stream.cuda_stream)

        def one():
            self.flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(self.stream)
            fn()
            e.record(self.stream)
            return s, e
        for _ in range(warmup):
            one()
        self.barrier()
        t0 = time.perf_counter()
        evs = [one() for _ in range(steps)]
        self.barrier()
        t1 = time.perf_counter()
        return self.max_over_ranks(sum(s.elapsed_time(e) for s, e in evs) * 1e-3), t0, t1

    def timed_calls(self, fn, steps, warmup):
        """Host-buffer (synchronous) calls: wall clock between barriers, max over ranks."""
        self.ctx.set_stream(0, own=True)
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = time.perf_counter() - t0
        self.barrier()
        return self.max_over_ranks(dt)


def atlas_part(name, cfg, scaling, rank, world, workers):
    """This rank's glyphs of the config (raw synthesis by numpy, preparation by the host library) and its global range."""
    from msdfgen_b200 import prep, shard
    from msdfgen_b200.synth import synth_atlas_range, glyph_segment_counts, glyph_transform
    n = cfg["glyphs"]
    if scaling == "weak":
        g0, g1 = rank * n, (rank + 1) * n           # every rank its own atlas: stream seed SEED + global glyph index
    else:
        costs = glyph_segment_counts(n, SEED) * cfg["size"] ** 2
        b = shard.partition(costs, world)
        g0, g1 = int(b[rank]), int(b[rank + 1])
    raw = synth_atlas_range(g0, g1, SEED, workers=workers)
    return prep.prepare(raw), glyph_transform(cfg["size"]), (g0, g1)


def run_atlas(rig, name, cfg, scaling, steps, warmup, atlas, xf, full=False):
    """value (+ e2e, + per-kernel roofline when `full`) of a glyph-atlas config on this rank's part."""
    api, ctx = rig.api, rig.ctx
    mode = {"sdf": api.SDF, "psdf": api.PSDF, "msdf": api.MSDF, "mtsdf": api.MTSDF}[cfg["mode"]]
    channels, size = api.CHANNELS[mode], cfg["size"]
    gcfg = api.MSDFGeneratorConfig()
    px_rank = atlas.n_glyphs * size * size
    px_total = rig.sum_over_ranks(float(px_rank))
    ctx.set_stream(rig.stream.cuda_stream)
    batch = ctx.upload(mode, atlas, xf, [size, size])
    launches0 = ctx.launch_count
    t_dev, w0, w1 = rig.timed_steps(lambda: batch.launch(gcfg), steps, warmup)
    launches = (ctx.launch_count - launches0) * steps // (steps + warmup)
    res = {"value": px_total * steps / t_dev, "unit": UNIT, "ms_per_step": 1e3 * t_dev / steps, "steps": steps, "warmup": warmup, "scaling": scaling,
           "workload": workload_name(name, cfg, scaling), "glyphs_this_rank": atlas.n_glyphs, "segments_this_rank": int(atlas.n_segments),
           "gpu_launches": int(launches), "_wall": (w0, w1)}

    if full:
        torch = rig.torch

        def time_stage(mask, reps=10):
            ts = []
            for _ in range(reps):
                rig.flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(rig.stream)
                batch.launch_stages(mask, gcfg)
                e.record(rig.stream)
                ts.append((s, e))
            torch.cuda.synchronize()
            return float(np.mean([s.elapsed_time(e) for s, e in ts])) * 1e-3
        batch.launch(gcfg)
        torch.cuda.synchronize()
        t_k1, t_k2, t_k0 = time_stage(2), time_stage(4), time_stage(1)
        t_k1a, t_k1b = time_stage(16), time_stage(32)
        out_b = atlas.n_glyphs * size * size * channels * 4
        alg_k1 = algorithmic_bytes(atlas, size, channels)                       # distance field: shapes in, texels out
        alg_k2 = 2 * out_b + (alg_k1 - out_b)                                   # error correction: shapes + field in, texels out
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # the distance field is two kernels (k1a_filter: fp32 bounds, k1b_exact: fp64 evaluation); the roofline line is about ONE
        # kernel, the slowest; the distance stage's algorithmic bytes (shapes in, texels out) are charged to k1b, which writes them
        dominant = "k1b_exact" if t_k1b >= t_k2 else "k2_error_correction"
        t_dom, alg_dom = (t_k1b, alg_k1) if t_k1b >= t_k2 else (t_k2, alg_k2)
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr_path):
            traffic = json.load(open(tr_path)).get(dominant)
        achieved = alg_dom / t_dom / 1e9
        res["roofline"] = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_dom,
                           "kernel_ms": {"k0_windings": 1e3 * t_k0, "k1_distance_field": 1e3 * t_k1, "k1a_filter": 1e3 * t_k1a, "k1b_exact": 1e3 * t_k1b,
                                         "k2_error_correction": 1e3 * t_k2},
                           "kernel_ms_per_rank": rig.gather([round(1e3 * t_k1, 4), round(1e3 * t_k2, 4)]) if rig.world > 1 else None,
                           "note": "fp64-ALU/latency bound by construction (thousands of fp64 ops per 13.5-byte texel, SURVEY.md §8d); "
                                   "see profiles/ for fp64-pipe utilisation and instruction counts"}
    batch.free()

    # ---- e2e: the C-ABI call with pinned host buffers -------------------------------------------------------------------
    from msdfgen_b200.flat import FlatShapes

    def pin(a):
        b = ctx.pinned(a.shape, a.dtype)
        b[...] = a
        return b
    host_atlas = FlatShapes(pin(atlas.gbegin), pin(atlas.cbegin), pin(atlas.points), pin(atlas.orders), pin(atlas.colors), pin(atlas.ydown))
    host_out = ctx.pinned((atlas.n_glyphs, size, size, channels), np.float32)
    e_steps = max(3, min(steps, 10)) if not full else steps
    t_e2e = rig.timed_calls(lambda: ctx.generate_batch(mode, host_out, host_atlas, xf, gcfg), e_steps, min(warmup, 3))
    h2d = int(atlas.points.nbytes + atlas.orders.nbytes + atlas.colors.nbytes + atlas.cbegin.nbytes + 128 * atlas.n_glyphs)
    res["e2e"] = {"value": px_total * e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(host_out.nbytes),
                  "ms_per_step": 1e3 * t_e2e / e_steps, "steps": e_steps, "api": "b200msdf_generate_batch (pinned host buffers in/out)"}
    # (release the big pinned block: the next config allocates its own)
    ctx.release_pinned()
    return res


def run_tiled(rig, name, cfg, steps, warmup):
    """c4: ONE shape, size x size sdf, rows split over the ranks.  Rank 0 builds the shape; one NCCL broadcast of the packed
    segment pool; every rank generates its band at full-image pixel coordinates (bit-identical to the untiled image)."""
    from msdfgen_b200 import prep, tiled
    from msdfgen_b200.synth import synth_big_shape, glyph_transform
    api, ctx, torch = rig.api, rig.ctx, rig.torch
    size = cfg["size"]
    xf = glyph_transform(size, px_range=8.0)
    shape = prep.prepare(synth_big_shape(cfg["contours"], cfg["segments"])) if rig.rank == 0 else None
    bcast_ms, bcast_bytes = 0.0, 0
    if rig.world > 1:
        dev = torch.device("cuda", rig.local)
        shape_w, _ = tiled.broadcast_shape(shape, 0, dev)             # untimed: NCCL channel set-up
        rig.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        shape, bcast_bytes = tiled.broadcast_shape(shape if rig.rank == 0 else None, 0, dev)
        e.record()
        torch.cuda.synchronize()
        bcast_ms = rig.max_over_ranks(s.elapsed_time(e))
    r0, r1 = tiled.band(size, rig.rank, rig.world)
    gcfg = api.MSDFGeneratorConfig()
    ctx.set_stream(rig.stream.cuda_stream)
    batch = ctx.upload(api.SDF, shape, xf, [size, r1 - r0], origins=[0, r0])
    launches0 = ctx.launch_count
    t_dev, _, _ = rig.timed_steps(lambda: batch.launch(gcfg), steps, warmup)
    launches = (ctx.launch_count - launches0) * steps // (steps + warmup)
    mine = batch.download()[: (r1 - r0) * size].reshape(r1 - r0, size)
    batch.free()
    # seam check: the first row of this band regenerated as the second row of a band that starts one row earlier
    ok = True
    if rig.rank > 0 and r1 > r0:
        seam = ctx.upload(api.SDF, shape, xf, [size, 2], origins=[0, r0 - 1])
        seam.launch(gcfg)
        two = seam.download()[: 2 * size].reshape(2, size)
        seam.free()
        ok = bool(np.array_equal(two[1].view(np.uint32), mine[0].view(np.uint32)))
    bad = rig.sum_over_ranks(0.0 if ok else 1.0)
    # e2e: the public tiled call with host buffers (shape H2D + kernels + D2H of the band)
    t_e2e = rig.timed_calls(lambda: tiled.generate_band(ctx, api.SDF, shape, xf, size, size, rig.rank, rig.world), max(2, min(steps, 5)), 1)
    e_steps = max(2, min(steps, 5))
    px = float(size) * size
    return {"value": px * steps / t_dev, "unit": UNIT, "ms_per_step": 1e3 * t_dev / steps, "steps": steps, "warmup": warmup, "scaling": "strong",
            "workload": workload_name(name, cfg), "segments": int(shape.n_segments), "contours": int(shape.n_contours), "rows_this_rank": r1 - r0,
            "collective": (f"1 NCCL broadcast of the packed segment pool ({bcast_bytes} B)" if rig.world > 1 else "none (1 GPU)"),
            "broadcast_ms": bcast_ms, "seams_bit_identical": bad == 0, "gpu_launches": int(launches),
            "e2e": {"value": px * e_steps / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / e_steps, "steps": e_steps,
                    "h2d_bytes_per_step": int(shape.points.nbytes + shape.orders.nbytes + shape.colors.nbytes + shape.cbegin.nbytes + 128),
                    "d2h_bytes_per_step": int((r1 - r0) * size * 4), "api": "tiled.generate_band (b200msdf_batch_upload_regions + launch + download)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="BASELINE config of the line (default c2 = the metric's workload)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="atlas configs: own atlas per rank (weak) or one atlas sharded (strong)")
    ap.add_argument("--no-extras", action="store_true", help="default run: skip the short c2_strong/c3/c4/c5 measurements")
    ap.add_argument("--ref-sample", type=int, default=4096, help="glyphs prepared for the reference arm (its steps time a bounded prefix)")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="glyphs for the cpu_baseline leg of the b200 arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    args.warmup = max(args.warmup, 3)
    name = args.config or "c2"
    cfg = CONFIGS[name]
    scaling = args.scaling or cfg["scaling"]
    extras = args.config is None and not args.no_extras
    workers = max(1, min(16, host_threads() // max(world, 1)))

    # ---- host-side workload synthesis BEFORE CUDA is initialised (forked workers) ---------------------------------------------
    parts = {}
    if cfg["kind"] == "atlas":
        parts["main"] = atlas_part(name, cfg, scaling, rank, world, workers)
    if extras:
        parts["c2_strong"] = atlas_part("c2", CONFIGS["c2"], "strong", rank, world, workers)
        parts["c3"] = (parts["main"][0], None, parts["main"][2])            # same glyphs, 64x64 mtsdf
        parts["c5"] = atlas_part("c5", CONFIGS["c5"], "strong", rank, world, workers)

    # nvidia-smi takes a few hundred ms to start streaming: launch it before torch / CUDA start-up (seconds), so that it is
    # sampling every 50 ms by the time the ~0.1 s timed region runs
    sampler = ClockSampler(local_rank) if rank == 0 else None
    rig = Rig(rank, local_rank, world)
    if not hasattr(rig.ctx, "release_pinned"):
        raise SystemExit("msdfgen_b200.api.Context.release_pinned missing")

    if cfg["kind"] == "atlas":
        atlas, xf, _ = parts["main"]
        res = run_atlas(rig, name, cfg, scaling, args.steps, args.warmup, atlas, xf, full=True)
    else:
        res = run_tiled(rig, name, cfg, args.steps, args.warmup)
    w0, w1 = res.pop("_wall", (time.perf_counter(), time.perf_counter()))
    clocks = sampler.stop(w0, w1) if sampler else None

    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": scaling if cfg["kind"] == "atlas" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": res["workload"], "l2": "flushed between timed steps (512 MB memset)",
                       "parallelism": (f"glyph sharding x{world}, no data-path collective" if cfg["kind"] == "atlas" else
                                       f"row bands x{world}; {res.get('collective')}")},
            "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": clocks, "wall_s": w1 - w0}
    line["config"]["host_binding"] = (f"rank thread + pinned buffers on the GPU's NUMA node ({rig.numa_cpus} CPUs, b200msdf_bind_host_thread)"
                                      if rig.numa_cpus else "none")
    for k in ("glyphs_this_rank", "segments_this_rank", "rows_this_rank", "segments", "contours", "broadcast_ms", "seams_bit_identical"):
        if k in res:
            line["config"][k] = res[k]
    if "roofline" in res:
        line["roofline"] = res["roofline"]

    if extras:
        ex = {}
        short = max(3, min(args.steps, 5))
        a, x, _ = parts["c2_strong"]
        ex["c2_strong"] = run_atlas(rig, "c2", CONFIGS["c2"], "strong", max(5, min(args.steps, 10)), 3, a, x)
        from msdfgen_b200.synth import glyph_transform
        ex["c3"] = run_atlas(rig, "c3", CONFIGS["c3"], "weak", short, 3, parts["c3"][0], glyph_transform(64))
        ex["c4"] = run_tiled(rig, "c4", CONFIGS["c4"], short, 3)
        a, x, _ = parts["c5"]
        ex["c5"] = run_atlas(rig, "c5", CONFIGS["c5"], "strong", short, 3, a, x)
        for v in ex.values():
            v.pop("_wall", None)
        line["configs"] = ex

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, rig.affinity0)          # the CPU leg uses every core of the box, not just the GPU's NUMA node
        threads = host_threads()
        if cfg["kind"] == "atlas":
            cb = cpu_atlas_rates(parts["main"][0], parts["main"][1], cfg["size"], cfg["mode"], args.cpu_sample, threads)
        else:
            from msdfgen_b200 import prep
            from msdfgen_b200.synth import synth_big_shape, glyph_transform
            cb = cpu_tiled_rates(prep.prepare(synth_big_shape(cfg["contours"], cfg["segments"])), glyph_transform(cfg["size"], px_range=8.0), cfg["size"], 32, threads)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "threads_used", "single_thread_value", "row_parallel_value")}
    if rank == 0:
        print(json.dumps(line))
    rig.ctx.close()
    if world > 1:
        rig.dist.destroy_process_group()


if __name__ == "__main__":
    main()
