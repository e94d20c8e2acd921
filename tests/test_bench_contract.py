"""bench.py contract checks that need no GPU: the reference arm runs on the host cores and prints one JSON line with the
agreed keys; the committed B200 line (profiles/bench_r1_n1.json) carries every key the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
          "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample", "96"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert COMMON <= set(d), COMMON - set(d)
    assert d["impl"] == "reference" and d["metric"] == "msdf_pixels_per_sec" and d["unit"] == "px/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] == d["value"]
    # the reference's own single-thread and OpenMP (row-parallel) rates are numeric fields, and the thread count is the one used
    assert c["single_thread_value"] > 0 and c["row_parallel_value"] > 0 and c["threads_used"] == c["cores"] == d["config"]["host_threads"]


def test_reference_arm_loads_no_repo_library_and_overrides_omp_threads():
    """The arm's shapes are prepared by the reference's own normalize + edgeColoringSimple: no msdfgen_b200/*.so is mapped.
    OMP_NUM_THREADS=1 in the environment (what torchrun exports) must not throttle it."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--ref-sample', '64']; "
            "runpy.run_path('bench.py', run_name='__main__'); "
            "maps = open('/proc/self/maps').read(); "
            "assert 'libmsdfgen_b200' not in maps, 'repo library mapped into the reference arm'; "
            "assert 'libmsdfgen_ref' in maps")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    import bench
    assert d["cpu_baseline"]["threads_used"] == bench.host_threads()


def test_reference_arm_tiled_config():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c4", "--steps", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["scaling"] == "strong" and d["value"] > 0 and "4096x4096" in d["config"]["workload"]


def test_committed_b200_line_has_every_key():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r1_n1.json")))
    assert COMMON | {"roofline", "clocks"} <= set(d)
    assert d["metric"] == "msdf_pixels_per_sec" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"] * 1.05
    assert d["gpu_launches"] >= d["steps"] * 4          # k0, k1, k2, k3 per step
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
