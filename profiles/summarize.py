#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep + the launch list into the text summaries committed under profiles/ (run in the build container).

    python profiles/summarize.py r1
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = os.environ.get("SUMMARY_OUT") or os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_branch_targets_threads_divergent.sum", "smsp__sass_branch_targets.sum"]

traffic = {}
for kernel in ("k1a_filter", "k1b_exact", "k2_classify", "k3_exact_distance", "k1_big_blocks"):
    rep = os.path.join(ROOT, "gpurun_out", f"{kernel}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    what = ("one launch of `SIZE=2048 MODE=1 python tools/_gpu_big_one.py`: the 10k-segment shape, sdf 2048x2048, overlapping combiner"
            if kernel == "k1_big_blocks" else "one launch of `python bench.py --steps 1 --warmup 3`, 4096 glyphs msdf 32x32")
    lines = [f"# {kernel} — ncu --set full --clock-control none ({what})",
             f"# kernel: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else kernel}", ""]
    get = {}
    for i, h in enumerate(hdr):
        if h in KEYS or "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            lines.append(f"{h:78s} {units[i]:12s} {vals[i]}")
            get[h] = (units[i], vals[i])

    def to_bytes(k):
        u, v = get.get(k, ("byte", "0"))
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    traffic[kernel if kernel != "k2_classify" else "k2_error_correction"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    # opcode histogram
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    ia, ie, it = sh.index("Source"), sh.index("Instructions Executed"), sh.index("Thread Instructions Executed")
    ops, thr, tot = collections.Counter(), collections.Counter(), 0
    for r in srows[2:]:
        if len(r) <= ie:
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
        op = m.group(2).split(".")[0] if m else "?"
        n = int(r[ie] or 0)
        ops[op] += n
        thr[op] += int(r[it] or 0)
        tot += n
    lines += ["", f"SASS opcode histogram (warp instructions executed, total {tot}):"]
    for op, n in ops.most_common(18):
        lines.append(f"  {op:10s} {n:12d} {100*n/max(tot,1):5.1f}%   avg active threads {thr[op]/max(n,1):5.1f}")
    sass = " ".join(r[ia] for r in srows[2:] if len(r) > ia)
    lines += ["", "Blackwell evidence in SASS: " + ", ".join(f"{k}={'yes' if k in sass else 'no'}" for k in ("UBLKCP", "UTMALDG", "LDGSTS", "LDS", "STS", "DFMA", "MUFU.RSQ64H"))]
    open(os.path.join(OUT, f"{kernel}_{tag}.txt"), "w").write("\n".join(lines) + "\n")
    print("wrote", f"profiles/{kernel}_{tag}.txt")

launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 8]
    h = rows[0]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(re.sub(r"\(.*", "", r[ki]), []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in agg.values())
    out = [f"# launch list: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 (cold-cache, serialised: compare SHARES)",
           f"{'kernel':60s} {'launches':>8s} {'mean_us':>10s} {'share':>7s}"]
    for k, v in agg.items():
        out.append(f"{k[:60]:60s} {len(v):8d} {sum(v)/len(v)/1e3:10.1f} {100*sum(v)/total:6.1f}%")
    open(os.path.join(OUT, f"launches_{tag}.txt"), "w").write("\n".join(out) + "\n")
    print("wrote", f"profiles/launches_{tag}.txt")
if traffic:
    json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print("traffic (bytes per launch):", traffic)
