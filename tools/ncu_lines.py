#!/usr/bin/env python
"""Per-source-line / per-region instruction and stall-sample shares of one .ncu-rep (run in the build container).
    python tools/ncu_lines.py gpurun_out/k1_glyph_tiles_f2.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum"]
for i, h in enumerate(hdr):
    if h in KEYS or ("pcsamp_warps_issue_stalled" in h and "not_issued" not in h and float(vals[i].replace(",", "") or 0) > 2000):
        print(f"{h:75s} {units[i]:10s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; cur = None; data = {}
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        data[(cur, int(r[0]))] = (int(r[ie] or 0), int(r[isamp] or 0), r[1])
tot = sum(v[0] for v in data.values()); tots = sum(v[1] for v in data.values())
print("total inst (line-attributed)", tot, "samples", tots)
byfile = collections.Counter(); sfile = collections.Counter()
for (f, l), (n, s, _) in data.items(): byfile[f] += n; sfile[f] += s
for f, n in byfile.most_common(): print(f"  {f:28s} inst {100*n/tot:5.1f}%  samples {100*sfile[f]/max(tots,1):5.1f}%")
for (f, l), (n, s, t) in sorted(data.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{f}:{l:4d} inst {100*n/tot:5.2f}% samp {100*s/max(tots,1):5.2f}%  {t.strip()[:105]}")
