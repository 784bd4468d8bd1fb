#!/usr/bin/env python
"""Summarises an ncu report of the solve kernel: headline metrics + per-function instruction / stall-sample shares.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md] [source file for the function ranges]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "local_load", "local_store"]
out = []
for h, u, v in zip(hdr, units, vals):
    if h in want or "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        out.append(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = None; lines = []
for r in rows:
    if len(r) > 4 and r[0] == "Line No": h2 = r; continue
    if h2 and len(r) > 8 and r[0].isdigit(): lines.append(r)
ci, si = h2.index("Instructions Executed"), h2.index("# Samples")
tot = sum(int(r[ci]) for r in lines if r[ci].isdigit()); tots = sum(int(r[si]) for r in lines if r[si].isdigit())
# function ranges from the source file itself
text = open(sys.argv[3] if len(sys.argv) > 3 else "cilqr_b200/csrc/cilqr_kernel.cuh").read().split("\n")
funcs = []
for i, l in enumerate(text, 1):
    m = re.match(r"^(?:template <[^>]*>\s*)?__(?:device|global)__.*?\b(\w+)\(", l)
    if m: funcs.append((i, m.group(1)))
def fn(ln):
    name = "other"
    for a, n in funcs:
        if a <= ln: name = n
    return name
agg = {}
for r in lines:
    ln = int(r[0]); n = int(r[ci]) if r[ci].isdigit() else 0; s_ = int(r[si]) if r[si].isdigit() else 0
    k = fn(ln); agg.setdefault(k, [0, 0]); agg[k][0] += n; agg[k][1] += s_
out.append(f"\ntotal warp-instructions {tot}, stall samples {tots}")
out.append("function                 instr%  samples%")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:24s} {v[0]/tot*100:6.1f}  {v[1]/tots*100:6.1f}")
out.append("\ntop source lines by stall samples: line samples instr source")
for r in sorted(lines, key=lambda r: -int(r[si]) if r[si].isdigit() else 0)[:30]:
    out.append(f"{r[0]:>5s} {r[si]:>8s} {r[ci]:>11s}  {r[1].strip()[:100]}")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2: open(sys.argv[2], "w").write(txt + "\n")
