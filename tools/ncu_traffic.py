#!/usr/bin/env python
"""ncu report of ONE solve launch at the bench config -> profiles/traffic.json: DRAM bytes and FP64 operations per
launch, tied to the kernel sources they were captured from (bench.py only reports them while the hash matches).

usage: tools/ncu_traffic.py gpurun_out/prof.ncu-rep <scenarios> <horizon> "<workload note>"
The capture must include the three FP64 counters:
  ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum ...
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_hash)

rep = sys.argv[1]
scen, hor = int(sys.argv[2]), int(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]


def get(name, default=None):
    if name not in hdr:
        return default
    i = hdr.index(name)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[i], 1)
    return float(vals[i].replace(",", "")) * mult


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
dfma = get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
dadd = get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
dmul = get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
path = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
out = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "fp64_flops_per_launch": (2 * dfma + dadd + dmul) if dfma is not None else None,
       "fp64_thread_inst": {"dfma": dfma, "dadd": dadd, "dmul": dmul},
       "gpu_time_ms_under_ncu": get("gpu__time_duration.sum") / (1e6 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1),
       "source": os.path.basename(rep), "workload": note, "scenarios": scen, "horizon": hor,
       "kernel_source_sha1": bench.kernel_source_hash(), "commit": commit + " (+ working tree)"}
for k in ("corridor", "dp"):
    if k in old:
        out[k] = old[k]
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
