#!/usr/bin/env python
"""ncu report of ONE solve launch at the bench config -> profiles/traffic.json (dram bytes per launch).
usage: tools/ncu_traffic.py gpurun_out/prof.ncu-rep "<workload note>" """
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
    return float(vals[i]) * mult
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "gpu_time_ms_under_ncu": float(vals[hdr.index("gpu__time_duration.sum")]),
       "source": os.path.basename(rep), "workload": note}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
