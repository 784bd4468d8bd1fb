#!/usr/bin/env python
"""Static SASS size of the solve kernel per source function (code-footprint / I-cache budget).
usage: tools/sass_size.py [path/to/libcilqr_b200.so]"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cilqr_b200", "lib", "libcilqr_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
text = open(os.path.join(ROOT, "cilqr_b200", "csrc", "cilqr_kernel.cuh")).read().split("\n")
funcs = []
for i, l in enumerate(text, 1):
    m = re.match(r"^(?:template <[^>]*>\s*)?__(?:device|global)__.*?\b(\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))
def fn(ln):
    name = "other"
    for a, n in funcs:
        if a <= ln:
            name = n
    return name
cur = None; curfile = None; cnt = {}
for l in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        curfile, cur = m.group(1), int(m.group(2)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        key = fn(cur) if curfile and curfile.endswith("cilqr_kernel.cuh") else "lib:" + (curfile.split("/")[-1] if curfile else "?")
        cnt[key] = cnt.get(key, 0) + 1
tot = sum(cnt.values())
print(f"total {tot} instructions = {tot * 16 / 1024:.0f} KB")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:30]:
    print(f"{v:7d}  {v * 16 / 1024:6.1f} KB  {k}")
