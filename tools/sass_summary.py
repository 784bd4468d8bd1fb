#!/usr/bin/env python
"""SASS evidence for profiles/: per-function static size and instruction mix of the solve kernel (the phase bodies
are __noinline__ device functions inside it), the asynchronous-copy / FP64 mnemonics the design relies on, and
excerpts of the Riccati inner loop and of a cp.async ring.
usage: tools/sass_summary.py [out.txt]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "cilqr_b200", "lib", "cilqr_capi.o")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
# functions: ".type NAME,@function" ... until the next one
segs, cur = [], None
for l in dis:
    m = re.match(r"\s*\.type\s+(\S+),@function", l)
    if m:
        cur = [m.group(1), []]
        segs.append(cur)
    elif cur is not None:
        cur[1].append(l)


def pretty(n):
    n = n.split("$")[-1] if "$_ZN5cilqr" in n else n.lstrip("$")
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*", "", d)


out = [f"# nvdisasm of {os.path.relpath(obj, ROOT)} (sm_100a, nvcc -O3 -lineinfo): static SASS per function of the solve kernel",
       "# an SM feeds unaligned warps at full rate only while their hot code fits ~32 KB (DESIGN.md 2.1): every phase body below is",
       "# its own __noinline__ function, and the CTA runs one phase type at a time", ""]
out.append(f"{'function':44s} {'instr':>6s} {'bytes':>7s} {'DFMA':>5s} {'DADD':>5s} {'DMUL':>5s} {'LDG':>4s} {'LDGSTS':>6s} {'LDS':>4s} {'STS':>4s} {'STG':>4s} {'LDL':>4s} {'STL':>4s}")
total = collections.Counter()
rows = []
for name, lines in segs:
    if "cilqr" not in name:
        continue
    ops = [m.group(1) for l in lines for m in [re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)] if m]
    c = collections.Counter(ops)
    total.update(c)
    rows.append((pretty(name), len(ops), c, lines))
for name, n, c, _ in sorted(rows, key=lambda r: -r[1]):
    out.append(f"{name[:44]:44s} {n:6d} {n * 16:7d} {c['DFMA']:5d} {c['DADD']:5d} {c['DMUL']:5d} {c['LDG']:4d} {c['LDGSTS']:6d} {c['LDS']:4d} "
               f"{c['STS']:4d} {c['STG']:4d} {c['LDL']:4d} {c['STL']:4d}")
n_all = sum(r[1] for r in rows)
out += ["", f"solve kernel in total: {n_all} instructions = {n_all * 16 / 1024:.0f} KB; mix: " + ", ".join(f"{k} {v}" for k, v in total.most_common(20)),
        f"LDGSTS (cp.async: global -> shared without passing registers; rings of ROLL / BACK, plane tiles of LIN / EVAL): {total['LDGSTS']}",
        f"tensor-core / TMA mnemonics (UTCHMMA, HMMA, UTMALDG, UBLKCP): {total['UTCHMMA'] + total['HMMA'] + total['UTMALDG'] + total['UBLKCP']} "
        "-- none: the dense blocks are 6x6 / 6x2 / 2x2 fp64 and serial in the knot index (DESIGN.md 2.2)"]
for key, title in (("backward_pass", "Riccati step (S2: Qh = Hh + F^T G): operands from shared memory (LDS.64), DFMA chains"),
                   ("roll_multi", "two-stage cp.async ring of the rollout: LDGSTS.E.BYPASS.128 + LDGDEPBAR / DEPBAR")):
    for name, n, c, lines in rows:
        if key in name:
            code = [l for l in lines if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
            pick = "DFMA" if key == "backward_pass" else "LDGSTS"
            idx = [i for i, l in enumerate(code) if pick in l]
            if idx:
                s0 = max(0, idx[len(idx) // 3] - 8)
                out += ["", f"excerpt of {name} -- {title}:"]
                out += ["  " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip() for l in code[s0:s0 + 30]]
            break
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt + "\n")
