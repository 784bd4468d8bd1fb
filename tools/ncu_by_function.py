#!/usr/bin/env python
"""Per-function view of one ncu capture of the solve kernel: executed instructions, stall samples and the main stall
reasons for each __noinline__ phase function, plus the dynamic opcode mix.

ncu's source page lists the kernel's SASS in address order without function names; nvdisasm of the SAME build lists the
same instructions in the same order with the function labels.  The two are joined by position (and the opcode of every
row is compared, so a stale object file is detected, not silently mis-attributed).

usage: tools/ncu_by_function.py gpurun_out/prof.ncu-rep [cilqr_b200/lib/cilqr_capi.o] > profiles/rNN_x_by_function.txt
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "cilqr_b200", "lib", "cilqr_capi.o")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")

# instructions of the solve kernel's section, in order, each with the function it belongs to
seq, fn, inside = [], None, False
for ln in dis:
    m = re.match(r"^(\$?[_A-Za-z$][\w$]*):\s*$", ln)
    if m:
        name = m.group(1)
        if name.startswith(".text."):
            continue
        if name.startswith("_ZN5cilqr18cilqr_solve_kernel"):
            inside, fn = True, "cilqr_solve_kernel (scheduler)"
        elif name.startswith("$_ZN5cilqr18cilqr_solve_kernel"):
            inside = True
            d = subprocess.run(["c++filt", name.split("$")[2]], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", d).replace("cilqr::", "") or name.split("$")[2]
        elif name.startswith("$__internal") and inside:
            fn = name.split("$")[-1]
        elif name.startswith("_Z") or (name.startswith("$_Z") and "cilqr_solve_kernel" not in name):
            inside = False
        continue
    m = re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and inside:
        op = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip()).split()[0]
        seq.append((fn, op))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, body = rows[h], [r for r in rows[h + 1:] if len(r) > 10]
col = {n: i for i, n in enumerate(hdr)}
if len(body) != len(seq):
    sys.exit(f"instruction counts differ: ncu {len(body)} vs nvdisasm {len(seq)} -- the object file is not the captured build")
bad = 0
for r, (f, op) in zip(body, seq):
    o = re.sub(r"^\s*@!?U?P\w+\s+", "", r[col["Source"]].strip()).split()[0]
    bad += (o.split(".")[0] != op.split(".")[0])
if bad:
    sys.exit(f"{bad} opcodes differ between ncu and nvdisasm -- the object file is not the captured build")

stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
F = collections.OrderedDict()
ops_e, ops_s = collections.Counter(), collections.Counter()
TE = TS = 0
for r, (f, op) in zip(body, seq):
    e, s = int(r[col["Instructions Executed"]]), int(r[col["# Samples"]])
    d = F.setdefault(f, {"n": 0, "e": 0, "s": 0, "st": collections.Counter()})
    d["n"] += 1; d["e"] += e; d["s"] += s
    for n in stalls:
        d["st"][n] += int(r[col[n]] or 0)
    ops_e[op.split(".")[0]] += e; ops_s[op.split(".")[0]] += s
    TE += e; TS += s

print(f"# {os.path.basename(rep)} joined with nvdisasm of {os.path.relpath(obj, ROOT)}: {len(seq)} SASS instructions, "
      f"{TE:.3e} warp instructions executed, {TS} stall samples")
print(f"{'function':34s} {'static':>6s} {'exec %':>7s} {'samples %':>9s}  top stall reasons (share of the function's samples)")
for f, d in sorted(F.items(), key=lambda kv: -kv[1]["s"]):
    tot = sum(d["st"].values()) or 1
    top = ", ".join(f"{n[6:]} {v / tot * 100:.0f}%" for n, v in d["st"].most_common(4))
    print(f"{f[:34]:34s} {d['n']:6d} {d['e'] / TE * 100:7.2f} {d['s'] / TS * 100:9.2f}  {top}")
allst = collections.Counter()
for d in F.values():
    allst.update(d["st"])
tot = sum(allst.values())
print("\nall samples by reason: " + ", ".join(f"{n[6:]} {v / tot * 100:.1f}%" for n, v in allst.most_common(10)))
print("\ndynamic opcode mix (share of executed warp instructions / of samples):")
for op, e in ops_e.most_common(22):
    print(f"  {op:8s} {e / TE * 100:5.1f}% {ops_s[op] / TS * 100:5.1f}%")
fp = sum(ops_e[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print(f"  FP64 arithmetic (DFMA+DMUL+DADD+DSETP+MUFU): {fp / TE * 100:.1f}% of executed instructions")
