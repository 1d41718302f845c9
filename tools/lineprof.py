"""Join ncu's per-SASS-instruction counts with nvdisasm line info -> instructions executed per source line.
usage: python tools/lineprof.py <ncu_source_sass.csv> <nvdisasm -g -c output> <mangled_kernel_name> [top] [name:lo-hi ...]"""
import collections
import csv
import re
import sys

srccsv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(srccsv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = hdr.index("Instructions Executed")
si = hdr.index("# Samples")
inst = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > ci:
        inst.append((r[1], int(float(r[ci] or 0)), int(float(r[si] or 0))))
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = None
seq = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2)))
print("sass instrs in csv:", len(inst), " in nvdisasm:", len(seq))
agg = collections.Counter()
smp = collections.Counter()
ops = collections.Counter()
n = min(len(inst), len(seq))
for k in range(n):
    agg[seq[k][0]] += inst[k][1]
    smp[seq[k][0]] += inst[k][2]
    ops[seq[k][1].split()[0] if not seq[k][1].startswith("@") else seq[k][1].split()[1]] += inst[k][1]
tot = sum(agg.values())
ts = sum(smp.values())
print("total warp instr:", tot)
src = {}
for (f, ln), c in agg.most_common(top):
    if f not in src:
        try:
            src[f] = open(__import__("os").environ.get("LINEPROF_SRC", "/root/repo/drone_image_stitch_cpp_b200/csrc/") + f).read().split("\n")
        except Exception:
            src[f] = []
    text = src[f][ln - 1].strip()[:90] if src[f] and ln <= len(src[f]) else ""
    print(f"{100*c/tot:5.1f}% inst {100*smp[(f,ln)]/max(ts,1):5.1f}% smp  {f}:{ln}  {text}")
print("---- by opcode")
for o, c in ops.most_common(25):
    print(f"{100*c/tot:5.1f}%  {o}")
if len(sys.argv) > 5:
    rng = []
    for a in sys.argv[5:]:
        nm, r = a.split(":")
        lo, hi2 = r.split("-")
        rng.append((nm, int(lo), int(hi2)))
    out = collections.Counter()
    outs = collections.Counter()
    for (f, ln), c in agg.items():
        key = "other:" + str(f)
        if f == "ds_kernels.h":
            for nm, lo, hi2 in rng:
                if lo <= ln <= hi2:
                    key = nm
                    break
        out[key] += c
        outs[key] += smp[(f, ln)]
    print("---- by phase")
    for k2, c in out.most_common():
        print(f"{100*c/tot:5.1f}% inst {100*outs[k2]/max(ts,1):5.1f}% smp  {k2}")
    # positional attribution: an instruction whose line is outside every named range (inlined helpers in other files
    # or at the top of ds_kernels.h) belongs to the phase of the nearest earlier instruction that has one
    pos = collections.Counter()
    poss = collections.Counter()
    curp = "prologue"
    for k in range(n):
        f, ln = seq[k][0] if seq[k][0] else ("", 0)
        if f == "ds_kernels.h":
            for nm, lo, hi2 in rng:
                if lo <= ln <= hi2:
                    curp = nm
                    break
        pos[curp] += inst[k][1]
        poss[curp] += inst[k][2]
    print("---- by phase, positional")
    for k2, c in pos.most_common():
        print(f"{100*c/tot:5.1f}% inst {100*poss[k2]/max(ts,1):5.1f}% smp  {c/1e6:8.1f} M  {k2}")
