"""Join ncu's per-SASS-instruction counts with nvdisasm line info -> instructions executed per source line.
usage: python tools_lineprof.py <ncu_source_sass.csv> <all.sass> <mangled_kernel_name> [top]"""
import csv, re, sys, collections
srccsv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(srccsv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
inst = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > ci:
        inst.append((r[1], int(float(r[ci] or 0)), int(float(r[si] or 0))))
# parse nvdisasm
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((cur, m.group(2)))
print("sass instrs in csv:", len(inst), " in nvdisasm:", len(seq))
agg = collections.Counter(); smp = collections.Counter()
n = min(len(inst), len(seq))
for k in range(n):
    agg[seq[k][0]] += inst[k][1]; smp[seq[k][0]] += inst[k][2]
tot = sum(agg.values()); ts = sum(smp.values())
print("total warp instr:", tot)
src = {}
for (f, ln), c in agg.most_common(top):
    if f not in src:
        try: src[f] = open("/root/repo/drone_image_stitch_cpp_b200/csrc/" + f).read().split("\n")
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if src[f] and ln <= len(src[f]) else ""
    print(f"{100*c/tot:5.1f}% inst {100*smp[(f,ln)]/max(ts,1):5.1f}% smp  {f}:{ln}  {text}")
