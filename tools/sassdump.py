"""SASS instructions (with ncu execution counts) attributed to a range of source lines.
usage: python tools/sassdump.py <ncu_source_sass.csv> <nvdisasm -g -c output> <mangled kernel> <file> <lo> <hi>"""
import csv, re, sys
srccsv, sass, kname, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
rows = list(csv.reader(open(srccsv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = rows[h].index("Instructions Executed")
inst = []
for r in rows[h + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    inst.append(int(float(r[ci] or 0)))
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kname + ":"))
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2)))
tot = 0
for k in range(min(len(inst), len(seq))):
    (f, ln, _), txt = seq[k]
    if f == fname and lo <= ln <= hi:
        print(f"{inst[k]:>12d}  {ln:5d}  {txt}")
        tot += inst[k]
print("total", tot)
