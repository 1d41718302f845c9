"""Per-kernel SASS evidence of the built library: instruction counts that show what each kernel is made of
(TMA tile loads UTMALDG, TMA L2 prefetches UTMAPF, integer dot products IDP, global / shared loads and stores, barriers,
tensor-core ops - none are expected on this path), with the hash of the sources the library was built from.
usage: python tools/sass_summary.py [out.json]     (default: profiles/sass_summary.json)"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drone_image_stitch_cpp_b200 import build  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.json")
so = build.build_cuda()
sass = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
res = subprocess.check_output(["cuobjdump", "-res-usage", so], text=True, stderr=subprocess.STDOUT)
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    usage[m.group(1)] = {"registers": int(m.group(2)), "stack_bytes": int(m.group(3)), "static_shared": int(m.group(4))}
kernels = {}
cur = None
WATCH = ["UTMALDG", "UTMAPF", "UTMASTG", "UBLKCP", "IDP", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "SYNCS", "FFMA", "IMAD", "PRMT",
         "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "F2I", "MUFU", "ATOMS", "ATOMG", "RED"]
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("UTMALDG", "UTMAPF", "IDP", "HMMA") and op.startswith(w)):
                kernels[cur][w] += 1
        if op.startswith("IDP.2A"):
            kernels[cur]["IDP.2A"] += 1
        if op.startswith("IDP.4A"):
            kernels[cur]["IDP.4A"] += 1
def demangle(k):
    m = re.match(r"_Z(\d+)", k)
    return k[len(m.group(0)):len(m.group(0)) + int(m.group(1))] if m else k


doc = {"library": os.path.relpath(so, ROOT), "kernel_source_sha16": build.source_hash(),
       "how": "cuobjdump -sass / -res-usage of the in-tree build (tools/sass_summary.py)",
       "kernels": {}}
for k, c in sorted(kernels.items()):
    name = demangle(k)
    d = {w: c[w] for w in ["total"] + WATCH + ["IDP.2A", "IDP.4A"] if c[w]}
    d.update(usage.get(k, {}))
    doc["kernels"][name] = d
doc["tensor_core_instructions"] = sum(c["HMMA"] + c["UTCHMMA"] + c["UTCQMMA"] for c in kernels.values())
doc["tma_tile_loads"] = {demangle(k): c["UTMALDG"] for k, c in kernels.items() if c["UTMALDG"]}
json.dump(doc, open(out, "w"), indent=1)
print(out, doc["kernel_source_sha16"], doc["tma_tile_loads"])
