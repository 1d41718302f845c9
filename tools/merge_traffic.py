"""profiles/dominant_kernel_traffic.json from the captures of tools/gpu_traffic.sh. usage: python tools/merge_traffic.py gpurun_out/<tag>"""
import csv, json, sys
pre = sys.argv[1]
sha = open(pre + "_so.sources").read().strip()
doc = {"kernel_source_sha16": sha, "capture": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:ds_mb_feed_l0 -s 3 -c 1 on bench.py --workload <w> (tools/gpu_traffic.sh): bytes of one launch"}
for w in ("cfg4", "cfg2"):
    rows = list(csv.reader(open(f"{pre}_traffic_{w}.csv")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    mi, vi, ui = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}
    m = {r[mi]: float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0) for r in rows[hi + 1:] if len(r) > vi}
    doc[f"{w}:mb_feed:0"] = int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"])
    doc[f"{w}:duration_ns_under_ncu"] = m["gpu__time_duration.sum"]
json.dump(doc, open("profiles/dominant_kernel_traffic.json", "w"), indent=1)
print(doc)
