"""Assembles profiles/r2_scale_cfg4_strong.json from bench.py lines taken at several N (builder's own gpurun calls).
usage: python tools/make_scale_profile.py out.json what n1.json [n2.json ...]"""
import json, sys
out, what, files = sys.argv[1], sys.argv[2], sys.argv[3:]
runs = []
for f in files:
    d = json.loads(open(f).read().strip().split("\n")[-1])
    r = d["roofline"]
    runs.append({"n_gpus": d["n_gpus"], "value_MPps": d["value"], "ms_per_composite": d["ms_per_step"], "steps": d["steps"],
                 "halo": d["run"]["halo"], "band_edges": d["run"]["band_edges"], "frames_rank0": d["run"]["frames_rank0"],
                 "device_GB_rank0": d["run"]["device_GB_rank0"], "parity": d["parity"], "e2e": d["e2e"],
                 "whole_step": r["whole_step"], "dominant": {k: r[k] for k in ("kernel", "achieved", "frac", "ms_per_launch")},
                 "kernels_rank0": r["kernels"], "clocks": d["clocks"], "config": d["config"], "host_affinity": d["run"]["host_affinity"]})
runs.sort(key=lambda r: r["n_gpus"])
base = next((r for r in runs if r["n_gpus"] == 1), None)
for r in runs:
    if base:
        r["speedup_vs_1"] = r["value_MPps"] / base["value_MPps"]
        r["efficiency"] = r["speedup_vs_1"] / r["n_gpus"]
json.dump({"what": what, "runs": runs}, open(out, "w"), indent=1)
for r in runs:
    print(r["n_gpus"], round(r["value_MPps"]), round(r["ms_per_composite"], 2), r.get("efficiency"), r["parity"].get("identical"), round(r["e2e"]["value"]))
