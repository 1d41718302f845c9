"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarise_profiles.py <tag> <launches.csv> <sources-hash-file> [<kernel>=<file.ncu-rep> ...]
  tag                round tag of the output names (r2 -> profiles/r2_launches_cfg2.json ...)
  sources-hash-file  the .so.sources file copied from the GPU box next to the capture (hash of the kernel sources the
                     captured library was built from); bench.py only reports roofline.traffic when it matches its own build"""
import collections
import csv
import json
import subprocess
import sys

tag, launch_csv, hash_file = sys.argv[1], sys.argv[2], sys.argv[3]
reps = [a.split("=", 1) for a in sys.argv[4:]]
SHA = open(hash_file).read().strip()
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]   # GB/s, measured copy bandwidth of this pool
rows = list(csv.reader(open(launch_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ki, mi, vi, idi, gi, bi = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size", "Block Size"))
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[idi], {"name": r[ki], "grid": r[gi], "block": r[bi]})
    d[r[mi]] = float(r[vi].replace(",", ""))
mine = [l for l in launch.values() if l["name"].startswith("ds_")]
starts = [i for i, l in enumerate(mine) if l["name"].startswith("ds_mb_feed_l0")]
start = starts[4]          # 3 warm-up + 3 timed composites, then the e2e ones: take a timed one
seq = []
for l in mine[start:]:
    if seq and l["name"].startswith("ds_mb_feed_l0"):
        break
    if l["name"].startswith("ds_mb_"):
        seq.append(l)
tot = sum(l["gpu__time_duration.sum"] for l in seq)
out = [dict(kernel=l["name"].split("(")[0], grid=l["grid"], block=l["block"], us=round(l["gpu__time_duration.sum"] / 1e3, 2),
            share=round(l["gpu__time_duration.sum"] / tot, 4), dram_read_MB=round(l["dram__bytes_read.sum"] / 1e6, 1),
            dram_write_MB=round(l["dram__bytes_write.sum"] / 1e6, 1),
            dram_GBps=round((l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) / l["gpu__time_duration.sum"], 1),
            frac_of_measured_hbm_peak=round((l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) / l["gpu__time_duration.sum"] / PEAK, 3))
       for l in seq]
json.dump(dict(command="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
                       "python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-parity",
               kernel_source_sha16=SHA,
               note="one ds_composite of cfg2; per-launch times under ncu are cold-cache / serialised: compare shares. dram_GBps = "
                    "(dram read + write) / duration of the launch: the REAL traffic rate, against MEASURED_PEAKS.json hbm_gbs "
                    "(%.1f GB/s; nominal ~8000). The bench line's roofline.achieved uses the algorithmic bytes instead." % PEAK,
               launches=out, total_us=round(tot / 1e3, 1), dram_total_MB=round(sum(o["dram_read_MB"] + o["dram_write_MB"] for o in out), 1)),
          open(f"profiles/{tag}_launches_cfg2.json", "w"), indent=1)
with open(f"profiles/{tag}_launches_cfg2.csv", "w") as f:
    f.write("idx,kernel,grid,block,time_ns,dram_read_bytes,dram_write_bytes\n")
    for i, l in enumerate(mine):
        f.write(f"{i},{l['name'].split('(')[0]},\"{l['grid']}\",\"{l['block']}\",{l['gpu__time_duration.sum']:.0f},"
                f"{l['dram__bytes_read.sum']:.0f},{l['dram__bytes_write.sum']:.0f}\n")
json.dump({"cfg2:mb_feed:0": int(seq[0]["dram__bytes_read.sum"] + seq[0]["dram__bytes_write.sum"]), "kernel_source_sha16": SHA,
           "capture": f"profiles/{tag}_launches_cfg2.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum of one ds_mb_feed_l0 launch, cfg2)"},
          open("profiles/dominant_kernel_traffic.json", "w"))
for o in out:
    print(o)
print("total us", tot / 1e3)

keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for kname, rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, r = rr[0], rr[1], rr[2]
    summ = {k: (r[h.index(k)] + " " + units[h.index(k)]).strip() for k in keys if k in h}
    summ["kernel_source_sha16"] = SHA
    json.dump(summ, open(f"profiles/{tag}_{kname}_ncu_full.json", "w"), indent=1)
    print(kname)
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__shared_mem_per_block_dynamic", "launch__registers_per_thread"):
        print("  ", k, summ.get(k))
