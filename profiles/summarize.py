"""Turns the ncu captures of profiles/capture.sh (in gpurun_out/) into the text summaries committed under profiles/.
usage: python profiles/summarize.py <tag>"""
import csv, io, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__inst_executed_op_shared_atom.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
traffic = {}
counts = {}
names = {"setup_kernel": "setup", "bin_fill_kernel": "bin_fill", "raster_kernel": "raster", "shade_kernel": "shade"}
for k, short in names.items():
    rep = os.path.join(src, f"{tag}_{k}_hall.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# ncu --set full --clock-control none, kernel {k}, workload hall 1920x1080 (one launch, cold-ish caches, serialised)"]
    d = {}
    for m in RAW:
        if m in h:
            i = h.index(m)
            lines.append(f"{m:75s} {vals[i]:>18s} {units[i]}")
            d[m] = vals[i]
    def num(x):
        return float(x.replace(",", ""))
    if "dram__bytes_read.sum" in d:
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        rd = num(d["dram__bytes_read.sum"]) * scale[units[h.index("dram__bytes_read.sum")]]
        wr = num(d["dram__bytes_write.sum"]) * scale[units[h.index("dram__bytes_write.sum")]]
        traffic[short] = int(rd + wr)
        counts[short] = {"dram_bytes": int(rd + wr), "warp_inst": int(num(d.get("smsp__inst_executed.sum", "0"))),
                         "duration_us": num(d["gpu__time_duration.sum"]),
                         "issue_active_pct": num(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "0")),
                         "capture": f"{tag}_{k}_hall.ncu-rep"}
        lines.append(f"dram traffic per launch (read+write) = {int(rd + wr)} bytes")
    hot = subprocess.run(f"ncu -i {rep} --page source --csv | python {os.path.join(dst, 'ncu_hot.py')} 25", shell=True, capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{tag}_{short}_ncu.txt"), "w").write("\n".join(lines) + "\n\n# hottest SASS lines (warp stall samples)\n" + hot)
for which, what in (("hall", "profiles/prof_frames.py hall 12 (skipping the first frames)"),
                    ("bench", "bench.py --steps 2 --warmup 3 --frames-per-step 16 --no-cpu-baseline (480 launches from the timed steps)")):
  lcsv = os.path.join(src, f"{tag}_launches_{which}.csv")
  if os.path.exists(lcsv):
      rows = list(csv.reader(open(lcsv)))
      hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
      h = rows[hi]
      ki, vi = h.index("Kernel Name"), h.index("Metric Value")
      agg = {}
      for r in rows[hi + 1:]:
          if len(r) > vi:
              name = r[ki].split("(")[0].split("::")[-1]
              agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
      tot = sum(sum(v) for v in agg.values())
      with open(os.path.join(dst, f"{tag}_launches_{which}.txt"), "w") as f:
          f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: launches of {what}\n")
          f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
          for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
              f.write(f"{k:24s} launches {len(v):3d}  mean {sum(v)/len(v)/1000:8.1f} us  share {100*sum(v)/tot:5.1f} %\n")
      import shutil
      shutil.copy(lcsv, os.path.join(dst, f"{tag}_launches_{which}.csv"))
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
json.dump(counts, open(os.path.join(dst, "ncu_counts.json"), "w"), indent=1)
print("traffic", traffic)
