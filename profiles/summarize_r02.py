"""Turns the captures of profiles/capture_r02.sh (in gpurun_out/) into the summaries committed under profiles/:
  <tag>_<kernel>_ncu.txt     key metrics + hottest SASS lines of the ncu --set full capture of each kernel
  <tag>_launches_{hall,bench}.{csv,txt}   launch lists (shares of the step)
  <tag>_atomics.txt          L2 atomic / reduction traffic of the front end against the atomic unit's peak
  <tag>_texel_taps.txt       L1 / L2 hit rates of the texel taps alone (differential: real taps - all taps on texel 0)
  ncu_counts.json            per-launch DRAM bytes and executed warp instructions that bench.py reports
usage: python profiles/summarize_r02.py <tag> [git-rev the captures were taken at]"""
import csv, io, json, os, shutil, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rev = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__inst_executed_op_shared_atom.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def num(x):
    return float(x.replace(",", ""))


counts = {"_capture": {"tag": tag, "tree": rev, "workload": "hall 1920x1080, frame 388 of the camera path (profiles/prof_frames.py hallpath)",
                       "made_by": "profiles/capture_r02.sh + profiles/summarize_r02.py"}}
names = {"setup_direct_kernel": "setup", "clip_scan_kernel": "clip_scan", "bin_fill_kernel": "bin_fill", "raster_kernel": "raster", "shade_kernel": "shade"}
for k, short in names.items():
    rep = os.path.join(src, f"{tag}_{k}_hall.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, vals = rows[0], rows[1], rows[2]
    kname = vals[h.index("Kernel Name")] if "Kernel Name" in h else k
    lines = [f"# ncu --set full --clock-control none, kernel {kname}, workload hall 1920x1080 on the camera path (one launch, cold caches, serialised); tree {rev}"]
    d = {}
    for m in RAW:
        if m in h:
            i = h.index(m)
            lines.append(f"{m:75s} {vals[i]:>18s} {units[i]}")
            d[m] = vals[i]
    if "dram__bytes_read.sum" in d:
        rd = num(d["dram__bytes_read.sum"]) * SCALE[units[h.index("dram__bytes_read.sum")]]
        wr = num(d["dram__bytes_write.sum"]) * SCALE[units[h.index("dram__bytes_write.sum")]]
        counts[short] = {"dram_bytes": int(rd + wr), "warp_inst": int(num(d.get("smsp__inst_executed.sum", "0"))),
                         "duration_us": num(d["gpu__time_duration.sum"]),
                         "issue_active_pct": num(d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "0")),
                         "capture": f"{tag}_{k}_hall.ncu-rep"}
        lines.append(f"dram traffic per launch (read+write) = {int(rd + wr)} bytes")
    hot = subprocess.run(f"ncu -i {rep} --page source --csv | python {os.path.join(dst, 'ncu_hot.py')} 25", shell=True, capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{tag}_{short}_ncu.txt"), "w").write("\n".join(lines) + "\n\n# hottest SASS lines (warp stall samples)\n" + hot)


def metric_rows(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    out = []
    for r in rows[hi + 1:]:
        if len(r) == len(h):
            out.append(dict(zip(h, r)))
    return out


for which, what in (("hall", "profiles/prof_frames.py hallpath 16 (50 launches after the first frames)"),
                    ("bench", "bench.py --steps 2 --warmup 3 --frames-per-step 16 --no-cpu-baseline --no-configs --no-geometry-upload (400 launches from the timed steps)")):
    lcsv = os.path.join(src, f"{tag}_launches_{which}.csv")
    if not os.path.exists(lcsv):
        continue
    agg = {}
    for r in metric_rows(lcsv):
        name = r["Kernel Name"].split("(")[0].split("::")[-1].split("<")[0]
        agg.setdefault(name, []).append(num(r["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(dst, f"{tag}_launches_{which}.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: launches of {what}; tree {rev}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:24s} launches {len(v):3d}  mean {sum(v)/len(v)/1000:8.1f} us  share {100*sum(v)/tot:5.1f} %\n")
    shutil.copy(lcsv, os.path.join(dst, f"{tag}_launches_{which}.csv"))

# ---- L2 atomics of the front end --------------------------------------------------------------------------------------
with open(os.path.join(dst, f"{tag}_atomics.txt"), "w") as f:
    f.write(f"# L2 atomic traffic of the front-end kernels (ncu --clock-control none, one frame, tree {rev}).\n"
            "# atom = atomics that return a value (ATOMG), red = reductions (REDG); sectors at the L2 (lts__t_sectors_op_*),\n"
            "# rate = sectors / kernel duration, unit_busy = lts__d_atomic_input_cycles_active, % of its peak over the kernel.\n"
            "# The binning atomics are CTA-aggregated through shared memory (smem_atom = shared-memory atomic instructions),\n"
            "# which is why so few reach the L2: its atomic unit is idle (well under 1 % busy) in every kernel.\n")
    for sc, label in (("hallpath", "hall 1080p (155 k tile references)"), ("rand", "1 M random triangles (1.01 M tile references)")):
        p = os.path.join(src, f"{tag}_atomics_{sc}.csv")
        if not os.path.exists(p):
            continue
        per = {}
        for r in metric_rows(p):
            name = r["Kernel Name"].split("(")[0].split("::")[-1]
            per.setdefault((r["ID"], name), {})[r["Metric Name"]] = num(r["Metric Value"])
        f.write(f"\n## {label}\n{'kernel':18s} {'us':>7s} {'atom sect':>10s} {'red sect':>10s} {'atom G/s':>9s} {'red G/s':>9s} {'unit_busy %':>11s} {'smem_atom':>10s}\n")
        for (_, name), m in per.items():
            f.write(f"{name:18s} {m['gpu__time_duration.sum']/1000:7.1f} {m['lts__t_sectors_op_atom.sum']:10.0f} {m['lts__t_sectors_op_red.sum']:10.0f} "
                    f"{m['lts__t_sectors_op_atom.sum.per_second']/1e9:9.3f} {m['lts__t_sectors_op_red.sum.per_second']/1e9:9.3f} "
                    f"{m['lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed']:11.2f} {m['smsp__inst_executed_op_shared_atom.sum']:10.0f}\n")
        shutil.copy(p, os.path.join(dst, f"{tag}_atomics_{sc}.csv"))

# ---- texel taps ----------------------------------------------------------------------------------------------------------
with open(os.path.join(dst, f"{tag}_texel_taps.txt"), "w") as f:
    f.write(f"# L1 and L2 hit rates of the TEXEL TAPS of the shade kernel alone (tree {rev}).\n"
            "# Differential measurement with the statistics build (make STATS=1): the same frame is shaded twice under ncu, once as it\n"
            "# is and once with every tap reading texel 0 of its texture (profiles/prof_frames.py ... nulltaps); everything else the\n"
            "# kernel loads (keys, shade records, RCPPS table) is identical, so the differences of the sector counters are the taps'.\n")
    for sc, label in (("hallpath", "hall 1080p: 25 textures, 2.07 M shaded pixels"), ("rand", "1 M random triangles: one 1024^2 texture, random UVs")):
        pr, pn = os.path.join(src, f"{tag}_taps_{sc}_real.csv"), os.path.join(src, f"{tag}_taps_{sc}_null.csv")
        if not (os.path.exists(pr) and os.path.exists(pn)):
            continue
        a = {r["Metric Name"]: num(r["Metric Value"]) for r in metric_rows(pr)}
        b = {r["Metric Name"]: num(r["Metric Value"]) for r in metric_rows(pn)}
        l1s = a["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"] - b["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
        l1h = a["l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum"] - b["l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum"]
        l2s = a["lts__t_sectors_srcunit_tex_op_read.sum"] - b["lts__t_sectors_srcunit_tex_op_read.sum"]
        l2h = a["lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"] - b["lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"]
        dr = a["dram__bytes_read.sum"] - b["dram__bytes_read.sum"]
        f.write(f"\n## {label}\n")
        f.write(f"                                   real taps      taps on texel 0      difference = the taps\n")
        f.write(f"L1 sectors (global loads)        {a['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']:12.0f} {b['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']:18.0f} {l1s:18.0f}\n")
        f.write(f"L1 sector hits                   {a['l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum']:12.0f} {b['l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum']:18.0f} {l1h:18.0f}\n")
        f.write(f"L2 read sectors from the SMs     {a['lts__t_sectors_srcunit_tex_op_read.sum']:12.0f} {b['lts__t_sectors_srcunit_tex_op_read.sum']:18.0f} {l2s:18.0f}\n")
        f.write(f"L2 read sector hits              {a['lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum']:12.0f} {b['lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum']:18.0f} {l2h:18.0f}\n")
        f.write(f"DRAM bytes read                  {a['dram__bytes_read.sum']:12.0f} {b['dram__bytes_read.sum']:18.0f} {dr:18.0f}\n")
        f.write(f"kernel duration under ncu (us)   {a['gpu__time_duration.sum']/1000:12.1f} {b['gpu__time_duration.sum']/1000:18.1f}\n")
        f.write(f"=> texel taps: L1 hit rate {100*l1h/max(l1s,1):.1f} %, L2 hit rate {100*l2h/max(l2s,1):.1f} %, "
                f"{l1s*32/max(a['smsp__inst_executed.sum'],1):.2f} L1 bytes per warp instruction, {dr/1e6:.2f} MB from DRAM per frame\n")
        counts.setdefault("texel_taps", {})[sc] = {"l1_hit_rate": l1h / max(l1s, 1), "l2_hit_rate": l2h / max(l2s, 1), "l2_sectors": l2s, "dram_bytes": dr}
        for p in (pr, pn):
            shutil.copy(p, os.path.join(dst, os.path.basename(p)))
json.dump(counts, open(os.path.join(dst, "ncu_counts.json"), "w"), indent=1)
print(json.dumps(counts, indent=1))
