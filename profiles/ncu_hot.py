"""Summarises an ncu --page source --csv dump: total stall samples per reason and the hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_hot.py [top_n]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: 0 for n in stall_cols}
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    s = int(r[col["# Samples"]] or 0)
    for n in stall_cols:
        tot[n] += int(r[col[n]] or 0)
    lines.append((s, r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0),
                  {n: int(r[col[n]] or 0) for n in stall_cols if int(r[col[n]] or 0) > 0}))
total = sum(l[0] for l in lines)
print("total samples", total, "instructions", sum(l[2] for l in lines))
print("by reason:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for i, (s, src, ex, st) in enumerate(lines):
    pass
order = sorted(range(len(lines)), key=lambda i: -lines[i][0])[:n]
for i in sorted(order):
    s, src, ex, st = lines[i]
    print(f"{i:5d} {s:6d} {100.0*s/total:5.1f}% exec={ex:8d} {src[:70]:70s} {st}")
