"""Groups an ncu --page source --csv dump into runs of SASS lines with the same execution count (= basic-block regions)
and prints the regions that matter: where the executed warp instructions of a kernel go.
usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_regions.py [min_total]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
out = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) < len(h):
        continue
    ex = int(r[col["Instructions Executed"]] or 0)
    te = int(r[col["Thread Instructions Executed"]] or 0) if "Thread Instructions Executed" in col else 0
    out.append((k, ex, te, r[col["Source"]].strip()[:60]))
thresh = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
total = sum(o[1] for o in out)
print("total warp instructions", total)
i = 0
while i < len(out):
    j = i
    while j + 1 < len(out) and out[j + 1][1] == out[i][1]:
        j += 1
    n = j - i + 1
    if out[i][1] * n > thresh:
        thr = sum(o[2] for o in out[i:j + 1]) / max(1, out[i][1] * n)
        print(f"{i:4d}-{j:4d} n={n:3d} exec={out[i][1]:8d} total={out[i][1]*n:9d} ({100.0*out[i][1]*n/total:4.1f}%) thr/inst={thr:4.1f}  {out[i][3]}")
    i = j + 1
