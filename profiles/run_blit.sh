cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blit" 2>&1 | tail -2
SRB_BULK_DETILE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blit" 2>&1 | tail -2
for sz in "1920 1080" "3840 2160"; do for k in "" SRB_BULK_DETILE=1; do
  env $k ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:detile -s 2 -c 3 --csv python profiles/prof_blit.py $sz 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r]
if hi:
    h=rows[hi[0]]
    for r in rows[hi[0]+1:]:
        if len(r)==len(h) and r[h.index('Metric Name')]=='gpu__time_duration.sum': print('$sz', '$k', r[h.index('Kernel Name')][:40], r[h.index('Metric Value')], 'ns')
"
done; done
