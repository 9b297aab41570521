#!/bin/bash
# Run on the GPU box (gpurun): launch list + full ncu captures of the frame's kernels on the bench workload (hall 1080p).
# usage: bash profiles/capture.sh <tag>
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 14 -c 70 --csv --log-file $out/${tag}_launches_hall.csv python profiles/prof_frames.py hall 12 > /dev/null 2>&1
for k in setup_kernel bin_fill_kernel raster_kernel shade_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $out/${tag}_${k}_hall python profiles/prof_frames.py hall 4 > /dev/null 2>&1
done
# the launch list of bench.py itself (a short run: ncu serialises every kernel, so only the SHARES mean anything)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 480 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --frames-per-step 16 --no-cpu-baseline > /dev/null 2>&1
ls -la $out
