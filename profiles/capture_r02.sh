#!/bin/bash
# Run on the GPU box (gpurun): launch lists, full ncu captures of the frame's kernels on the bench workload (hall 1080p on
# the camera path), the L2 atomic counters of the front end and the differential texel-tap counters of the shade kernel.
# usage: bash profiles/capture_r02.sh <tag>     (then, here: python profiles/summarize_r02.py <tag>)
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# 1. every launch of a few frames with its device time (cold-cache, serialised: SHARES, not absolutes)
$NCU --metrics gpu__time_duration.sum -s 30 -c 50 --csv --log-file $out/${tag}_launches_hall.csv python profiles/prof_frames.py hallpath 16 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -s 200 -c 400 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 3 --frames-per-step 16 --no-cpu-baseline --no-configs --no-geometry-upload > /dev/null 2>&1
# 2. ncu --set full of each kernel of one frame (the 5th frame rendered)
for k in setup_direct_kernel clip_scan_kernel bin_fill_kernel raster_kernel shade_kernel; do
  $NCU --set full --import-source on -k regex:$k -s 4 -c 1 -o $out/${tag}_${k}_hall python profiles/prof_frames.py hallpath 6 > /dev/null 2>&1
done
# 3. L2 atomics of the front end, hall and the 1 M-triangle scene
AT=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_requests_op_atom.sum,lts__t_requests_op_red.sum,lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_atom.sum.per_second,lts__t_sectors_op_red.sum.per_second,gpu__time_duration.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed_op_shared_atom.sum,lts__cycles_elapsed.avg.per_second
for sc in hallpath rand; do
  $NCU --metrics $AT -k regex:"setup_direct_kernel|clip_scan_kernel|bin_fill_kernel|raster_kernel" -s 16 -c 4 --csv --log-file $out/${tag}_atomics_${sc}.csv python profiles/prof_frames.py $sc 6 > /dev/null 2>&1
done
# 4. texel taps of the shade kernel: the same frame with the taps as they are, and with every tap reading texel 0
#    (statistics build) -> the difference of the sector / hit counters is the taps' own traffic
TX=lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,dram__bytes_read.sum,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
export SRB_LIB=$GRAFT_REPO_ROOT/softrast_b200/lib/libsoftrast_b200_stats.so
for sc in hallpath rand; do
  $NCU --metrics $TX -k regex:shade_kernel -s 4 -c 1 --csv --log-file $out/${tag}_taps_${sc}_real.csv python profiles/prof_frames.py $sc 6 > /dev/null 2>&1
  $NCU --metrics $TX -k regex:shade_kernel -s 4 -c 1 --csv --log-file $out/${tag}_taps_${sc}_null.csv python profiles/prof_frames.py $sc 6 nulltaps > /dev/null 2>&1
done
unset SRB_LIB
ls -la $out | tail -30
