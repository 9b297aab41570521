#!/bin/bash
# raster CTAs per SM x smallest unit, one and twelve frames in flight
cd "$GRAFT_REPO_ROOT" || exit 1
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["scene"], d["us_per_frame_12_in_flight"], d["us_per_frame_1_in_flight"], d["kernel_us"]["raster"])'
for scene in hall rand; do
  for ctas in 2 3 4 5 6 8; do
    SRB_RASTER_CTAS_PER_SM=$ctas python profiles/ab.py $scene 256 12 2>&1 | tail -1 | python -c "$show" "ctas $ctas"
  done
  for mu in 64 96 192; do
    SRB_MIN_UNIT=$mu python profiles/ab.py $scene 256 12 2>&1 | tail -1 | python -c "$show" "minunit $mu"
  done
done
