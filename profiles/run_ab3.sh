#!/bin/bash
# ab.py over the bench scenes (twice); "full" as first argument: then the GPU test suite
cd "$GRAFT_REPO_ROOT" || exit 1
show='import sys,json; d=json.loads(sys.stdin.read()); print(d["scene"], d["us_per_frame_12_in_flight"], d["us_per_frame_1_in_flight"], d["kernel_us"])'
for rep in 1 2; do
  for scene in hall rand cubes100 hall4k; do
    python profiles/ab.py $scene 256 12 2>&1 | tail -1 | python -c "$show"
  done
done
[ "$1" = full ] && timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
