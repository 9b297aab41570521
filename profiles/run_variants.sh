#!/bin/bash
# A/B of experiment builds (make -C softrast_b200/csrc VARIANT=name DEFS="-D...") against the default library.
# usage: run_variants.sh "name1 name2 ..." "scene1 scene2 ..." [full]
cd "$GRAFT_REPO_ROOT" || exit 1
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], d["scene"], d["us_per_frame_12_in_flight"], d["us_per_frame_1_in_flight"], d["kernel_us"])'
for rep in 1 2; do
  for scene in $2; do
    python profiles/ab.py $scene 256 12 2>&1 | tail -1 | python -c "$show" "default"
    for v in $1; do
      SRB_LIB=$PWD/softrast_b200/lib/libsoftrast_b200_$v.so python profiles/ab.py $scene 256 12 2>&1 | tail -1 | python -c "$show" "$v"
    done
  done
done
if [ "$3" = full ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
