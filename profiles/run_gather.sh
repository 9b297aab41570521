#!/bin/bash
# geometry re-uploaded every frame (SRB_FLAG_UPLOAD_ALWAYS): gather by loads (default) against bulk asynchronous copies (SRB_GATHER_BULK=1)
# and against one cudaMemcpyBatchAsync per frame (SRB_GATHER_BATCH=1)
cd "$GRAFT_REPO_ROOT" || exit 1
show='import sys,json; d=json.loads([l for l in sys.stdin.read().strip().split("\n") if l.startswith("{")][-1]); g=d["e2e_geometry_upload"]; print(sys.argv[1], "e2e+geometry", round(g["value"]), "frames/s; upload alone", round(g["without_readback"]), "frames/s =", round(g["h2d_gbs_without_readback"],1), "GB/s; e2e", round(d["e2e"]["value"]))'
for rep in 1 2; do
  timeout 300 python bench.py --no-configs --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "$show" "loads"
  SRB_GATHER_BULK=1 timeout 300 python bench.py --no-configs --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "$show" "bulk "
  SRB_GATHER_BATCH=1 timeout 300 python bench.py --no-configs --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "$show" "batch"
done
SRB_GATHER_BATCH=1 timeout 600 python -m pytest tests/test_gpu_shim.py tests/test_gpu_obj.py tests/test_gpu_parity.py -m gpu -x -q -k "shim or upload or always or pinned or obj" 2>&1 | tail -3
