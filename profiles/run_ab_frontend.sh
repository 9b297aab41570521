set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" ; tail -5 gpurun_out/r02a_pytest.log
for cfg in "default" "SRB_NO_GRAPH=1" "SRB_NO_GRAPH=1 SRB_SEPARATE_SCAN=1"; do
  echo "== $cfg"
  env $( [ "$cfg" = default ] || echo $cfg ) timeout 300 python profiles/sweep.py hall 1,12 512 2>&1 | tail -6
done
