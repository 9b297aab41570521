# usage: bash profiles/run_tests_ab.sh <tag> <scene> <frames> "<knobs>"...   : gpu test suite, then A/B of the knob sets
cd $GRAFT_REPO_ROOT
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
bash profiles/run_ab.sh "$@"
