# usage: bash profiles/run_bench.sh <tag>   : gpu test suite, reference arm, bench (both with the driver's flags)
cd $GRAFT_REPO_ROOT
tag=$1
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log; fi
SECONDS=0; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"; echo "ref wall ${SECONDS}s"
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; echo "bench wall ${SECONDS}s"; tail -3 gpurun_out/${tag}_bench.err
