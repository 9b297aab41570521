# usage: bash profiles/run_scale.sh <tag> <N> [ref]   : bench.py at N GPUs (torchrun), optionally the reference arm first
cd $GRAFT_REPO_ROOT
tag=$1; n=$2
mkdir -p gpurun_out
if [ "$3" = "ref" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference_${n}gpu.json 2> gpurun_out/${tag}_ref_${n}gpu.err; echo "ref rc=$?"; fi
SECONDS=0
if [ "$n" = "1" ]; then timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err; fi
echo "bench rc=$? wall ${SECONDS}s"; tail -2 gpurun_out/${tag}_bench_${n}gpu.err
if [ "$n" != "1" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29563 tests/multigpu_tile_split.py --frames 200 2>&1 | tail -1 > gpurun_out/${tag}_tile_split_${n}gpu.json; cut -c1-330 gpurun_out/${tag}_tile_split_${n}gpu.json; fi
