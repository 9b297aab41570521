"""Multi-GPU parity: the 4K screen-tile split (BASELINE config 4) composited over NVLink must equal the single-GPU frame bit
for bit.  Needs >= 2 GPUs in one box (skipped otherwise): `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multigpu.py -m gpu`."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import ctypes

        cuda = ctypes.CDLL("libcuda.so.1")
        if cuda.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def _run_split(world, width, height, frames, port, detail=1.0):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_tile_split.py"), "--width", str(width), "--height",
           str(height), "--frames", str(frames), "--detail", str(detail)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert lines, p.stdout[-2000:] + p.stderr[-2000:]
    return json.loads(lines[-1])


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs in one box")
def test_tile_split_two_gpus_composite_is_bit_exact():
    out = _run_split(2, 3840, 2160, 20, 29531)
    assert out["n_gpus"] == 2 and out["composite_bit_exact_vs_single_gpu"] == [True, True], out
    assert out["depth_of_owned_tiles_bit_exact_on_every_rank"] is True, out
    # every rank sets up only the triangles that touch its tiles
    assert all(r["tris_setup"] > 0 for r in out["per_rank"])


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 4, reason="needs four GPUs in one box")
def test_tile_split_all_gpus_composite_is_bit_exact():
    n = 8 if _gpu_count() >= 8 else 4
    out = _run_split(n, 3840, 2160, 20, 29532)
    assert out["n_gpus"] == n and out["composite_bit_exact_vs_single_gpu"] == [True, True], out
    assert out["depth_of_owned_tiles_bit_exact_on_every_rank"] is True, out
    # ragged size: the tile count is not a multiple of the rank count
    out = _run_split(n, 1000, 600, 5, 29533, detail=0.3)
    assert out["composite_bit_exact_vs_single_gpu"] == [True, True], out
