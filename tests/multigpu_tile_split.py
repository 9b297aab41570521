"""BASELINE config 4: one 3840x2160 frame of the hall scene, screen-tile split across the GPUs of one box with the
composite done by peer stores over NVLink (each rank's shade kernel writes its tiles into rank 0's framebuffer).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_tile_split.py [--width 3840 --height 2160 --frames 50]

Checks the composited frame bit-for-bit against a single-GPU render of the same frame, then times frames (every rank
renders its tiles of frame f, device sync, barrier) and prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--frames", type=int, default=50)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from softrast_b200 import capi, scenes

    scene = scenes.hall_scene(args.width, args.height)
    blob = [None]
    if rank == 0:
        r = capi.SceneRenderer(scene, device=local)
        blob[0] = r.ctx.export_framebuffer(r.fb)
    if world > 1:
        dist.broadcast_object_list(blob, src=0)
    if rank != 0:
        r = capi.SceneRenderer(scene, device=local, fb_import=blob[0])
    r.ctx.set_tile_ownership(world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    r.render()
    barrier()
    ok = None
    if rank == 0:
        colour, depth = r.read_tiles()
        full = capi.SceneRenderer(scene, device=local)
        full.render()
        fc, fd = full.read_tiles()
        ok = bool(np.array_equal(colour, fc) and np.array_equal(depth.view(np.uint32), fd.view(np.uint32)))
        # single-GPU timing of the same frame for the speed-up
        for _ in range(5):
            full.render()
        t0 = time.perf_counter()
        for _ in range(args.frames):
            full.render()
        single_ms = (time.perf_counter() - t0) / args.frames * 1e3
        full.close()
    barrier()
    for _ in range(5):
        r.render()
        barrier()
    t0 = time.perf_counter()
    for _ in range(args.frames):
        r.render()
        barrier()
    split_ms = (time.perf_counter() - t0) / args.frames * 1e3
    # per-rank kernel times (CUDA events on the library's stream) of a few more frames
    r.ctx.set_timing(True)
    acc = {}
    for _ in range(8):
        r.render()
        barrier()
        for k, v in r.ctx.kernel_times().items():
            acc[k] = acc.get(k, 0.0) + v / 8
    r.ctx.set_timing(False)
    per_rank = [None] * world
    if world > 1:
        dist.all_gather_object(per_rank, {k: round(v, 1) for k, v in acc.items()})
    else:
        per_rank = [{k: round(v, 1) for k, v in acc.items()}]
    if rank == 0:
        print(json.dumps({"config": f"hall {args.width}x{args.height} screen-tile split, NVLink composite by peer stores",
                          "n_gpus": world, "composite_bit_exact_vs_single_gpu": ok,
                          "ms_per_frame_split": split_ms, "ms_per_frame_single_gpu": single_ms,
                          "tiles": r.fb.num_tiles, "kernel_us_per_rank": per_rank, "counters": r.ctx.counters()}), flush=True)
    barrier()
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
