"""BASELINE config 4: one 3840x2160 frame of the hall scene, screen-tile split across the GPUs of one box with the
composite done by peer stores over NVLink (each rank's shade kernel writes its tiles into rank 0's framebuffer) and the
frame's completion signalled on the device: every rank stamps an arrival flag in rank 0's memory, rank 0's shade kernel
waits for all stamps, and a release stamp keeps the other ranks from overwriting a frame rank 0 is still reading.
There is NO host barrier inside the frame loop.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_tile_split.py [--width 3840 --height 2160 --frames 50]

Checks composited frames bit-for-bit against single-GPU renders of the same frames (while the other ranks are already a
frame ahead), then times frames and prints one JSON line on rank 0.  Run by tests/test_gpu_multigpu.py."""
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
    ap.add_argument("--detail", type=float, default=1.0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from softrast_b200 import capi, scenes

    scene = scenes.hall_scene(args.width, args.height, detail=args.detail)
    path = scenes.hall_camera_path(scene, 64)
    blob = [None]
    full = None
    if rank == 0:
        r = capi.SceneRenderer(scene, device=local)
        blob[0] = r.ctx.export_framebuffer(r.fb)
        full = capi.SceneRenderer(scene, device=local)  # the same frames on one GPU, for the comparison
    if world > 1:
        dist.broadcast_object_list(blob, src=0)
    if rank != 0:
        r = capi.SceneRenderer(scene, device=local, fb_import=blob[0])
    r.ctx.set_tile_ownership(world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def same_as_single_gpu(f):
        colour, depth = r.read_tiles()
        full.render(mvps=path[f % len(path)])
        fc, fd = full.read_tiles()
        own = np.arange(rank, colour.shape[0], world)  # depth stays on the GPU that owns the tile
        return bool(np.array_equal(colour, fc) and np.array_equal(depth.view(np.uint32)[own], fd.view(np.uint32)[own]))

    barrier()
    ok = []
    depth_ok = True
    # frames 0..5 back to back with no barrier: rank 0 stops after frame 2 and after frame 5 to look at its framebuffer,
    # the others run ahead as far as the release stamp lets them (one frame)
    for f in range(6):
        r.render(mvps=path[f])
        if rank == 0 and f in (2, 5):
            time.sleep(0.2)  # the other ranks are already waiting inside frame f + 1
            ok.append(same_as_single_gpu(f))
    if rank != 0:
        # every other rank checks the depth of ITS tiles against a full render of its own
        mine = capi.SceneRenderer(scene, device=local)
        mine.render(mvps=path[5])
        _, fd = mine.read_tiles()
        _, d = r.read_tiles()
        own = np.arange(rank, d.shape[0], world)
        depth_ok = bool(np.array_equal(d.view(np.uint32)[own], fd.view(np.uint32)[own]))
        mine.close()
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, depth_ok)
        depth_ok = all(flags)
    barrier()
    single_ms = None
    if rank == 0:
        for f in range(5):
            full.render(mvps=path[f])
        mv1 = path[np.arange(args.frames) % len(path)]
        t0 = time.perf_counter()
        capi.render_frames([full], args.frames, mv1)
        torch.cuda.synchronize()
        single_ms = (time.perf_counter() - t0) / args.frames * 1e3
    barrier()
    for f in range(5):
        r.render(mvps=path[f])
    barrier()
    mv = path[np.arange(args.frames) % len(path)]
    t0 = time.perf_counter()
    capi.render_frames([r], args.frames, mv)  # the C loop: no Python between the frames
    torch.cuda.synchronize()
    split_ms = (time.perf_counter() - t0) / args.frames * 1e3
    if world > 1:
        t = torch.tensor([split_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        split_ms = float(t.item())
    barrier()
    # per-rank kernel times (CUDA events on the library's stream) of a few more frames
    r.ctx.set_timing(True)
    acc = {}
    for f in range(8):
        r.render(mvps=path[f])
        for k, v in r.ctx.kernel_times().items():
            acc[k] = acc.get(k, 0.0) + v / 8
    r.ctx.set_timing(False)
    counters = r.ctx.counters()
    per_rank = [None] * world
    mine = {"kernel_us": {k: round(v, 1) for k, v in acc.items()}, "tris_setup": counters["tris_setup"], "tile_refs": counters["tile_refs"]}
    if world > 1:
        dist.all_gather_object(per_rank, mine)
    else:
        per_rank = [mine]
    if rank == 0:
        print(json.dumps({"config": f"hall {args.width}x{args.height} screen-tile split, NVLink composite by peer stores, device-side completion",
                          "n_gpus": world, "composite_bit_exact_vs_single_gpu": ok, "depth_of_owned_tiles_bit_exact_on_every_rank": depth_ok,
                          "ms_per_frame_split": split_ms, "ms_per_frame_single_gpu": single_ms,
                          "tiles": r.fb.num_tiles, "per_rank": per_rank}), flush=True)
        full.close()
    barrier()
    r.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
