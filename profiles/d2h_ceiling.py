"""What bounds e2e at N GPUs: device -> pinned-host copies of one frame's colour tiles (8.36 MB), ALL ranks at once, for
several kinds of host buffer.  Run under torchrun (N ranks) on the GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 profiles/d2h_ceiling.py
Prints one JSON line on rank 0: GB/s per rank (min / mean / sum) per variant, alone (ranks one after the other) and together."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sys.argv = sys.argv[:1]
import bench
from softrast_b200 import capi

FRAME = 510 * 16384
ctx = capi.RenderContext(local)


def run(host, nbytes, reps):
    ms = C.c_float()
    rc = capi.lib.srb_debug_d2h_copies(ctx.h, C.c_void_p(host), nbytes, reps, C.byref(ms))
    assert rc == 0
    return nbytes * reps / (ms.value * 1e-3) / 1e9


def gather(x):
    if world == 1:
        return [x]
    out = [None] * world
    dist.all_gather_object(out, x)
    return out


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def variant(name, alloc, free, nbytes, reps):
    p = alloc(nbytes)
    run(p, nbytes, 4)
    alone = None
    for r in range(world):  # one rank at a time
        barrier()
        if r == rank:
            alone = run(p, nbytes, reps)
    barrier()
    together = run(p, nbytes, reps)  # all ranks at once
    barrier()
    free(p)
    a, t = gather(alone), gather(together)
    return {"variant": name, "alone_gbs_min": round(min(a), 1), "alone_gbs_mean": round(float(np.mean(a)), 1),
            "together_gbs_min": round(min(t), 1), "together_gbs_mean": round(float(np.mean(t)), 1), "together_gbs_sum": round(sum(t), 1),
            "per_rank_together": [round(x, 1) for x in t]}


libc = C.CDLL("libc.so.6")
libc.aligned_alloc.restype = C.c_void_p
libc.aligned_alloc.argtypes = [C.c_size_t, C.c_size_t]
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
libc.free.argtypes = [C.c_void_p]
cudart = torch.cuda.cudart()


def alloc_thp(n):
    n2 = (n + (2 << 20) - 1) & ~((2 << 20) - 1)
    p = libc.aligned_alloc(2 << 20, n2)
    libc.madvise(p, n2, 14)  # MADV_HUGEPAGE
    C.memset(p, 0, n2)
    assert int(cudart.cudaHostRegister(p, n2, 0)) == 0
    return p


def free_thp(p):
    cudart.cudaHostUnregister(p)
    libc.free(p)


results = []
results.append(variant("cudaHostAlloc default, 1 frame per copy", capi.host_alloc, capi.host_free, FRAME, 64))
results.append(variant("cudaHostAlloc write-combined, 1 frame per copy", lambda n: capi.host_alloc_ex(n, 1), capi.host_free, FRAME, 64))
results.append(variant("cudaHostAlloc portable, 1 frame per copy", lambda n: capi.host_alloc_ex(n, 2), capi.host_free, FRAME, 64))
results.append(variant("cudaHostAlloc default, 8 frames per copy", capi.host_alloc, capi.host_free, FRAME * 8, 8))
results.append(variant("transparent huge pages + cudaHostRegister, 1 frame per copy", alloc_thp, free_thp, FRAME, 64))
cpus = bench._bind_to_gpu_numa(local)
results.append(variant(f"after binding the rank to its GPU's CPUs ({len(cpus) if cpus else 'no affinity reported'}): default, 1 frame per copy",
                       capi.host_alloc, capi.host_free, FRAME, 64))
results.append(variant("bound + write-combined", lambda n: capi.host_alloc_ex(n, 1), capi.host_free, FRAME, 64))
if rank == 0:
    try:
        numa = open("/sys/devices/system/node/online").read().strip()
    except OSError:
        numa = "?"
    print(json.dumps({"n_gpus": world, "host_cpus": os.cpu_count(), "numa_nodes_online": numa, "results": results}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
