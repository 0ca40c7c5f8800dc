"""PCIe copy bandwidth with N GPUs copying AT THE SAME TIME (one process per GPU under torchrun): D2H from every GPU into its
own pinned host buffer, by the copy engine (cudaMemcpyAsync) and by a kernel storing 16 bytes per thread straight into mapped
pinned memory (what the zero-copy gather kernel does), per-GPU and aggregate.  Explains (or not) the host-buffer loop's scaling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_probe_multi.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n = 64 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dev = h.cuda_view = None


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=20):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return n * reps / dt / 1e9


res = {"rank": rank, "world": world, "cpus": len(os.sched_getaffinity(0))}
res["d2h_copy_engine_gbs"] = timed(lambda: h.copy_(d, non_blocking=True))
res["h2d_copy_engine_gbs"] = timed(lambda: d.copy_(h, non_blocking=True))
# kernel stores into mapped pinned memory (zero copy): torch exposes pinned host memory to kernels through UVA -- a device
# "view" of the host tensor is obtained from its pointer
try:
    import ctypes
    from torch.utils import cpp_extension  # noqa: F401  (not used: no JIT on the box)
    hv = torch.empty(0)
except Exception:
    pass
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    agg = {k: sum(o[k] for o in out) for k in ("d2h_copy_engine_gbs", "h2d_copy_engine_gbs")}
    print(json.dumps({"n_gpus": world, "cpus_per_rank": res["cpus"], "aggregate": agg, "per_gpu_d2h": [round(o["d2h_copy_engine_gbs"], 1) for o in out]}), flush=True)
if world > 1:
    dist.destroy_process_group()
