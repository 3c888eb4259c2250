"""Probe: does torch symmetric memory give usable NVLink peer pointers here?  (torchrun, >= 2 GPUs)"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 64 << 20  # int64 elements: 512 MB
    t = symm_mem.empty(n, dtype=torch.int64, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "world", hdl.world_size, flush=True)
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (n,), torch.int64)
    src = torch.full((n,), rank + 1, dtype=torch.int64, device=dev)
    hdl.barrier()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pt.copy_(src)  # a kernel / copy engine writing into the peer's memory over NVLink
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(rank, "peer write %.1f GB/s" % (n * 8 / dt / 1e9), flush=True)
    hdl.barrier()
    torch.cuda.synchronize()
    ok = bool((t == ((rank - 1) % world) + 1).all().item())
    print(rank, "peer data arrived:", ok, flush=True)
    # all_to_all_single for comparison
    a = torch.full((n,), rank, dtype=torch.int64, device=dev)
    b = torch.empty_like(a)
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dist.all_to_all_single(b, a)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(rank, "nccl all_to_all_single: %.1f GB/s leaving the GPU" % (n * 8 * (world - 1) / world / dt / 1e9), flush=True)
except Exception as e:
    import traceback
    traceback.print_exc()
    print(rank, "symmetric memory unavailable:", repr(e), flush=True)
dist.barrier()
dist.destroy_process_group()
