"""Per-phase wall times of the distributed frame (tuning aid).  Launch with torchrun:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_trace.py [log2_objects_per_gpu]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import _loadpkg

bp = _loadpkg.load()
from broadphase_rs_b200 import dist as bpd
from broadphase_rs_b200 import dist_bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
sc = dist_bench._scene_slice(bp, n, world, rank, 6)
from broadphase_rs_b200 import _lib
dl = dist_bench._context(bp, bpd, 2, local, n)
dl.set_stream(torch.cuda.current_stream(local).cuda_stream)
dl.set_option(_lib.DIST_OPT_TRACE, 1)   # bp_dist_frame synchronises and stamps every phase
db = torch.from_numpy(sc["bounds"]).cuda()
di = torch.from_numpy(sc["ids"].view(np.int32)).cuda()
acc = {}
for it in range(8):
    dl.frame(sc["sys_bounds"], db, di, n, None)
    if it >= 3:
        for k, v in dl.last["phases_ms"].items():
            acc[k] = acc.get(k, 0.0) + v / 5
for r in range(world):
    if rank == r:
        print("rank %d objects/gpu=%d world=%d  total %.3f ms" % (rank, n, world, sum(acc.values())))
        print("   " + "  ".join("%s=%.3f" % (k, v) for k, v in acc.items()))
        print("   records local=%d owned=%d halo=%d raw=%d pairs=%d" % (dl.last["records_local"], dl.last["records_owned"], dl.last["halo"], dl.last["raw_pairs"], dl.last["pairs"]))
        sys.stdout.flush()
    dist.barrier()
dist.barrier()
dist.destroy_process_group()
