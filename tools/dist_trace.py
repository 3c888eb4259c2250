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
# the result of the last frame against the bench's own record of this scene (profiles/*_bench_n<world>.json, parity.timed_shape):
# pair count, order-sensitive hash of the concatenated slices, strict increase inside and across the slices
pairs = dl.frame(sc["sys_bounds"], db, di, n, None)
cnt = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=db.device)
cnts = [torch.empty_like(cnt) for _ in range(world)]
dist.all_gather(cnts, cnt)
cnts = [int(c.item()) for c in cnts]
h = torch.tensor([dist_bench._i64(dist_bench.pair_hash(pairs, sum(cnts[:rank])))], dtype=torch.int64, device=db.device)
edge = torch.zeros(4, dtype=torch.int64, device=db.device)
if pairs.shape[0]:
    edge[:2] = pairs[0].to(torch.int64) & 0xFFFFFFFF
    edge[2:] = pairs[-1].to(torch.int64) & 0xFFFFFFFF
edges = [torch.empty_like(edge) for _ in range(world)]
dist.all_gather(edges, edge)
inc, prev = dist_bench._strictly_increasing(pairs), None
for r in range(world):
    if cnts[r]:
        e = [int(x) for x in edges[r].tolist()]
        inc = inc and (prev is None or prev < (e[0], e[1]))
        prev = (e[2], e[3])
incf = torch.tensor([1 if inc else 0], dtype=torch.int64, device=db.device)
dist.all_reduce(incf, op=dist.ReduceOp.MIN)
dist.all_reduce(h, op=dist.ReduceOp.SUM)
if rank == 0:
    print("check: objects=%d pairs=%d hash=%016x sorted_unique_global=%s" % (n * world, sum(cnts), int(h.item()) & ((1 << 64) - 1), bool(incf.item())))
for r in range(world):
    if rank == r:
        print("rank %d objects/gpu=%d world=%d  total %.3f ms" % (rank, n, world, sum(acc.values())))
        print("   " + "  ".join("%s=%.3f" % (k, v) for k, v in acc.items()))
        print("   records local=%d owned=%d halo=%d raw=%d pairs=%d" % (dl.last["records_local"], dl.last["records_owned"], dl.last["halo"], dl.last["raw_pairs"], dl.last["pairs"]))
        st = dl.layers()[1].stats()
        print("   shard sort passes=%d pair sort passes=%d" % (st["sort_passes"], st["pair_sort_passes"]))
        sys.stdout.flush()
    dist.barrier()
dist.barrier()
dist.destroy_process_group()
