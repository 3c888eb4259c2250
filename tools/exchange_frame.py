"""The multi-GPU exchange kernels on ONE GPU (the short command ncu wraps): 2^24 uniform cubes are encoded, counted and
scattered into g buckets whose destinations are slices of local buffers (peers' symmetric memory in the real path), then
the raw pairs of the scene's scan go through the pair exchange the same way.  Usage: python tools/exchange_frame.py [g] [log2 n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import _loadpkg

bp = _loadpkg.load()
from broadphase_rs_b200.dist import _view

g = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
sc = bp.scenes.uniform_cubes(n, 6)
db = torch.from_numpy(sc["bounds"]).cuda()
di = torch.from_numpy(sc["ids"].view(np.int32)).cuda()
L = bp.Layer(2, "u32")
L.set_stream(torch.cuda.current_stream().cuda_stream)   # so that torch's events bracket the layer's kernels
dev = torch.device("cuda")
for rep in range(2):
    L.clear()
    L.extend_device(sc["sys_bounds"], db, di, n)
    kp, ip, r, _ = L.records_device()
    keys, ids = _view(kp, r, torch.int64, dev), _view(ip, r, torch.int32, dev)
    sample = keys[:: max(1, r // 4096)].cpu().numpy().view(np.uint64)
    spl = np.sort(sample)[[(i * sample.shape[0]) // g for i in range(1, g)]].astype(np.uint64)
    counts, halo = L.count_records(keys, r, spl)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ok = torch.empty(r + 16, dtype=torch.int64, device=dev)
    oi = torch.empty(r + 16, dtype=torch.int32, device=dev)
    L.scatter_records(keys, ids, r, spl, [ok.data_ptr() + 8 * int(off[b]) for b in range(g)],
                      [oi.data_ptr() + 4 * int(off[b]) for b in range(g)], None, None, fold_cell_flags=True)
    S = bp.Layer(2, "u32")
    S.set_records(ok, oi, sorted_=False, on_device=True, n=r, flagged=True)
    S.sort()
    ptr, nraw = S.scan_raw_device(None)
    raw = _view(ptr, nraw, torch.int64, dev)
    a = ((raw[:: max(1, nraw // 4096)] >> 32) & 0xFFFFFFFF).cpu().numpy().astype(np.uint64)
    aspl = np.sort(a)[[(i * a.shape[0]) // g for i in range(1, g)]].astype(np.uint64)
    pc = S.count_pairs(raw, nraw, aspl)
    poff = np.concatenate([[0], np.cumsum(pc)]).astype(np.int64)
    op = torch.empty(nraw + 16, dtype=torch.int64, device=dev)
    S.scatter_pairs(raw, nraw, aspl, [op.data_ptr() + 8 * int(poff[b]) for b in range(g)])
    torch.cuda.synchronize()
print("records %d raw pairs %d buckets %d; launches: enc layer %d, shard layer %d" % (r, nraw, g, L.stats()["launches_total"], S.stats()["launches_total"]))

# ---- cost of counting inside the encode (cached splitters): encode alone vs encode + per-shard counts ----
rows = torch.zeros((2, 2 * g + 7), dtype=torch.int64, device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def plain():
    L.clear()
    L.extend_device(sc["sys_bounds"], db, di, n)
    len(L)


def counted():
    L.extend_count_rows(sc["sys_bounds"], db, di, n, spl, True, [rows[0].data_ptr(), rows[1].data_ptr()])
    len(L)


print("encode %.3f ms, encode + counts for %d shards %.3f ms (%d objects)" % (timed(plain), g, timed(counted), n))

# ---- the record exchange itself, all destinations local ----
plain()
kp, ip, r, _ = L.records_device()
keys, ids = _view(kp, r, torch.int64, dev), _view(ip, r, torch.int32, dev)
dk = [ok.data_ptr() + 8 * int(off[b]) for b in range(g)]
dv = [oi.data_ptr() + 4 * int(off[b]) for b in range(g)]
t = timed(lambda: L.scatter_records(keys, ids, r, spl, dk, dv, None, None, fold_cell_flags=True))
print("record exchange, %d local destinations: %.3f ms for %d records = %.0f GB/s of algorithmic bytes (BP_EXCHANGE_TMA=%s)"
      % (g, t, r, 24.0 * r / (t * 1e-3) / 1e9, os.environ.get("BP_EXCHANGE_TMA", "1")))
