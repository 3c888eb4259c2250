"""Timings of the BASELINE configs that bench.py does not put on its headline line: config 1 (10,000
circles, Index32_2D, min_depth 4) and config 4 (2^26 static objects sorted once + 2^22 dynamic objects
per frame through Layer::merge).  Device-resident inputs, CUDA events, one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import _loadpkg

bp = _loadpkg.load()
out = {}
stream = torch.cuda.current_stream()


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), r


# ---- config 1 ----
sc = bp.scenes.example_circles(10_000, 1)
L = bp.LayerBuilder().with_min_depth(4).build(bp.Index32_2D, "u32")
L.set_stream(stream.cuda_stream)
db = torch.from_numpy(sc["bounds"]).cuda()
di = torch.from_numpy(sc["ids"].view(np.int32)).cuda()


def frame1():
    L.clear()
    L.extend_device(sc["sys_bounds"], db, di, 10_000)
    L.par_sort()
    return L.scan_device(None)[1]


ms, pairs = timed(frame1, 50)
out["cfg1_10k_circles_Index32_2D"] = {"ms_per_frame": ms, "objects_per_s": 10_000 / (ms * 1e-3), "pairs": pairs,
                                     "records": L.stats()["n_records"], "reference_readme_ms": 6.0}

# ---- config 4 ----
ns, nd = (1 << 26), (1 << 22)
if len(sys.argv) > 1:
    ns, nd = 1 << int(sys.argv[1]), 1 << int(sys.argv[2])
st = bp.scenes.uniform_cubes(ns, 4)
S = bp.Layer(bp.Index64_3D, "u32")
S.set_stream(stream.cuda_stream)
sb = torch.from_numpy(st["bounds"]).cuda()
si = torch.from_numpy(st["ids"].view(np.int32)).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
S.extend_device(st["sys_bounds"], sb, si, ns)
S.sort()
len(S)
e1.record(stream)
torch.cuda.synchronize()
static_ms = e0.elapsed_time(e1)
static_records = len(S)
del sb, si, st
dy = bp.scenes.uniform_cubes(nd, 5, id_base=ns, edge_factor=0.4 * (ns / nd) ** (-1.0 / 3.0))
D = bp.Layer(bp.Index64_3D, "u32")
D.set_stream(stream.cuda_stream)
ddb = torch.from_numpy(dy["bounds"]).cuda()
ddi = torch.from_numpy(dy["ids"].view(np.int32)).cuda()


def frame4():
    D.clear()
    D.extend_device(dy["sys_bounds"], ddb, ddi, nd)
    D.sort()
    D.merge(S)
    return D.scan_device(None)[1]


ms, pairs = timed(frame4, 5, 2)
stt = D.stats()
out["cfg4_static_2^%d_dynamic_2^%d" % (int(np.log2(ns)), int(np.log2(nd)))] = {
    "static_build_ms": static_ms, "static_records": static_records, "ms_per_frame": ms,
    "objects_in_scan_per_s": (ns + nd) / (ms * 1e-3), "dynamic_objects_per_s": nd / (ms * 1e-3),
    "records": stt["n_records"], "raw_pairs": stt["n_raw_pairs"], "pairs": pairs, "merged": stt["merged"],
    "sort_passes_dynamic": stt["sort_passes"], "pair_sort_passes": stt["pair_sort_passes"]}
print(json.dumps(out))
