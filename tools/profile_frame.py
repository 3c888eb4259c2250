"""Runs a few frames of one workload on cuda:0 -- the short command that ncu wraps (see
/opt/skills/guides/B200_PROFILING.md).  Usage: python tools/profile_frame.py cfg3 [n] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import _loadpkg
import bench

bp = _loadpkg.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = bench.make_scene(bp, wl, n)
n = sc["bounds"].shape[0]
flt = bench.scene_filter(bp, wl)
d_bounds = torch.from_numpy(sc["bounds"]).cuda()
d_ids = torch.from_numpy(sc["ids"].view(np.int32)).cuda()
layer = bp.LayerBuilder().with_min_depth(sc["min_depth"]).build(sc["kind"], "u32")
for f in range(frames):
    _, pairs = bench.gpu_frame_device(layer, sc, d_bounds, d_ids, n, flt)
torch.cuda.synchronize()
st = layer.stats()
print("frames=%d objects=%d records=%d raw=%d pairs=%d sort_passes=%d pair_passes=%d launches=%d" % (
    frames, n, st["n_records"], st["n_raw_pairs"], pairs, st["sort_passes"], st["pair_sort_passes"], st["launches_total"]))
