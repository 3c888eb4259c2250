import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import _loadpkg, bench
bp = _loadpkg.load()
sc = bench.make_scene(bp, "cfg2", None)
n = sc["bounds"].shape[0]
d_bounds = torch.from_numpy(sc["bounds"]).cuda(); d_ids = torch.from_numpy(sc["ids"].view(np.int32)).cuda()
layer = bp.LayerBuilder().with_min_depth(sc["min_depth"]).build(sc["kind"], "u32")
for f in range(3):
    bench.gpu_frame_device(layer, sc, d_bounds, d_ids, n, None)
torch.cuda.synchronize()
layer.set_profiling(True)
layer.reset_stats()
bench.gpu_frame_device(layer, sc, d_bounds, d_ids, n, None)
torch.cuda.synchronize()
layer.stats()
layer.set_profiling(False)
