"""A small frame of every index kind / ID width / filter / merge path, for compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import _loadpkg

bp = _loadpkg.load()
rng = np.random.Generator(np.random.Philox(1))
for kind, dim in ((0, 2), (1, 2), (2, 3)):
    for idt in ("u32", "u64"):
        n = 3000
        sysb = np.concatenate([np.zeros(dim), np.ones(dim)]).astype(np.float32)
        size = (0.05 * rng.random((n, dim)) ** 2).astype(np.float32)
        mn = (rng.random((n, dim)) * (1 - size)).astype(np.float32)
        b = np.concatenate([mn, mn + size], axis=1).astype(np.float32)
        ids = rng.integers(0, n // 2, size=n).astype(np.uint32 if idt == "u32" else np.uint64)
        s = bp.LayerBuilder().with_min_depth(2).build(kind, idt)
        s.extend(sysb, b[:2000], np.sort(ids[:2000]))
        s.sort()
        d = bp.LayerBuilder().with_min_depth(3).build(kind, idt)
        d.extend(sysb, b[2000:], ids[2000:])
        p0 = d.scan_filtered(bp.ScanFilter.id_parity()).shape[0]
        d.merge(s)
        p1 = d.scan().shape[0]
        d.extend(sysb, b[:100], ids[:100])
        p2 = d.scan_filtered(bp.ScanFilter.xor_mask(6)).shape[0]
        k, i = d.iter()
        print(kind, idt, len(d), p0, p1, p2, k.shape)
print("done")
