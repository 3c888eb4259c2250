"""Executable model of record_finish_kernel / pair_finish_kernel's ordering rule (broadphase-rs_b200/csrc/bp_radix.cuh),
for the CPU tests: the algorithm is checked here on small windows, the CUDA kernel against the oracle in the gpu tests.

Input: records already ordered (stably) by key >> gshift.  Every record looks at most `halo` records to either side --
what a tile of the kernel holds around it -- and computes its output position
    group start + #(records of its group that sort before it)      ("before": smaller key, or equal key and earlier).
A group of more than `halo` records is "big": each of its records can tell (it misses an end of the group, or sees both
and counts more than `halo`), stays where it is, and raises the flag.
"""
import numpy as np


def record_finish(keys, gshift, halo, tile):
    keys = np.asarray(keys, dtype=np.uint64)
    n = keys.shape[0]
    out = np.empty(n, dtype=np.int64)
    big_any = False
    g = keys >> np.uint64(gshift)
    for t0 in range(0, n, tile):
        w0 = max(t0 - halo, 0)
        w1 = min(n, t0 + tile + halo)
        for i in range(t0, min(t0 + tile, n)):
            h, pos = i, 0
            while h > w0 and i - h <= halo:
                if g[h - 1] != g[i]:
                    break
                h -= 1
                pos += int(keys[h] <= keys[i])
            e = i + 1
            while e < w1 and e - i <= halo:
                if g[e] != g[i]:
                    break
                pos += int(keys[e] < keys[i])
                e += 1
            closed_l = (w0 == 0) if h == w0 else g[h - 1] != g[i]
            closed_r = (w1 == n) if e == w1 else g[e] != g[i]
            big = (not closed_l) or (not closed_r) or (e - h > halo)
            big_any |= big
            out[i] = i if big else h + pos
    return out, big_any
