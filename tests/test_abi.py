"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bp.h declares, fails loudly without a GPU, and the host-side radix planner behaves."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(bp):
    declared = _declared_symbols()
    assert len(declared) >= 20
    L = bp.lib()
    for name in declared:
        assert hasattr(L, name), "libbroadphase_b200.so does not export %s" % name
    # and the ctypes table binds exactly the declared set
    from broadphase_rs_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared


def test_version_and_status_strings(bp):
    L = bp.lib()
    assert L.bp_version() == 100
    assert L.bp_status_string(0) == b"ok"
    assert L.bp_status_string(2) == b"CUDA error"


def test_no_cpu_fallback(bp):
    """Without a CUDA device layer creation must fail with BP_ERR_CUDA (never compute on the host)."""
    if bp.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(bp.BpError) as e:
        bp.Layer(bp.Index64_3D, "u32")
    assert e.value.status == 2


def test_dist_context_arguments(bp):
    """bp_dist_create validates its configuration before it touches a device; without one it fails like bp_layer_create."""
    from broadphase_rs_b200._lib import DistConfig
    L = bp.lib()
    h = ctypes.c_void_p()
    assert L.bp_dist_handle_bytes() >= 64                                    # a CUDA IPC handle + the layout check
    assert L.bp_dist_create(None, ctypes.byref(h)) == 1
    for cfg in (DistConfig(2, 0, -1, 0, 0, 1 << 10, 1 << 10),                # world 0
                DistConfig(2, 0, -1, 0, 17, 1 << 10, 1 << 10),               # more than 16 ranks
                DistConfig(2, 0, -1, 3, 2, 1 << 10, 1 << 10),                # rank outside the world
                DistConfig(0, 0, -1, 0, 1, 1 << 10, 1 << 10)):               # Index32_2D: the sharded frame is for the 64-bit indices
        assert L.bp_dist_create(ctypes.byref(cfg), ctypes.byref(h)) == 1 and not h.value
    if bp.device_count() == 0:
        cfg = DistConfig(2, 0, -1, 0, 1, 1 << 10, 1 << 10)
        assert L.bp_dist_create(ctypes.byref(cfg), ctypes.byref(h)) == 2     # BP_ERR_CUDA: no CPU fallback
    assert L.bp_dist_destroy(None) == 0
    assert L.bp_dist_frame(None, None, None, None, 0, None, None, None) == 1
    assert L.bp_dist_layer(None, 0) is None


def test_invalid_arguments(bp):
    from broadphase_rs_b200._lib import LayerConfig
    L = bp.lib()
    h = ctypes.c_void_p()
    assert L.bp_layer_create(None, ctypes.byref(h)) == 1
    cfg = LayerConfig(7, 4, 0, -1, 0, 0, 0)
    assert L.bp_layer_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    cfg = LayerConfig(2, 3, 0, -1, 0, 0, 0)
    assert L.bp_layer_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    assert L.bp_layer_destroy(None) == 0


def _covered(plan):
    c = 0
    for f in plan:
        c |= ((1 << f[1]) - 1) << f[0]
        if len(f) == 4:
            c |= ((1 << f[3]) - 1) << f[2]
    return c


def test_radix_planner(bp):
    plan = bp.plan_radix_passes
    assert plan(0) == []
    assert plan(0xFF) == [(0, 8)]
    assert plan(0x1) == [(0, 1)]
    # Index64_3D, every record at depth 7: only the top 21 origin bits vary -> 3 passes, not 8
    mask = ((1 << 21) - 1) << (5 + 57 - 21)
    p = plan(mask)
    assert len(p) == 3 and p[0][0] == 41 and _covered(p) & mask == mask
    # multi-depth keys: 4 depth bits + 39 origin bits with a gap in between: 43 bits -> 6 passes, the
    # first digit made of two bit-fields (depth bits + the lowest origin bits)
    mask = 0xF | (((1 << 39) - 1) << 23)
    p = plan(mask)
    assert len(p) == 6 and p[0] == (0, 4, 23, 4)
    assert _covered(p) & mask == mask
    # packed ID pairs with 20-bit IDs: 40 bits in two fields -> 5 passes
    mask = 0xFFFFF | (0xFFFFF << 32)
    p = plan(mask)
    assert len(p) == 5 and _covered(p) & mask == mask
    # scattered bits: the single 8-bit window wins over two 1-bit runs
    p = plan(0b10101010)
    assert p == [(1, 7)]
    # digits never overlap, are processed from the least significant bit up, and hold <= 8 bits
    for mask in (0xDEADBEEFCAFEF00D, (1 << 62) - 1, (1 << 64) - 1, 0x8000000000000001):
        p = plan(mask)
        assert _covered(p) & mask == mask
        seen, last_top = 0, -1
        for f in p:
            fields = [(f[0], f[1])] + ([(f[2], f[3])] if len(f) == 4 else [])
            assert sum(b for _, b in fields) <= 8
            for sft, b in fields:
                m = ((1 << b) - 1) << sft
                assert m & seen == 0 and sft + b <= 64 and sft > last_top
                seen |= m
                last_top = sft + b - 1
    assert len(plan((1 << 62) - 1)) == 8
    assert len(plan((1 << 64) - 1)) == 8


def test_scene_recipes_are_deterministic(bp):
    a = bp.scenes.uniform_cubes(1000, 2)
    b = bp.scenes.uniform_cubes(1000, 2)
    assert (a["bounds"] == b["bounds"]).all() and a["bounds"].dtype.name == "float32"
    c = bp.scenes.lognormal_cubes(1000, 3)
    assert (c["bounds"][:, 3:] >= c["bounds"][:, :3]).all() and (c["bounds"][:, 3:] <= 1.0).all()
    d = bp.scenes.example_circles(1000, 1)
    assert d["kind"] == 0 and d["min_depth"] == 4 and d["bounds"].shape == (1000, 4)


def test_sort_finish_planner(bp):
    """Which record sorts take "radix passes over the top bits + one finish pass" (bp_plan_sort_finish)."""
    plan = bp.plan_sort_finish
    origin = lambda nbits: ((1 << nbits) - 1) << (5 + 57 - nbits)     # the top origin bits of an Index64_3D key
    # config 3: 4 depth bits + 39 origin bits, 87 M records -> 24 top bits in 3 passes (+ finish) instead of 6 passes
    mask = 0xF | origin(39)
    top, gs = plan(mask, 87_322_146)
    assert top == origin(24) and gs == 62 - 24 and len(bp.plan_radix_passes(top)) == 3
    # the same keys, 2.7 M records (the test-sized config 3): 19 bits wanted, whole passes -> again 24 bits
    assert plan(mask, 2_700_000) == (origin(24), 38)
    # config-5 shape: 27 varying bits, 146 M records want 25 bits -> every pass is needed, plain plan
    assert plan(origin(27), 146_536_534) is None
    # config 2: 21 bits in 3 passes: nothing to save
    assert plan(origin(21), 3_554_446) is None
    # all 62 bits, 1 M records: 17 bits wanted -> 3 passes (24 bits) instead of 8
    top, gs = plan((1 << 62) - 1, 1 << 20)
    assert top == ((1 << 24) - 1) << 38 and gs == 38
    # a plan that would save a single pass is not worth the finish pass
    assert plan(origin(32), 1 << 20) is None      # 4 passes against 3 + finish
    assert plan(origin(33), 1 << 20) is not None  # 5 passes against 3 + finish
    # scattered varying bits: the group shift is the lowest of the bits sorted on
    mask = 0x8000_0000_0000_0000 | (0xFFFF_FFFF << 20) | 0xFF
    top, gs = plan(mask, 1 << 12)                  # 9 bits wanted -> 16 bits
    assert bin(top).count("1") == 16 and top & mask == top and top >> gs << gs == top and (mask >> gs << gs) == top
    assert plan(0, 100) is None and plan(0xFF, 1) is None and plan((1 << 62) - 1, 0) is None


def test_finish_model_orders_groups():
    """The ordering rule of the finish kernels on windows small enough to hit every edge: tile boundaries inside groups,
    groups at the array ends, groups of exactly / just over the window -- tests/finish_model.py."""
    from tests.finish_model import record_finish
    rng = np.random.Generator(np.random.Philox(5))
    for trial in range(60):
        n = int(rng.integers(1, 400))
        gshift = int(rng.integers(2, 7))
        halo, tile = int(rng.integers(2, 12)), int(rng.integers(4, 40))
        ngroups = max(1, int(rng.integers(1, max(2, n // max(1, int(rng.integers(1, 14)))))))
        keys = (rng.integers(0, ngroups, n).astype(np.uint64) << np.uint64(gshift)) | rng.integers(0, 1 << gshift, n).astype(np.uint64)
        order = np.argsort(keys >> np.uint64(gshift), kind="stable")           # what the radix passes leave behind
        pre = keys[order]
        out, big = record_finish(pre, gshift, halo, tile)
        assert sorted(out.tolist()) == list(range(n))                             # a permutation, big groups or not
        res = np.empty(n, dtype=np.uint64)
        src = np.empty(n, dtype=np.int64)
        res[out] = pre
        src[out] = order
        want = np.argsort(keys, kind="stable")
        sizes = np.unique(pre >> np.uint64(gshift), return_counts=True)[1]
        assert big == bool((sizes > halo).any())
        if not big:
            assert (src == want).all()                                            # = the stable sort by the whole key
        else:   # big groups stay as they were: still stable, the remaining passes finish the job
            assert (src[np.argsort(res, kind="stable")] == want).all()
