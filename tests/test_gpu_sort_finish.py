"""GPU parity tests of the "top bits + finish" record sort and of the pair sort that leaves the lowest later-ID bits to the
finish kernel (csrc/bp_radix.cuh record_finish_kernel, csrc/bp_scan.cuh pair_finish_kernel): the same equalities as
tests/test_gpu_parity.py -- record sequence after sort, pair sequence after scan, bit for bit against the oracle
(src/layer.rs:146-165 and :473-474 sort total orders: any correct sort yields the reference's sequence) -- on scenes built
to take these plans, their window overflow and the fall-back behind it.

BP_SORT_FINISH_MIN (read when a layer is created) lowers the record count from which the plan is considered, so that
scenes the oracle finishes in a second take it."""
import numpy as np
import pytest

from oracle import cpu_oracle as co
from tests.test_gpu_parity import (_assert_pairs_equal, _assert_records_equal, _is_strictly_increasing, _pair, _random_scene)

pytestmark = pytest.mark.gpu


@pytest.fixture
def finish_everywhere(monkeypatch):
    monkeypatch.setenv("BP_SORT_FINISH_MIN", "64")


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("id_bytes", [4, 8])
@pytest.mark.parametrize("shuffle", [False, True])
def test_finish_plan_matches_oracle(bp, finish_everywhere, kind, id_bytes, shuffle):
    """Multi-depth scenes (sizes over three decades): 28 (Index32_2D) to ~60 varying key bits, 50 k records.  With IDs in
    random order the ID digits are sorted first and the finish pass must keep that order among equal keys."""
    sysb, bounds, ids = _random_scene(kind, 30_000, 300 + kind, span=0.05, shuffle_ids=shuffle)
    if id_bytes == 8:
        ids = ids.astype(np.uint64) * np.uint64(0x100000001) + np.uint64(1 << 41)
    g, o = _pair(bp, kind, id_bytes, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    g.par_sort()
    o.par_sort()
    _assert_records_equal(g, o)
    st = g.stats()
    assert st["launches"]["sort_finish"] == 1
    plan = bp.plan_sort_finish(_varying_mask(o), len(o))
    assert plan is not None
    if not shuffle:     # (shuffled IDs add the passes over the ID digits)
        assert st["sort_passes"] == len(bp.plan_radix_passes(plan[0])) <= len(bp.plan_radix_passes(_varying_mask(o))) - 2
    gp, op = g.par_scan(), o.par_scan()
    _assert_pairs_equal(gp, op)
    assert gp.shape[0] > 0 and _is_strictly_increasing(gp)
    _assert_records_equal(g, o)


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_finish_plan_on_a_few_hundred_records(bp, finish_everywhere, kind):
    """~1000 records: 8 top bits in one pass, groups ordered by 50+ low bits -- more than fit the finish kernel's packed
    (low bits, position) words, so the 64-bit kinds take its general form (record_finish_walk_kernel)."""
    sysb, bounds, ids = _random_scene(kind, 500, 400 + kind, span=0.05, shuffle_ids=False)
    g, o = _pair(bp, kind, 4, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    g.par_sort()
    o.par_sort()
    _assert_records_equal(g, o)
    plan = bp.plan_sort_finish(_varying_mask(o), len(o))
    assert plan is not None and (kind == 0 or plan[1] > 52)
    assert g.stats()["launches"]["sort_finish"] == 1
    _assert_pairs_equal(g.par_scan(), o.par_scan())


def _varying_mask(o):
    k, _ = o.records()
    return int(np.bitwise_or.reduce(k)) & ~int(np.bitwise_and.reduce(k)) & ((1 << 64) - 1)


def test_finish_plan_group_sizes_around_the_window(bp, finish_everywhere):
    """Crowds of tiny boxes inside one cell of the top-bit grid, each with its own low bits: groups of a few records up
    to thousands that agree on the sorted top bits.  Up to the 256-record window they are ordered by the finish pass;
    beyond it the pass copies the group through, raises its flag, and the host sorts the long way -- the same sequence
    either way -- and the next sorts of that layer skip the plan for a while."""
    rng = np.random.Generator(np.random.Philox(77))
    sysb = np.array([0, 0, 0, 1, 1, 1], dtype=np.float32)
    bg = bp.scenes.lognormal_cubes(60_000, 31)
    outcomes = set()
    for crowd in (1, 40, 100, 400, 5000):
        # `crowd` tiny boxes (depth ~14-17) inside one cell of side 2^-9: equal top origin bits, different low bits
        base = np.array([0.3, 0.6, 0.2], dtype=np.float32) + np.float32(2.0 ** -12)
        mn = (base + rng.random((crowd, 3)).astype(np.float32) * np.float32(2.0 ** -11)).astype(np.float32)
        size = (np.float32(2.0 ** -17) * (1 + rng.random((crowd, 1)) * 6)).astype(np.float32)
        cb = np.concatenate([mn, mn + size], axis=1).astype(np.float32)
        bounds = np.concatenate([bg["bounds"][:30_000], cb, bg["bounds"][30_000:]]).astype(np.float32)
        ids = np.arange(bounds.shape[0], dtype=np.uint32)
        g, o = _pair(bp, 2, 4, 0)
        g.extend(sysb, bounds, ids)
        o.extend(sysb, bounds, ids)
        g.par_sort()
        o.par_sort()
        _assert_records_equal(g, o)
        st = g.stats()
        assert st["launches"]["sort_finish"] == 1
        plan = bp.plan_sort_finish(_varying_mask(o), len(o))
        assert plan is not None
        top_passes, full_passes = len(bp.plan_radix_passes(plan[0])), len(bp.plan_radix_passes(_varying_mask(o)))
        k, _ = o.records()
        biggest = int(np.unique(k >> np.uint64(plan[1]), return_counts=True)[1].max())
        fallback = biggest > 256
        outcomes.add(fallback)
        assert st["sort_passes"] == (top_passes + full_passes if fallback else top_passes), (crowd, biggest, st["sort_passes"])
        _assert_pairs_equal(g.par_scan(), o.par_scan())
        # again, same layer: after an overflow the plan is not tried (no second finish launch), otherwise it is
        g.clear(); o.clear()
        g.extend(sysb, bounds, ids); o.extend(sysb, bounds, ids)
        g.par_sort(); o.par_sort()
        _assert_records_equal(g, o)
        assert g.stats()["launches"]["sort_finish"] == (1 if fallback else 2)
    assert outcomes == {False, True}


def test_finish_plan_sorts_a_tail_behind_a_sorted_prefix(bp, finish_everywhere):
    """extend onto a sorted tree: the new records are sorted on their own (a sub-range of the buffers) and merged."""
    a = bp.scenes.lognormal_cubes(50_000, 41)
    b = bp.scenes.lognormal_cubes(20_000, 42)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(a["sys_bounds"], a["bounds"], a["ids"]); o.extend(a["sys_bounds"], a["bounds"], a["ids"])
    g.par_sort(); o.par_sort()
    nb, ni = b["bounds"], b["ids"] + np.uint32(50_000)
    g.extend(a["sys_bounds"], nb, ni); o.extend(a["sys_bounds"], nb, ni)
    g.par_sort(); o.par_sort()
    _assert_records_equal(g, o)
    st = g.stats()
    assert st["merged"] == 1 and st["launches"]["sort_finish"] == 2
    _assert_pairs_equal(g.par_scan_filtered(bp.ScanFilter.id_parity()), o.par_scan(co.FILTER_ID_PARITY))


def test_default_threshold_takes_the_plan_at_config3_scale(bp):
    """No environment override: the test-sized config 3 (2^19 log-normal objects, 2.7 M records, 43+ varying bits) sorts
    with 3 radix passes + the finish pass."""
    sc = bp.scenes.lognormal_cubes(1 << 19, 3)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"]); o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    g.par_sort(); o.par_sort()
    _assert_records_equal(g, o)
    st = g.stats()
    assert st["sort_passes"] == 3 and st["launches"]["sort_finish"] == 1
    # small sorts keep the plain plan (no host round trip in their sort)
    g2, o2 = _pair(bp, 2, 4, 0)
    g2.extend(sc["sys_bounds"], sc["bounds"][:20_000], sc["ids"][:20_000]); o2.extend(sc["sys_bounds"], sc["bounds"][:20_000], sc["ids"][:20_000])
    g2.par_sort(); o2.par_sort()
    _assert_records_equal(g2, o2)
    assert g2.stats()["launches"]["sort_finish"] == 0


@pytest.mark.parametrize("n_objects,edge_factor,passes", [(1 << 17, 0.4, 2), (1 << 16, 0.4, 2), (1 << 17, 0.9, 3)])
def test_pair_sort_leaves_low_id_bits_to_the_finish_kernel(bp, n_objects, edge_factor, passes):
    """u32 IDs beyond the counting sort's range: radix passes over the later ID, its lowest bit left out when that saves a
    pass and the groups stay small.  2^17 IDs above 2^23: 17 varying bits -> 16 in 2 passes, the finish kernel orders
    groups of 2 later IDs by the packed pair; 2^16 IDs: 16 bits, nothing to save; 2^17 IDs with ~10 pairs each: too many
    pairs per group, all 17 bits in 3 passes."""
    sc = bp.scenes.uniform_cubes(n_objects, 51, id_base=1 << 23, edge_factor=edge_factor)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"]); o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    gp, op = g.par_scan(), o.par_scan()
    _assert_pairs_equal(gp, op)
    assert gp.shape[0] > n_objects and _is_strictly_increasing(gp)
    assert g.stats()["pair_sort_passes"] == passes, (g.stats()["pair_sort_passes"], gp.shape[0])
    gf, of = g.par_scan_filtered(bp.ScanFilter.id_parity()), o.par_scan(co.FILTER_ID_PARITY)
    _assert_pairs_equal(gf, of)


def test_pair_sort_low_bits_with_duplicates_and_twins(bp):
    """IDs that own several bounds: duplicates and (a, b) / (b, a) twins inside groups that span two later IDs."""
    sysb, bounds, ids = _random_scene(2, 120_000, 78, multi_bounds=True, span=0.03)
    ids = (ids + np.uint32(1 << 23)).astype(np.uint32)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sysb, bounds, ids); o.extend(sysb, bounds, ids)
    gp, op = g.par_scan(), o.par_scan()
    _assert_pairs_equal(gp, op)
    assert _is_strictly_increasing(gp)
