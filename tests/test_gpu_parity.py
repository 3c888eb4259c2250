"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Integer/byte work: the bar is bit-exact equality of the record sequence after extend, of the
record sequence after sort, and of the pair sequence after scan (tests/test_layer.rs:25-124 of the
reference make the same three comparisons against golden files)."""
import numpy as np
import pytest

from oracle import cpu_oracle as co

pytestmark = pytest.mark.gpu

KINDS = [0, 1, 2]


def _random_scene(kind, n, seed, multi_bounds=False, span=0.05, shuffle_ids=True):
    rng = np.random.Generator(np.random.Philox(seed))
    dim = co.DIM[kind]
    sysb = np.concatenate([np.full(dim, -3.0), np.full(dim, 5.0)]).astype(np.float32)
    size = (8.0 * span * rng.random((n, dim)) ** 3).astype(np.float32)
    mn = (-3.0 + rng.random((n, dim)) * (8.0 - size)).astype(np.float32)
    mx = (mn + size).astype(np.float32)
    bounds = np.concatenate([mn, mx], axis=1).astype(np.float32)
    bounds[::97, 0] = -3.5                       # outside the system -> dropped
    bounds[5::89, dim:] = bounds[5::89, :dim]    # zero extent
    if multi_bounds:
        ids = rng.integers(0, max(2, n // 3), size=n).astype(np.uint32)
    elif shuffle_ids:
        ids = rng.permutation(n).astype(np.uint32)
    else:
        ids = np.arange(n, dtype=np.uint32)
    return sysb, bounds, ids


def _pair(bp, kind, id_bytes, min_depth):
    g = bp.LayerBuilder().with_min_depth(min_depth).build(kind, "u32" if id_bytes == 4 else "u64")
    o = co.OracleLayer(kind, id_bytes, min_depth)
    return g, o


def _assert_records_equal(g, o):
    gk, gi = g.iter()
    ok, oi = o.records()
    assert gk.shape == ok.shape, (gk.shape, ok.shape)
    assert (gk.astype(np.uint64) == ok).all(), "keys differ at %s" % np.flatnonzero(gk.astype(np.uint64) != ok)[:5]
    assert (gi.astype(np.uint64) == oi).all(), "ids differ at %s" % np.flatnonzero(gi.astype(np.uint64) != oi)[:5]
    assert g.sorted == o.sorted
    assert len(g) == len(o)


def _assert_pairs_equal(gp, op):
    assert gp.shape == op.shape, (gp.shape, op.shape)
    assert (gp.astype(np.uint64) == op).all()


def _is_strictly_increasing(p):
    if p.shape[0] < 2:
        return True
    a, b = p[:, 0].astype(np.uint64), p[:, 1].astype(np.uint64)
    return bool(((a[1:] > a[:-1]) | ((a[1:] == a[:-1]) & (b[1:] > b[:-1]))).all())


# ---- extend -> sort -> scan, every index kind / ID width ----------------------------------------------

@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("id_bytes", [4, 8])
@pytest.mark.parametrize("min_depth", [0, 3])
def test_extend_sort_scan_small(bp, kind, id_bytes, min_depth):
    sysb, bounds, ids = _random_scene(kind, 5000, 100 + kind, span=0.06)
    if id_bytes == 8:
        ids = ids.astype(np.uint64) * np.uint64(0x100000001) + np.uint64(1 << 41)
    g, o = _pair(bp, kind, id_bytes, min_depth)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    _assert_records_equal(g, o)          # tests/test_layer.rs:25-40 `extend`
    g.sort()
    o.sort()
    _assert_records_equal(g, o)          # tests/test_layer.rs:56-90 `sort` / `par_sort`
    gp = g.scan()
    op = o.scan()
    _assert_pairs_equal(gp, op)          # tests/test_layer.rs:92-124 `scan` / `par_scan`
    assert gp.shape[0] > 0 and _is_strictly_increasing(gp)
    st = g.stats()
    assert st["n_raw_pairs"] == o.num_raw_collisions
    assert st["n_pairs"] == gp.shape[0]


@pytest.mark.parametrize("kind", KINDS)
def test_filters(bp, kind):
    sysb, bounds, ids = _random_scene(kind, 4000, 7, span=0.08)
    g, o = _pair(bp, kind, 4, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    rng = np.random.Generator(np.random.Philox(5))
    table = rng.integers(0, 16, size=(3000, 2)).astype(np.uint32)  # ids >= 3000 act as all-ones
    dim = co.DIM[kind]                                               # spheres around the objects (fused narrow phase)
    spheres = np.zeros((3500, 4), dtype=np.float32)
    spheres[ids[ids < 3500], :dim] = ((bounds[:, :dim] + bounds[:, dim:]) * np.float32(0.5))[ids < 3500]
    spheres[ids[ids < 3500], 3] = ((bounds[:, dim:] - bounds[:, :dim]).max(axis=1) * np.float32(0.45))[ids < 3500]
    for gf, of in [(bp.ScanFilter.id_parity(), (co.FILTER_ID_PARITY, 0, None)),
                   (bp.ScanFilter.spheres(spheres), (co.FILTER_SPHERES, 0, spheres)),
                   (bp.ScanFilter.xor_mask(6), (co.FILTER_XOR_MASK, 6, None)),
                   (bp.ScanFilter.category(table), (co.FILTER_CATEGORY, 0, table)),
                   (None, (co.FILTER_NONE, 0, None))]:
        gp = g.scan_filtered(gf).copy()
        op = o.scan(*of)
        _assert_pairs_equal(gp, op)
        assert _is_strictly_increasing(gp)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("id_bytes", [4, 8])
def test_multi_bounds_ids_trigger_inactive_records(bp, kind, id_bytes):
    """Several bounds per ID (static geometry, src/layer.rs:92-93): nested bounds of one ID make
    records the reference skips entirely (src/layer.rs:562-564)."""
    sysb, bounds, ids = _random_scene(kind, 6000, 21, multi_bounds=True, span=0.15)
    ids = ids.astype(np.uint64 if id_bytes == 8 else np.uint32)
    g, o = _pair(bp, kind, id_bytes, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    for flt, of in [(None, (co.FILTER_NONE, 0, None)), (bp.ScanFilter.id_parity(), (co.FILTER_ID_PARITY, 0, None))]:
        gp = g.scan_filtered(flt).copy()
        op = o.scan(*of)
        _assert_pairs_equal(gp, op)
    _assert_records_equal(g, o)
    assert g.stats()["rescans"] == 1


def test_edge_case_boxes(bp):
    sc = bp.scenes.edge_cases_3d()
    for min_depth in (0, 2):
        g, o = _pair(bp, sc["kind"], 4, min_depth)
        g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
        o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
        _assert_records_equal(g, o)
        _assert_pairs_equal(g.scan(), o.scan())
        _assert_records_equal(g, o)
        assert g.stats()["n_invalid"] == 0  # cleared by scan, like self.invalid (src/layer.rs:468)


def test_empty_and_flags(bp):
    g, o = _pair(bp, 2, 4, 0)
    assert g.sorted and len(g) == 0                       # src/layer.rs:681
    assert g.scan().shape == (0, 2)
    sc = bp.scenes.uniform_cubes(300, 1)
    g.extend(sc["sys_bounds"], sc["bounds"][:0], sc["ids"][:0])
    assert g.sorted and len(g) == 0                       # nothing appended: flag untouched
    outside = sc["bounds"][:4] + 10.0
    g.extend(sc["sys_bounds"], outside, sc["ids"][:4])
    assert g.sorted and len(g) == 0                       # every object rejected (src/layer.rs:108-111)
    g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    assert not g.sorted
    _assert_records_equal(g, o)
    g.clear()
    assert g.sorted and len(g) == 0                       # src/layer.rs:84-88
    g.extend(sc["sys_bounds"], sc["bounds"][:1], sc["ids"][:1])
    assert g.scan().shape == (0, 2) or len(g) > 1         # a single object has no pairs


def test_multiple_extends_and_ragged_sizes(bp):
    """Ragged batch sizes around the tile sizes of the kernels (1024 objects, 2048/4608-record tiles)."""
    sysb, bounds, ids = _random_scene(2, 9000, 33, span=0.05, shuffle_ids=False)
    g, o = _pair(bp, 2, 4, 0)
    cuts = [0, 1, 2, 1023, 1024, 1025, 3071, 4097, 9000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        g.extend(sysb, bounds[a:b], ids[a:b])
        o.extend(sysb, bounds[a:b], ids[a:b])
    _assert_records_equal(g, o)
    _assert_pairs_equal(g.scan(), o.scan())
    _assert_records_equal(g, o)


def test_extend_after_sort_uses_prefix_merge(bp):
    sysb, bounds, ids = _random_scene(2, 8000, 34, span=0.05)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sysb, bounds[:5000], ids[:5000])
    o.extend(sysb, bounds[:5000], ids[:5000])
    g.sort()
    o.sort()
    g.extend(sysb, bounds[5000:], ids[5000:])
    o.extend(sysb, bounds[5000:], ids[5000:])
    _assert_records_equal(g, o)           # sorted prefix followed by the new unsorted tail
    g.sort()
    o.sort()
    _assert_records_equal(g, o)
    assert g.stats()["merged"] == 1
    _assert_pairs_equal(g.scan(), o.scan())
    # a short sorted prefix followed by a long tail takes the full re-sort path instead
    g2, o2 = _pair(bp, 2, 4, 0)
    g2.extend(sysb, bounds[:300], ids[:300]); o2.extend(sysb, bounds[:300], ids[:300])
    g2.sort(); o2.sort()
    g2.extend(sysb, bounds[300:], ids[300:]); o2.extend(sysb, bounds[300:], ids[300:])
    g2.sort(); o2.sort()
    _assert_records_equal(g2, o2)
    assert g2.stats()["merged"] == 0


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("id_bytes", [4, 8])
def test_merge_static_into_dynamic(bp, kind, id_bytes):
    """Layer::merge (src/layer.rs:127-138): appended verbatim, min_depth lowered, flag cleared; the
    following sort + scan equal the oracle's full re-sort."""
    sysb, bounds, ids = _random_scene(kind, 12000, 55, span=0.05)
    ids = ids.astype(np.uint64 if id_bytes == 8 else np.uint32)
    gs, os_ = _pair(bp, kind, id_bytes, 2)
    gs.extend(sysb, bounds[:9000], ids[:9000]); os_.extend(sysb, bounds[:9000], ids[:9000])
    gs.sort(); os_.sort()
    gd, od = _pair(bp, kind, id_bytes, 3)
    gd.extend(sysb, bounds[9000:], ids[9000:]); od.extend(sysb, bounds[9000:], ids[9000:])
    gd.sort(); od.sort()
    gd.merge(gs); od.merge(os_)
    assert gd.min_depth == 2 and not gd.sorted
    _assert_records_equal(gd, od)
    gd.sort(); od.sort()
    _assert_records_equal(gd, od)
    assert gd.stats()["merged"] == 1
    _assert_pairs_equal(gd.scan(), od.scan())
    _assert_records_equal(gs, os_)        # the static layer is untouched
    # merging an unsorted layer, and merging into an unsorted layer
    ga, oa = _pair(bp, kind, id_bytes, 0)
    gb, ob = _pair(bp, kind, id_bytes, 0)
    ga.extend(sysb, bounds[:2000], ids[:2000]); oa.extend(sysb, bounds[:2000], ids[:2000])
    gb.extend(sysb, bounds[2000:5000], ids[2000:5000]); ob.extend(sysb, bounds[2000:5000], ids[2000:5000])
    ga.merge(gb); oa.merge(ob)
    _assert_records_equal(ga, oa)
    _assert_pairs_equal(ga.scan(), oa.scan())
    _assert_records_equal(ga, oa)
    # merging an empty layer still clears the flag (src/layer.rs:137)
    ge, _ = _pair(bp, kind, id_bytes, 0)
    ga.merge(ge)
    assert not ga.sorted
    ga.sort()
    assert ga.sorted


def test_set_records_roundtrip_and_scan(bp):
    sysb, bounds, ids = _random_scene(2, 3000, 77, span=0.08)
    o = co.OracleLayer(2, 4, 0)
    o.extend(sysb, bounds, ids)
    k, i = o.records()
    g = bp.Layer(2, "u32")
    g.set_records(k, i.astype(np.uint32), sorted_=False)
    gk, gi = g.iter()
    assert (gk == k).all() and (gi == i).all() and not g.sorted
    _assert_pairs_equal(g.scan(), o.scan())
    _assert_records_equal(g, o)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("id_bytes", [4, 8])
@pytest.mark.parametrize("min_depth", [0, 4])
def test_dedup_at_source_path(bp, kind, id_bytes, min_depth):
    """clear -> extend -> scan without looking at the records in between: the sort moves the cell flags
    into the IDs and the scan emits every ID pair from one canonical shared cell only.  The result must
    still be the oracle's, and fewer raw pairs must have been produced than the reference's sweep makes."""
    sysb, bounds, ids = _random_scene(kind, 20000, 300 + kind + min_depth, span=0.03, shuffle_ids=False)
    if id_bytes == 8:
        ids = ids.astype(np.uint64) + np.uint64(1 << 40)
    g, o = _pair(bp, kind, id_bytes, min_depth)
    table = np.random.Generator(np.random.Philox(3)).integers(0, 16, size=(15000, 2)).astype(np.uint32)
    cases = [(None, (co.FILTER_NONE, 0, None)), (bp.ScanFilter.id_parity(), (co.FILTER_ID_PARITY, 0, None)),
             (bp.ScanFilter.xor_mask(5), (co.FILTER_XOR_MASK, 5, None))]
    if id_bytes == 4:
        cases.append((bp.ScanFilter.category(table), (co.FILTER_CATEGORY, 0, table)))
    for flt, oflt in cases:
        g.clear(); o.clear()
        g.extend(sysb, bounds[:12000], ids[:12000]); o.extend(sysb, bounds[:12000], ids[:12000])
        g.extend(sysb, bounds[12000:], ids[12000:]); o.extend(sysb, bounds[12000:], ids[12000:])
        gp = g.scan_filtered(flt)
        op = o.scan(*oflt)
        _assert_pairs_equal(gp, op)
        st = g.stats()
        assert st["n_pairs"] <= st["n_raw_pairs"] < o.num_raw_collisions
        assert st["rescans"] == 0
    _assert_records_equal(g, o)   # the flags are gone again once the records are looked at
    _assert_pairs_equal(g.scan(), o.scan())


def test_dedup_at_source_large_ids_are_left_alone(bp):
    """IDs that use the top 3 bits of their type leave no room for the cell flags: plain emission."""
    sysb, bounds, ids = _random_scene(2, 5000, 41, span=0.05, shuffle_ids=False)
    big = (ids.astype(np.uint64) + np.uint64(0xE0000000)).astype(np.uint32)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sysb, bounds, big); o.extend(sysb, bounds, big)
    _assert_pairs_equal(g.scan(), o.scan())
    assert g.stats()["n_raw_pairs"] == o.num_raw_collisions
    _assert_records_equal(g, o)


@pytest.mark.parametrize("id_bytes", [4, 8])
def test_crowded_cell_takes_the_full_width_pair_sort(bp, id_bytes):
    """Hundreds of objects in one cell: the later ID with the most partners has more of them than
    pair_finish_kernel's window, so the scan must fall back to the full-width pair sort."""
    sc = bp.scenes.uniform_cubes(20_000, 12)
    crowd = 900
    sc["bounds"][:crowd] = np.array([0.4001, 0.4001, 0.4001, 0.4019, 0.4019, 0.4019], dtype=np.float32)
    # a second, smaller crowd that still fits the window (up to 8 raw copies of each of its ~40 pairs per ID)
    sc["bounds"][crowd:crowd + 40] = np.array([0.7001, 0.2001, 0.6001, 0.7012, 0.2012, 0.6012], dtype=np.float32)
    ids = sc["ids"].astype(np.uint64 if id_bytes == 8 else np.uint32)
    g = bp.Layer(2, "u32" if id_bytes == 4 else "u64")
    o = co.OracleLayer(2, id_bytes, 0)
    g.extend(sc["sys_bounds"], sc["bounds"], ids)
    o.extend(sc["sys_bounds"], sc["bounds"], ids)
    gp = g.par_scan()
    op = o.par_scan()
    _assert_pairs_equal(gp, op)
    assert gp.shape[0] > crowd * (crowd - 1) // 2
    # the full-width fallback (>= 4 passes over 2 x 15 ID bits), after the later-ID passes (u64 IDs) or the
    # counting sort by later ID (dense u32 IDs, no radix pass)
    assert g.stats()["pair_sort_passes"] >= (4 if id_bytes == 4 else 5)
    # without the big crowd the finish kernel handles everything (fewer passes)
    g2 = bp.Layer(2, "u32" if id_bytes == 4 else "u64")
    o2 = co.OracleLayer(2, id_bytes, 0)
    g2.extend(sc["sys_bounds"], sc["bounds"][crowd:], ids[crowd:])
    o2.extend(sc["sys_bounds"], sc["bounds"][crowd:], ids[crowd:])
    _assert_pairs_equal(g2.par_scan(), o2.par_scan())
    assert g2.stats()["pair_sort_passes"] <= 3


@pytest.mark.parametrize("id_base", [0, 1 << 23])
@pytest.mark.parametrize("multi", [False, True])
def test_pair_sort_counting_and_radix_paths(bp, id_base, multi):
    """Dense small u32 IDs take the counting sort of the pairs (count per later ID in scan_emit_kernel, scan,
    scatter); IDs at or above 2^22 take the radix passes over the later ID.  Same pairs either way; IDs that
    own several bounds add duplicates and (a, b) / (b, a) twins that only the finish kernel removes."""
    sysb, bounds, ids = _random_scene(2, 30_000, 77, multi_bounds=multi, span=0.08)
    ids = (ids + np.uint32(id_base)).astype(np.uint32)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    gp, op = g.par_scan(), o.par_scan()
    _assert_pairs_equal(gp, op)
    assert gp.shape[0] > 1000 and _is_strictly_increasing(gp)
    passes = g.stats()["pair_sort_passes"]
    if id_base:
        assert passes >= 1
    elif not multi:   # (with several bounds per ID a crowded later ID may still force the full-width fallback)
        assert passes == 0
    gf = g.par_scan_filtered(bp.ScanFilter.id_parity())
    of = o.par_scan(co.FILTER_ID_PARITY)
    _assert_pairs_equal(gf, of)


# ---- BASELINE.json configs at sizes the oracle finishes in seconds --------------------------------------

def _run_scene(bp, sc, id_type="u32", flt=None, oflt=(0, 0, None), check_unsorted=True):
    g = bp.LayerBuilder().with_min_depth(sc["min_depth"]).build(sc["kind"], id_type)
    o = co.OracleLayer(sc["kind"], 4 if id_type == "u32" else 8, sc["min_depth"])
    g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    if check_unsorted:
        _assert_records_equal(g, o)
    g.par_sort()
    o.par_sort()
    _assert_records_equal(g, o)
    gp = g.par_scan_filtered(flt)
    op = o.par_scan(*oflt)
    _assert_pairs_equal(gp, op)
    assert _is_strictly_increasing(gp)
    return g, o, gp


def test_config1_example_circles(bp):
    """BASELINE config 1: 10,000 circles, Index32_2D, min_depth 4, par_scan."""
    g, o, gp = _run_scene(bp, bp.scenes.example_circles(10_000, 1))
    assert gp.shape[0] > 1000


def test_config2_uniform_1m(bp):
    """BASELINE config 2 at full size: 2^20 uniform cubes, Index64_3D."""
    g, o, gp = _run_scene(bp, bp.scenes.uniform_cubes(1 << 20, 2))
    st = g.stats()
    assert st["sort_passes"] <= 4          # all keys at one depth: 21 varying origin bits
    assert 3.0 < st["n_records"] / (1 << 20) < 4.0


def test_config3_lognormal_parity_filter(bp):
    """BASELINE config 3 recipe at 2^19 objects (multi-depth keys, skewed run lengths), ID-parity filter."""
    sc = bp.scenes.lognormal_cubes(1 << 19, 3)
    g, o, gp = _run_scene(bp, sc, flt=bp.ScanFilter.id_parity(), oflt=(co.FILTER_ID_PARITY, 0, None))
    assert ((gp[:, 0] ^ gp[:, 1]) & 1).all()


def test_gen_boxes_reference_test_scene(bp):
    """The reference's own test-scene recipe (tests/gen_test_scenes.py: density 1/1000, sizes 1-10)."""
    for n in (100, 1000, 10_000, 100_000):
        _run_scene(bp, bp.scenes.gen_boxes(n, 0))


def test_config4_static_plus_dynamic_frames(bp):
    """BASELINE config 4 shape at reduced size: a static layer sorted once, a fresh dynamic layer per
    frame merged with it."""
    ns, nd = 1 << 18, 1 << 14
    st = bp.scenes.uniform_cubes(ns, 4)
    gs = bp.Layer(2, "u32"); os_ = co.OracleLayer(2, 4, 0)
    gs.extend(st["sys_bounds"], st["bounds"], st["ids"]); os_.extend(st["sys_bounds"], st["bounds"], st["ids"])
    gs.sort(); os_.par_sort()
    gd = bp.Layer(2, "u32"); od = co.OracleLayer(2, 4, 0)
    for frame in range(3):
        dy = bp.scenes.uniform_cubes(nd, 5 + frame, id_base=ns, edge_factor=0.4 * (ns / nd) ** (-1.0 / 3.0))
        gd.clear(); od.clear()
        gd.extend(dy["sys_bounds"], dy["bounds"], dy["ids"]); od.extend(dy["sys_bounds"], dy["bounds"], dy["ids"])
        gd.sort(); od.par_sort()
        gd.merge(gs); od.merge(os_)
        gp = gd.par_scan(); op = od.par_scan()
        _assert_pairs_equal(gp, op)
        _assert_records_equal(gd, od)
        assert gd.stats()["merged"] == 1


@pytest.mark.parametrize("mutation", ["none", "clear_other", "extend_other", "destroy_other", "merge_twice", "iter_first"])
def test_deferred_merge_is_unobservable(bp, mutation):
    """Layer::merge of a sorted layer into a sorted layer is deferred (the next sort merges straight out of both trees,
    the cell flags of dedup-at-the-source stay in the IDs).  Whatever happens to the other layer in between, the
    result is that of the reference's eager append (src/layer.rs:127-138)."""
    sc = bp.scenes.uniform_cubes(60_000, 21)
    a, b = slice(0, 45_000), slice(45_000, 60_000)
    gs, os_ = _pair(bp, 2, 4, 0)
    gd, od = _pair(bp, 2, 4, 0)
    gs.extend(sc["sys_bounds"], sc["bounds"][a], sc["ids"][a]); os_.extend(sc["sys_bounds"], sc["bounds"][a], sc["ids"][a])
    gd.extend(sc["sys_bounds"], sc["bounds"][b], sc["ids"][b]); od.extend(sc["sys_bounds"], sc["bounds"][b], sc["ids"][b])
    gs.sort(); os_.sort(); gd.sort(); od.sort()
    gd.merge(gs); od.merge(os_)
    assert len(gd) == len(od) and not gd.sorted
    extra = bp.scenes.uniform_cubes(5_000, 22, id_base=100_000)
    if mutation == "clear_other":
        gs.clear()
    elif mutation == "extend_other":
        gs.extend(extra["sys_bounds"], extra["bounds"], extra["ids"])
    elif mutation == "destroy_other":
        gs.close()
    elif mutation == "merge_twice":
        g2, o2 = _pair(bp, 2, 4, 0)
        g2.extend(extra["sys_bounds"], extra["bounds"], extra["ids"]); o2.extend(extra["sys_bounds"], extra["bounds"], extra["ids"])
        g2.sort(); o2.sort()
        gd.merge(g2); od.merge(o2)
    elif mutation == "iter_first":
        _assert_records_equal(gd, od)          # the appended, still unsorted tree
    gp, op = gd.par_scan(), od.par_scan()
    _assert_pairs_equal(gp, op)
    _assert_records_equal(gd, od)
    if mutation in ("none",):
        st = gd.stats()
        assert st["merged"] == 1 and st["n_raw_pairs"] == gp.shape[0]   # flags kept: no duplicate reached the pair sort


def test_extend_from_pinned_host_buffers(bp):
    """bp_layer_extend_host reads page-locked buffers in place (the encode kernel streams them over PCIe); pageable
    buffers go through a staging copy.  Same records either way."""
    import torch
    sc = bp.scenes.lognormal_cubes(300_001, 8)
    hb = torch.from_numpy(sc["bounds"]).pin_memory().numpy()
    hi = torch.from_numpy(sc["ids"].view(np.int32)).pin_memory().numpy().view(np.uint32)
    g, o = _pair(bp, 2, 4, 0)
    g.extend(sc["sys_bounds"], hb, hi)
    o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    _assert_records_equal(g, o)
    g.extend(sc["sys_bounds"], hb[1:1001], hi[1:1001])       # a misaligned slice of the pinned buffer: staged
    o.extend(sc["sys_bounds"], sc["bounds"][1:1001], sc["ids"][1:1001])
    _assert_records_equal(g, o)
    _assert_pairs_equal(g.par_scan(), o.par_scan())
    # beyond 64 MB the copy engine takes over: chunked copies on a second stream, one encode per chunk behind them
    big = bp.scenes.uniform_cubes(3_000_003, 12)
    hb = torch.from_numpy(big["bounds"]).pin_memory().numpy()
    hi = torch.from_numpy(big["ids"].view(np.int32)).pin_memory().numpy().view(np.uint32)
    g, o = _pair(bp, 2, 4, 0)
    for _ in range(2):                                       # twice: the staging buffers are reused
        g.clear(); o.clear()
        g.extend(big["sys_bounds"], hb, hi)
        o.extend(big["sys_bounds"], big["bounds"], big["ids"])
        _assert_records_equal(g, o)
    _assert_pairs_equal(g.par_scan(), o.par_scan())


def test_non_ascending_ids_need_id_passes(bp):
    """IDs in random order: the sort must order equal keys by ID (derived Ord of (Index, ID))."""
    sc = bp.scenes.uniform_cubes(200_000, 9)
    rng = np.random.Generator(np.random.Philox(1))
    sc["ids"] = rng.permutation(200_000).astype(np.uint32)
    sc["bounds"][:, 3:] = np.minimum(sc["bounds"][:, :3] + 0.02, 1.0)  # bigger cubes: many equal keys
    g, o, gp = _run_scene(bp, sc)
    assert g.stats()["sort_passes"] >= 4


# ---- BASELINE configs at FULL size against the oracle (about a minute of host time each) -------------------

def _full_size(bp, sc, flt=None, oflt=(0, 0, None)):
    """extend -> par_sort -> par_scan(_filtered) from device-resident inputs; records after the sort and the pair list
    compared bit for bit with the oracle (tests/test_layer.rs:57-124's equalities at BASELINE's sizes)."""
    import torch
    n = sc["bounds"].shape[0]
    g = bp.LayerBuilder().with_min_depth(sc["min_depth"]).build(sc["kind"], "u32")
    o = co.OracleLayer(sc["kind"], 4, sc["min_depth"])
    db = torch.from_numpy(sc["bounds"]).cuda()
    di = torch.from_numpy(sc["ids"].view(np.int32)).cuda()
    g.extend_device(sc["sys_bounds"], db, di, n)
    o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    g.par_sort()
    o.par_sort()
    _assert_records_equal(g, o)
    gp = g.par_scan_filtered(flt)
    op = o.par_scan(*oflt)
    _assert_pairs_equal(gp, op)
    assert _is_strictly_increasing(gp)
    return g, gp


def test_config3_full_size_against_oracle(bp):
    """BASELINE config 3 as quoted: 2^24 log-normal cubes, Index64_3D, scan_filtered with the ID-parity filter."""
    g, gp = _full_size(bp, bp.scenes.lognormal_cubes(1 << 24, 3), bp.ScanFilter.id_parity(), (co.FILTER_ID_PARITY, 0, None))
    assert gp.shape[0] > (1 << 23) and ((gp[:, 0] ^ gp[:, 1]) & 1).all()
    assert g.stats()["n_records"] > 4 * (1 << 24)


def test_config5_per_gpu_shape_full_size_against_oracle(bp):
    """The bench's headline shape on one GPU: 2^25 uniform cubes (BASELINE config 5's per-GPU share)."""
    g, gp = _full_size(bp, bp.scenes.uniform_cubes(1 << 25, 6))
    assert gp.shape[0] > (1 << 25)


def test_config4_static_plus_dynamic_large(bp):
    """BASELINE config 4 at 1/16 of its size: 2^22 static objects sorted once + 2^20 dynamic objects per frame through
    Layer::merge (the reference appends and re-sorts; here one merge-path merge), two frames."""
    import torch
    ns, nd = 1 << 22, 1 << 20
    st = bp.scenes.uniform_cubes(ns, 4)
    gs = bp.Layer(2, "u32"); os_ = co.OracleLayer(2, 4, 0)
    gs.extend(st["sys_bounds"], st["bounds"], st["ids"]); os_.extend(st["sys_bounds"], st["bounds"], st["ids"])
    gs.sort(); os_.par_sort()
    gd = bp.Layer(2, "u32"); od = co.OracleLayer(2, 4, 0)
    for frame in range(2):
        dy = bp.scenes.uniform_cubes(nd, 5 + frame, id_base=ns, edge_factor=0.4 * (ns / nd) ** (-1.0 / 3.0))
        gd.clear(); od.clear()
        gd.extend_device(dy["sys_bounds"], torch.from_numpy(dy["bounds"]).cuda(), torch.from_numpy(dy["ids"].view(np.int32)).cuda(), nd)
        od.extend(dy["sys_bounds"], dy["bounds"], dy["ids"])
        gd.sort(); od.par_sort()
        gd.merge(gs); od.merge(os_)
        gp = gd.par_scan(); op = od.par_scan()
        _assert_pairs_equal(gp, op)
        _assert_records_equal(gd, od)
        assert gd.stats()["merged"] == 1


# ---- full-size properties (no oracle): BASELINE config 3 at 2^24 objects -----------------------------

def test_config3_full_size_properties(bp):
    import torch
    n = 1 << 24
    sc = bp.scenes.lognormal_cubes(n, 3)
    g = bp.Layer(2, "u32")
    db = torch.from_numpy(sc["bounds"]).cuda()
    di = torch.from_numpy(sc["ids"].astype(np.int32)).cuda()
    g.extend_device(sc["sys_bounds"], db, di, n)
    g.par_sort()
    kp, ip, nrec, is_sorted = g.records_device()
    assert is_sorted and nrec > 4 * n
    keys, ids = g.iter()
    k64 = keys.astype(np.uint64)
    nondecreasing = (k64[1:] > k64[:-1]) | ((k64[1:] == k64[:-1]) & (ids[1:] >= ids[:-1]))
    assert nondecreasing.all()                              # tests/test_layer.rs:42-54, unique = false
    # the multiset of records is preserved by the sort: order-independent checksums
    g2 = bp.Layer(2, "u32")
    g2.extend_device(sc["sys_bounds"], db, di, n)
    k2, i2 = g2.iter()
    assert k2.shape == keys.shape
    assert int(k2.sum(dtype=np.uint64)) == int(keys.sum(dtype=np.uint64))
    assert int((k2 ^ (i2.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))).sum(dtype=np.uint64)) == \
        int((keys ^ (ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))).sum(dtype=np.uint64))
    p1 = g.par_scan_filtered(bp.ScanFilter.id_parity()).copy()
    assert p1.shape[0] > 0 and _is_strictly_increasing(p1)   # sorted + unique
    assert ((p1[:, 0] ^ p1[:, 1]) & 1).all()                 # filter applied
    p2 = g.par_scan_filtered(bp.ScanFilter.id_parity())
    assert p1.shape == p2.shape and (p1 == p2).all()         # idempotent
    # the filtered result is exactly the subset of the unfiltered one that passes the filter
    pall = g.par_scan()
    sub = pall[((pall[:, 0] ^ pall[:, 1]) & 1) == 1]
    assert sub.shape == p1.shape and (sub == p1).all()


def test_oversized_object_rejects_the_whole_extend(bp):
    """An object that wants more than 2^20 cells (min_depth far above its natural depth; the reference warn!s and
    heap-allocates, src/geom.rs:299-301) is a documented limit of the encoder: the extend call is rejected as a whole with
    BP_ERR_TOO_LARGE and the tree is what it was before the call."""
    sc = bp.scenes.uniform_cubes(5_000, 61)
    g = bp.LayerBuilder().with_min_depth(8).build(2, "u32")
    o = co.OracleLayer(2, 4, 8)
    g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    n_before = len(g)
    assert n_before == len(o)
    more = bp.scenes.uniform_cubes(1_000, 62, id_base=5_000)
    bad = more["bounds"].copy()
    bad[500] = np.array([0.1, 0.1, 0.1, 0.62, 0.62, 0.62], dtype=np.float32)   # 134^3 cells at depth 8
    with pytest.raises(bp.BpError) as e:
        g.extend(sc["sys_bounds"], bad, more["ids"])
        len(g)                                                                   # (the error may only surface at the next call)
    assert e.value.status == 4
    assert len(g) == n_before
    _assert_records_equal(g, o)
    _assert_pairs_equal(g.par_scan(), o.par_scan())
    g.extend(more["sys_bounds"], more["bounds"], more["ids"])                    # the layer goes on as if nothing had happened
    o.extend(more["sys_bounds"], more["bounds"], more["ids"])
    _assert_pairs_equal(g.par_scan(), o.par_scan())
    _assert_records_equal(g, o)
