"""Pins the CPU oracle: the reference's in-source known-answer tests, the restatement-derived
vectors of SURVEY.md section 8c, and agreement with the independent numpy restatement."""
import numpy as np
import pytest

from oracle import cpu_oracle as co
from oracle import pyref

import importlib.util, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "broadphase-rs_b200"))
import scenes  # noqa: E402  (numpy-only module of the product package)

KINDS = [co.INDEX32_2D, co.INDEX64_2D, co.INDEX64_3D]


# ---- reference KATs -------------------------------------------------------------------------

def test_kat_decode():  # src/index.rs:343-352
    assert co.decode_axis(co.INDEX64_3D, 0o0_001_111_111_111_111_111_111) == 0o1_777_777 << 13
    assert co.decode_axis(co.INDEX64_3D, 0o0_006_666_666_666_666_666_666) == 0
    assert int(pyref.decode_axis(2, np.uint64(0o0_001_111_111_111_111_111_111))) == 0o1_777_777 << 13
    assert int(pyref.decode_axis(2, np.uint64(0o0_006_666_666_666_666_666_666))) == 0


def test_kat_encode():  # src/index.rs:355-364
    assert co.encode_axis(co.INDEX64_3D, 0o1_777_777 << 13) == 0o0_001_111_111_111_111_111_111
    assert co.encode_axis(co.INDEX64_3D, 0) == 0
    assert int(pyref.encode_axis(2, np.uint32(0o1_777_777 << 13))) == 0o0_001_111_111_111_111_111_111


def test_kat_round_trip_axis():  # src/index.rs:367-374 (ChaCha stream replaced by Philox)
    rng = np.random.Generator(np.random.Philox(0))
    for v in rng.integers(0, 0o2_000_000, size=10000):
        expected = int(v) << 13
        assert co.decode_axis(co.INDEX64_3D, co.encode_axis(co.INDEX64_3D, expected)) == expected


def test_kat_system_bounds():  # src/geom.rs:696-706
    sysb = [-64, -64, -64, 64, 64, 64]
    box = np.array([-32, -32, -32, 32, 32, 32], dtype=np.float32)
    loc = co.to_local(3, sysb, box)
    assert (co.to_global(3, sysb, loc) == box).all()
    assert (pyref.to_global(sysb, pyref.to_local(sysb, box, 3), 3)[0] == box).all()


# ---- codec: fast spread == bit-by-bit definition ------------------------------------------------

@pytest.mark.parametrize("kind,axis_bits", [(0, 14), (2, 19)])
def test_encode_axis_exhaustive(kind, axis_bits):
    v = (np.arange(1 << axis_bits, dtype=np.uint64) << np.uint64(32 - axis_bits)).astype(np.uint32)
    ref = pyref.encode_axis(kind, v)
    step = 1 if axis_bits <= 14 else 37
    for i in range(0, v.shape[0], step):
        assert co.encode_axis(kind, int(v[i])) == int(ref[i])


def test_encode_axis_64_2d_sampled():
    rng = np.random.Generator(np.random.Philox(7))
    v = rng.integers(0, 1 << 32, size=5000, dtype=np.uint64).astype(np.uint32)
    ref = pyref.encode_axis(1, v)
    for i in range(v.shape[0]):
        got = co.encode_axis(1, int(v[i]))
        assert got == int(ref[i])
        assert co.decode_axis(1, got) == (int(v[i]) >> 3) << 3


def test_level_mask_and_layout():  # SURVEY.md section 8 bit-layout table
    assert co.level_mask(0, 1) == 0xC000_0000
    assert co.level_mask(1, 1) == 0x6000_0000_0000_0000
    assert co.level_mask(2, 1) == 0x3800_0000_0000_0000
    for kind in KINDS:
        assert co.level_mask(kind, 0) == 0
        _, dim, db, ab = pyref.KINDS[kind]
        for d in range(ab + 1):
            assert co.level_mask(kind, d) == pyref.level_mask(kind, d)
        full = ((1 << (dim * ab)) - 1) << db
        assert co.level_mask(kind, ab) == full


# ---- restatement-derived vectors (cross-checks, not reference output) -------------

def test_to_local_vectors():
    sysb = np.array([-64, -64, -64, 64, 64, 64], dtype=np.float32)
    for g, want in [(-32, 0x3FFFFFC0), (0, 0x7FFFFF80), (32, 0xBFFFFF00), (64, 0xFFFFFF00), (-64, 0)]:
        box = np.array([g] * 6, dtype=np.float32)
        assert co.to_local(3, sysb, box)[0] == want
        assert pyref.to_local(sysb, box, 3)[0, 0] == want


def _extend_keys(kind, sysb, box, min_depth=0):
    L = co.OracleLayer(kind, 4, min_depth)
    L.extend(sysb, np.array([box], dtype=np.float32), np.array([7], dtype=np.uint32))
    return [int(k) for k in L.records()[0]]


def test_extend_vectors():
    keys = _extend_keys(2, [-64, -64, -64, 64, 64, 64], [-32, -32, -32, 32, 32, 32])
    assert keys == [0x1 | (i << 59) for i in range(8)]
    b = [0.5, 0.25, 0.125, 0.5005, 0.2505, 0.1255]
    keys = _extend_keys(2, [0, 0, 0, 1, 1, 1], b)
    assert len(keys) == 8 and keys[0] == 0x017FFFFF0000000A and keys[-1] == 0x0A8000000000000A
    b2 = [0.5, 0.25, 0.5005, 0.2505]
    assert _extend_keys(0, [0, 0, 1, 1], b2) == [0x1FFFF00A, 0x4AAAA00A, 0x3555500A, 0x6000000A]
    assert _extend_keys(1, [0, 0, 1, 1], b2) == [
        0x0FFFF8000000000A, 0x255550000000000A, 0x1AAAA8000000000A, 0x300000000000000A]


# ---- oracle == independent numpy restatement -------------------------------------------------------

def _random_scene(kind, n, seed, multi_bounds=False, span=0.05):
    rng = np.random.Generator(np.random.Philox(seed))
    dim = co.DIM[kind]
    sysb = np.concatenate([np.full(dim, -3.0), np.full(dim, 5.0)]).astype(np.float32)
    size = (8.0 * span * rng.random((n, dim)) ** 3).astype(np.float32)
    mn = (-3.0 + rng.random((n, dim)) * (8.0 - size)).astype(np.float32)
    mx = (mn + size).astype(np.float32)
    bounds = np.concatenate([mn, mx], axis=1).astype(np.float32)
    # a few out-of-bounds and degenerate objects
    bounds[::97, 0] = -3.5
    bounds[5::89, dim:] = bounds[5::89, :dim]
    if multi_bounds:
        ids = rng.integers(0, max(2, n // 3), size=n).astype(np.uint32)
    else:
        ids = rng.permutation(n).astype(np.uint32)
    return sysb, bounds, ids


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("min_depth", [0, 3])
@pytest.mark.parametrize("multi", [False, True])
def test_oracle_matches_pyref(kind, min_depth, multi):
    sysb, bounds, ids = _random_scene(kind, 3000, 11 + kind + 10 * min_depth, multi, span=0.08)
    L = co.OracleLayer(kind, 4, min_depth)
    L.extend(sysb, bounds, ids)
    k, i = L.records()
    pk, pi = pyref.extend(kind, min_depth, sysb, bounds, ids)
    assert (k == pk).all() and (i == pi.astype(np.uint64)).all()
    assert not L.sorted
    L.sort()
    k, i = L.records()
    sk, si = pyref.sort_records(pk, pi)
    assert (k == sk).all() and (i == si.astype(np.uint64)).all()
    for fk, arg in [(co.FILTER_NONE, 0), (co.FILTER_ID_PARITY, 0), (co.FILTER_XOR_MASK, 6)]:
        got = L.scan(fk, arg)
        want, raw = pyref.scan(kind, sk, si, fk, arg)
        assert got.shape == want.shape and (got == want).all()
        assert L.num_raw_collisions == raw
        if got.shape[0] > 1:  # strictly increasing -- tests/test_layer.rs:42-54 with unique = true
            a, b = got[:, 0].astype(object), got[:, 1].astype(object)
            packed = a * (1 << 64) + b
            assert (packed[1:] > packed[:-1]).all()
    assert got.shape[0] > 0


@pytest.mark.parametrize("kind", KINDS)
def test_par_scan_equals_scan(kind):
    sysb, bounds, ids = _random_scene(kind, 20000, 5, False, span=0.02)
    for min_depth in (0, 4):
        A = co.OracleLayer(kind, 4, min_depth)
        B = co.OracleLayer(kind, 4, min_depth)
        A.extend(sysb, bounds, ids)
        B.extend(sysb, bounds, ids)
        B.par_sort()
        A.sort()
        assert all((x == y).all() for x, y in zip(A.records(), B.records()))
        assert (A.scan() == B.par_scan()).all()


def test_category_filter():
    sysb, bounds, ids = _random_scene(2, 2000, 3, False, span=0.1)
    rng = np.random.Generator(np.random.Philox(4))
    table = rng.integers(0, 16, size=(1500, 2)).astype(np.uint32)  # ids >= 1500 act as all-ones
    L = co.OracleLayer(2, 4, 0)
    L.extend(sysb, bounds, ids)
    got = L.scan(co.FILTER_CATEGORY, 0, table)
    k, i = L.records()
    want, _ = pyref.scan(2, k, i, pyref.FILTER_CATEGORY, 0, table)
    assert got.shape == want.shape and (got == want).all() and got.shape[0] > 0


def test_sphere_filter_is_the_narrow_phase():
    """FILTER_SPHERES = the pair test of the reference's example (examples/main.rs:461-479) as the scan filter: oracle vs
    pyref, and against the plain scan post-filtered the way the example does it."""
    sysb, bounds, ids = _random_scene(2, 3000, 11, False, span=0.1)
    centre = ((bounds[:, :3] + bounds[:, 3:]) * np.float32(0.5)).astype(np.float32)
    radius = ((bounds[:, 3:] - bounds[:, :3]).max(axis=1) * np.float32(0.5)).astype(np.float32)
    table = np.zeros((int(ids.max()) + 1, 4), dtype=np.float32)
    table[ids] = np.concatenate([centre, radius[:, None]], axis=1)
    table = table[:2500]                                    # ids >= 2500 pass unconditionally
    L = co.OracleLayer(2, 4, 0)
    L.extend(sysb, bounds, ids)
    got = L.scan(co.FILTER_SPHERES, 0, table)
    k, i = L.records()
    want, _ = pyref.scan(2, k, i, pyref.FILTER_SPHERES, 0, table)
    assert got.shape == want.shape and (got == want).all()
    every = L.scan()
    keep = pyref._filter(pyref.FILTER_SPHERES, 0, table, every[:, 0], every[:, 1])
    assert (every[keep] == got).all() and 0 < got.shape[0] < every.shape[0]


def test_u64_ids_and_merge():
    sysb, bounds, ids = _random_scene(2, 1500, 9, True, span=0.1)
    big = ids.astype(np.uint64) * np.uint64(0x1_0000_0001) + np.uint64(1 << 40)
    S = co.OracleLayer(2, 8, 2)
    S.extend(sysb, bounds[:1000], big[:1000])
    S.sort()
    D = co.OracleLayer(2, 8, 3)
    D.extend(sysb, bounds[1000:], big[1000:])
    D.sort()
    D.merge(S)
    assert D.min_depth == 2 and not D.sorted and len(D) > len(S)
    k, i = D.records()
    ks, is_ = S.records()
    assert (k[-len(S):] == ks).all() and (i[-len(S):] == is_).all()  # appended verbatim
    got = D.scan()
    sk, si = pyref.sort_records(k, i)
    want, _ = pyref.scan(2, sk, si)
    assert (got == want).all()
    E = co.OracleLayer(2, 8, 0)
    E.merge(co.OracleLayer(2, 8, 0))
    assert not E.sorted  # merge always clears the flag -- src/layer.rs:137


def test_edge_cases_scene():
    sc = scenes.edge_cases_3d()
    L = co.OracleLayer(sc["kind"], 4, 0)
    L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    k, i = L.records()
    pk, pi = pyref.extend(sc["kind"], 0, sc["sys_bounds"], sc["bounds"], sc["ids"])
    assert (k == pk).all() and (i == pi).all()
    present = set(int(x) for x in i)
    assert 7 not in present and 8 not in present and 11 not in present  # rejected by contains()
    assert 9 in present and 10 in present                                 # NaN passes
    assert [int(x) for x in k[i == 1]] == [0]                             # whole system -> default index
    assert (L.scan() == pyref.scan(sc["kind"], *pyref.sort_records(pk, pi))[0]).all()


def test_clear_and_flags():
    L = co.OracleLayer(2, 4, 0)
    assert L.sorted and len(L) == 0          # LayerBuilder::build -- src/layer.rs:681
    sc = scenes.uniform_cubes(100, 1)
    L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    assert not L.sorted
    L.clear()
    assert L.sorted and len(L) == 0          # src/layer.rs:84-88
    L.extend(sc["sys_bounds"], sc["bounds"][:0], sc["ids"][:0])
    assert L.sorted                           # no valid object -> flag untouched, src/layer.rs:119
    assert L.scan().shape == (0, 2)


def test_scene_recipes_have_expected_shape():
    sc = scenes.uniform_cubes(1 << 14, 2)
    L = co.OracleLayer(sc["kind"], 4, 0)
    L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    assert 3.0 < len(L) / (1 << 14) < 4.0     # SURVEY: R/N ~ 3.45
    p = L.par_scan()
    assert 1.0 < p.shape[0] / (1 << 14) < 3.0
    sc = scenes.example_circles(2000, 1)
    L = co.OracleLayer(sc["kind"], 4, sc["min_depth"])
    L.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    assert len(L) >= 2000 and L.scan().shape[0] > 0


# ---- the candidate set is a correct broadphase: every truly overlapping pair of boxes is in it -----------

@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("min_depth", [0, 3])
def test_candidate_pairs_cover_all_overlapping_boxes(kind, min_depth):
    """Independent of any restatement detail: two objects whose quantised boxes overlap share at least
    one cell chain, so the scan must report them (in one orientation or the other).  Brute force O(n^2)."""
    dim = co.DIM[kind]
    sysb, bounds, ids = _random_scene(kind, 1500, 77 + kind, False, span=0.12)
    ok = pyref.contains(sysb, bounds, dim)
    loc = pyref.to_local(sysb, bounds, dim).astype(np.int64)
    lo, hi = loc[:, :dim], loc[:, dim:]
    L = co.OracleLayer(kind, 4, min_depth)
    L.extend(sysb, bounds, ids)
    pairs = L.scan()
    got = set((int(a), int(b)) for a, b in pairs) | set((int(b), int(a)) for a, b in pairs)
    idx = np.flatnonzero(ok & (hi >= lo).all(axis=1))  # valid, non-inverted boxes
    missing = 0
    overlapping = 0
    for x in range(idx.shape[0]):
        i = idx[x]
        others = idx[x + 1:]
        ov = ((lo[others] <= hi[i]) & (hi[others] >= lo[i])).all(axis=1)
        for j in others[ov]:
            overlapping += 1
            if (int(ids[i]), int(ids[j])) not in got:
                missing += 1
    assert overlapping > 50
    assert missing == 0
