"""Batched box / ray queries (Layer::test_box / test_ray, src/layer.rs:244-351).

CPU part: the C++ oracle's literal test_impl recursion against pyref's closed form (every record replays its
own path of cells), plus a geometric completeness check.  GPU part: the CUDA descent through the C ABI against
the oracle, query by query, bit-exact.  The reference holds no test or fixture for its queries and the cell
centres come from an un-vendored dependency (cgmath's midpoint): "parity unpinned", see oracle/bp_oracle.cpp."""
import numpy as np
import pytest

from oracle import cpu_oracle as co
from oracle import pyref

KINDS = [0, 1, 2]
SYS = {2: np.array([-3, -3, 5, 5], dtype=np.float32), 3: np.array([-3, -3, -3, 5, 5, 5], dtype=np.float32)}


def _scene(kind, n, seed, span=0.05):
    rng = np.random.Generator(np.random.Philox(seed))
    dim = co.DIM[kind]
    size = (8.0 * span * rng.random((n, dim)) ** 3).astype(np.float32)
    mn = (-3.0 + rng.random((n, dim)) * (8.0 - size)).astype(np.float32)
    bounds = np.concatenate([mn, (mn + size).astype(np.float32)], axis=1).astype(np.float32)
    bounds[:3, :dim] = -3.0          # three scene-sized objects: shallow cells
    bounds[:3, dim:] = 5.0
    ids = rng.permutation(n).astype(np.uint32)
    ids[n // 2:] = ids[: n - n // 2]  # IDs that own two bounds: the result must still be duplicate-free
    return SYS[dim], bounds, ids


def _boxes(dim, nq, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    size = (8.0 * 0.3 * rng.random((nq, dim)) ** 4).astype(np.float32)
    mn = (-3.5 + rng.random((nq, dim)) * 8.5).astype(np.float32)
    b = np.concatenate([mn, (mn + size).astype(np.float32)], axis=1).astype(np.float32)
    b[0, :dim], b[0, dim:] = -10.0, 10.0      # everything
    b[1, :dim], b[1, dim:] = 6.0, 7.0         # outside the system
    b[2, :dim], b[2, dim:] = 1.0, 1.0         # a point on a cell boundary (cells are closed: both sides report)
    return b


def _rays(dim, nq, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    org = (-4.0 + rng.random((nq, dim)) * 10.0).astype(np.float32)
    d = rng.normal(size=(nq, dim)).astype(np.float32)
    rmin = np.full((nq, 1), -np.inf, dtype=np.float32)
    rmax = np.full((nq, 1), np.inf, dtype=np.float32)
    rmin[::3] = 0.0                            # half lines
    rmax[::5] = 2.5                            # segments
    d[4, 0] = 0.0                              # axis-parallel rays: the division yields +-inf / NaN
    d[5, :] = 0.0
    d[5, dim - 1] = -1.0
    d[6, :] = 0.0                              # a degenerate direction
    org[7, :] = 1.0                            # starts on a cell boundary
    return np.concatenate([org, d, rmin, rmax], axis=1).astype(np.float32)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("max_depth", [None, 0, 3, 40])
def test_oracle_queries_match_the_closed_form(kind, max_depth):
    dim = co.DIM[kind]
    sysb, bounds, ids = _scene(kind, 3000, 200 + kind)
    o = co.OracleLayer(kind, 4, 0)
    o.extend(sysb, bounds, ids)
    o.sort()
    keys, rids = o.records()
    for q in _boxes(dim, 40, 300 + kind):
        got = o.test_box(sysb, q, max_depth)
        want = pyref.test_box(kind, keys, rids, sysb, q, max_depth)
        assert got.shape == want.shape and (got == want).all()
    for q in _rays(dim, 40, 400 + kind):
        got = o.test_ray(sysb, q[:dim], q[dim:2 * dim], q[2 * dim], q[2 * dim + 1], max_depth)
        want = pyref.test_ray(kind, keys, rids, sysb, q[:dim], q[dim:2 * dim], q[2 * dim], q[2 * dim + 1], max_depth)
        assert got.shape == want.shape and (got == want).all()


@pytest.mark.parametrize("kind", KINDS)
def test_box_query_is_complete(kind):
    """Cells cover their objects, so every object whose own AABB overlaps the test box must be reported
    (the converse does not hold: the test is at cell granularity)."""
    dim = co.DIM[kind]
    sysb, bounds, _ = _scene(kind, 2000, 500 + kind)
    ids = np.arange(bounds.shape[0], dtype=np.uint32)
    o = co.OracleLayer(kind, 4, 0)
    o.extend(sysb, bounds, ids)
    inside = ((bounds[:, :dim] >= sysb[:dim]) & (bounds[:, dim:] <= sysb[dim:])).all(axis=1)
    for q in _boxes(dim, 25, 600 + kind):
        got = set(o.test_box(sysb, q).tolist())
        hit = inside & ((bounds[:, :dim] <= q[dim:]) & (bounds[:, dim:] >= q[:dim])).all(axis=1)
        assert set(np.flatnonzero(hit).tolist()) <= got


def test_empty_layer_and_results_are_sorted_unique():
    o = co.OracleLayer(2, 4, 0)
    assert o.test_box(SYS[3], [0, 0, 0, 1, 1, 1]).shape == (0,)
    sysb, bounds, ids = _scene(2, 500, 9)
    o.extend(sysb, bounds, ids)
    r = o.test_box(sysb, [-3, -3, -3, 5, 5, 5])
    assert (np.diff(r.astype(np.int64)) > 0).all() and r.shape[0] == np.unique(ids).shape[0]


# ---- pick_ray ---------------------------------------------------------------------------------------------------

def _sphere_dist(shape, org, d):
    """PICK_SPHERE in numpy float32, the same operations in the same order as oracle/bp_oracle.cpp shape_distance."""
    f = np.float32
    dim = org.shape[0]
    b = (shape[:dim] - org).astype(f)
    p, m = (d * b).astype(f), (b * b).astype(f)
    proj, mag2 = p[0], m[0]
    for i in range(1, dim):
        proj, mag2 = f(proj + p[i]), f(mag2 + m[i])
    with np.errstate(invalid="ignore"):
        ext = np.sqrt(f(f(f(proj * proj) - mag2) + f(shape[dim] * shape[dim])), dtype=f)
    lo, hi = f(proj - ext), f(proj + ext)
    if hi < 0:
        return f(np.inf)
    if lo < 0:
        return f(0)
    return lo if np.isfinite(lo) else f(np.inf)


def _pick_scene(kind, n, seed):
    """Disjoint IDs (one bound each); spheres inscribed in the bounds, so a hit point always lies inside the object's cells."""
    dim = co.DIM[kind]
    sysb, bounds, _ = _scene(kind, n, seed, span=0.04)
    bounds = bounds[3:]                                   # (drop the scene-sized objects: their inscribed spheres hide everything)
    ids = np.arange(bounds.shape[0], dtype=np.uint32)
    centre = ((bounds[:, :dim] + bounds[:, dim:]) * np.float32(0.5)).astype(np.float32)
    radius = ((bounds[:, dim:] - bounds[:, :dim]).min(axis=1) * np.float32(0.5)).astype(np.float32)
    spheres = np.concatenate([centre, radius[:, None]], axis=1).astype(np.float32)
    return sysb, bounds, ids, spheres


def _pick_rays(dim, nq, seed, targets=None):
    """Rays from random origins; two thirds of them aimed at an object's centre (something gets hit on the way)."""
    rng = np.random.Generator(np.random.Philox(seed))
    org = (-3.0 + rng.random((nq, dim)) * 8.0).astype(np.float32)
    d = rng.normal(size=(nq, dim)).astype(np.float32)
    if targets is not None:
        aim = targets[rng.integers(0, targets.shape[0], size=nq)] - org
        sel = np.arange(nq) % 3 != 0
        d[sel] = aim[sel]
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    d[0, :] = 0.0
    d[0, 0] = 1.0                                          # axis-parallel
    return np.concatenate([org, d.astype(np.float32)], axis=1).astype(np.float32)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("max_dist", [np.inf, 4.0])
def test_oracle_pick_ray_is_the_minimum_over_the_ray_test(kind, max_dist):
    """pick_ray's pruning must not change the answer: the distance equals the minimum of get_dist over everything
    test_ray(0, max_dist) reports (independent closed form, pyref), and the ID is one that attains it."""
    dim = co.DIM[kind]
    sysb, bounds, ids, spheres = _pick_scene(kind, 1500, 40 + kind)
    o = co.OracleLayer(kind, 4, 0)
    o.extend(sysb, bounds, ids)
    o.sort()
    keys, rids = o.records()
    hits = 0
    for ray in _pick_rays(dim, 60, 50 + kind, spheres[:, :dim]):
        org, d = ray[:dim], ray[dim:]
        got = o.pick_ray(sysb, org, d, max_dist, co.PICK_SPHERE, spheres)
        cand = pyref.test_ray(kind, keys, rids, sysb, org, d, 0.0, max_dist)
        dists = np.array([_sphere_dist(spheres[int(c)], org, d) for c in cand], dtype=np.float32)
        best = dists.min() if dists.size else np.float32(np.inf)
        if not (best < np.float32(max_dist)):
            assert got is None
            continue
        hits += 1
        assert got is not None and np.float32(got[0]) == best
        assert got[1] in set(int(c) for c in cand[dists == best])
        assert (got[2] == (org + d * np.float32(got[0])).astype(np.float32)).all()
    assert hits > (5 if np.isinf(max_dist) else 0)


def test_oracle_pick_ray_aabb_and_ties():
    """Two identical boxes: the one met first in the walk (smaller ID in the same cell) wins the tie; the slab distance
    is the entry distance, 0 from inside, None when the ray points away."""
    sysb = SYS[3]
    b = np.array([[0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 1], [2, 0, 0, 3, 1, 1]], dtype=np.float32)
    o = co.OracleLayer(2, 4, 0)
    o.extend(sysb, b, np.array([5, 4, 9], dtype=np.uint32))
    shapes = np.zeros((10, 6), dtype=np.float32)
    shapes[[5, 4, 9]] = b
    got = o.pick_ray(sysb, [-1, 0.5, 0.5], [1, 0, 0], np.inf, co.PICK_AABB, shapes)
    assert got is not None and got[0] == 1.0 and got[1] == 4 and got[2].tolist() == [0.0, 0.5, 0.5]
    assert o.pick_ray(sysb, [0.5, 0.5, 0.5], [1, 0, 0], np.inf, co.PICK_AABB, shapes)[0] == 0.0
    assert o.pick_ray(sysb, [1.5, 0.5, 0.5], [1, 0, 0], np.inf, co.PICK_AABB, shapes)[1] == 9
    assert o.pick_ray(sysb, [4, 0.5, 0.5], [1, 0, 0], np.inf, co.PICK_AABB, shapes) is None
    assert o.pick_ray(sysb, [-1, 0.5, 0.5], [1, 0, 0], 0.5, co.PICK_AABB, shapes) is None      # max_dist


# ---- GPU ------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("id_bytes", [4, 8])
@pytest.mark.parametrize("max_depth", [None, 0, 4])
def test_gpu_query_batches_match_the_oracle(bp, kind, id_bytes, max_depth):
    dim = co.DIM[kind]
    sysb, bounds, ids = _scene(kind, 6000, 700 + kind)
    if id_bytes == 8:
        ids = ids.astype(np.uint64) * np.uint64(0x100000001) + np.uint64(1 << 40)
    g = bp.LayerBuilder().build(kind, "u32" if id_bytes == 4 else "u64")
    o = co.OracleLayer(kind, id_bytes, 0)
    g.extend(sysb, bounds, ids)
    o.extend(sysb, bounds, ids)
    boxes, rays = _boxes(dim, 150, 800 + kind), _rays(dim, 150, 900 + kind)
    offs, got = g.test_box_batch(sysb, boxes, max_depth)      # sorts implicitly, like Layer::test
    assert g.sorted and offs.shape == (151,) and offs[0] == 0 and offs[-1] == got.shape[0]
    for q in range(boxes.shape[0]):
        want = o.test_box(sysb, boxes[q], max_depth)
        mine = got[offs[q]:offs[q + 1]].astype(np.uint64)
        assert mine.shape == want.shape and (mine == want).all(), "box query %d" % q
    offs, got = g.test_ray_batch(sysb, rays, max_depth)
    for q in range(rays.shape[0]):
        want = o.test_ray(sysb, rays[q, :dim], rays[q, dim:2 * dim], rays[q, 2 * dim], rays[q, 2 * dim + 1], max_depth)
        mine = got[offs[q]:offs[q + 1]].astype(np.uint64)
        assert mine.shape == want.shape and (mine == want).all(), "ray query %d" % q
    assert g.stats()["launches"]["query"] >= 4


@pytest.mark.gpu
def test_gpu_queries_after_a_scan_and_edge_cases(bp):
    """After a scan the sorted tree carries cell flags in its IDs' top bits (dedup at the source): queries must
    not see them.  Also: empty layer, empty batch, single-query wrappers, a box that reports every record
    (a group far larger than the finish kernel's window -> the full-width fallback)."""
    sc = bp.scenes.uniform_cubes(50_000, 21)
    g = bp.Layer(2, "u32")
    o = co.OracleLayer(2, 4, 0)
    assert g.test_box(sc["sys_bounds"], [0, 0, 0, 1, 1, 1]).shape == (0,)
    g.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    o.extend(sc["sys_bounds"], sc["bounds"], sc["ids"])
    pairs = g.par_scan()
    assert pairs.shape[0] > 0
    offs, ids = g.test_box_batch(sc["sys_bounds"], np.zeros((0, 6), dtype=np.float32))
    assert offs.tolist() == [0] and ids.shape == (0,)
    box = np.array([0.2, 0.2, 0.2, 0.4, 0.35, 0.3], dtype=np.float32)
    assert (g.test_box(sc["sys_bounds"], box).astype(np.uint64) == o.test_box(sc["sys_bounds"], box)).all()
    everything = g.test_box(sc["sys_bounds"], [0, 0, 0, 1, 1, 1])
    assert (everything == np.arange(50_000, dtype=np.uint32)).all()
    ray = g.test_ray(sc["sys_bounds"], [0.5, 0.5, -1.0], [0.0, 0.0, 1.0], -np.inf, np.inf)
    want = o.test_ray(sc["sys_bounds"], [0.5, 0.5, -1.0], [0.0, 0.0, 1.0], -np.inf, np.inf)
    assert ray.shape[0] > 0 and (ray.astype(np.uint64) == want).all()
    # the scan still works afterwards and returns the same pairs
    assert (g.par_scan() == pairs).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("id_bytes", [4, 8])
@pytest.mark.parametrize("shape_kind", [0, 1])
def test_gpu_pick_ray_matches_the_oracle(bp, kind, id_bytes, shape_kind):
    dim = co.DIM[kind]
    sysb, bounds, ids, spheres = _pick_scene(kind, 5000, 60 + kind)
    shapes = spheres if shape_kind == bp.PICK_SPHERE else bounds     # (IDs are 0..n-1 here, so the table rows line up)
    g = bp.LayerBuilder().build(kind, "u32" if id_bytes == 4 else "u64")
    o = co.OracleLayer(kind, id_bytes, 0)
    g.extend(sysb, bounds, ids.astype(np.uint64 if id_bytes == 8 else np.uint32))
    o.extend(sysb, bounds, ids.astype(np.uint64 if id_bytes == 8 else np.uint32))
    rays = _pick_rays(dim, 300, 70 + kind, spheres[:, :dim])
    for max_dist, max_depth in ((np.inf, None), (2.0, None), (np.inf, 5)):
        res = g.pick_ray_batch(sysb, rays, max_dist, shape_kind, shapes, max_depth)
        nhit = 0
        for q in range(rays.shape[0]):
            want = o.pick_ray(sysb, rays[q, :dim], rays[q, dim:], max_dist, shape_kind, shapes, max_depth)
            if want is None:
                assert res[q]["hit"] == 0, q
                continue
            nhit += 1
            assert res[q]["hit"] == 1 and np.float32(res[q]["dist"]) == np.float32(want[0]) and int(res[q]["id"]) == want[1], q
            assert (res[q]["point"][:dim] == want[2]).all(), q
        assert nhit > (10 if np.isinf(max_dist) else 0)
    single = g.pick_ray(sysb, rays[3, :dim], rays[3, dim:], np.inf, shape_kind, shapes)
    want = o.pick_ray(sysb, rays[3, :dim], rays[3, dim:], np.inf, shape_kind, shapes)
    assert (single is None) == (want is None) and (single is None or (single[0] == want[0] and single[1] == want[1]))
