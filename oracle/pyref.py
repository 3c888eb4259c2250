"""pyref.py -- an independent numpy restatement of the broadphase-rs hot path.

TEST INFRASTRUCTURE ONLY.  It exists to cross-check the C++ oracle (bp_oracle.cpp) with a second,
independent formulation.  Parity status: pinned -- like the C++ oracle it reproduces the reference's three
validation files (SHA-256 in the LFS pointers of tests/data/validation/) from the regenerated n = 10 000
gen_boxes scene (tests/test_reference_fixtures.py); the codec and quantiser are pinned by the reference's KATs.

It deliberately uses different formulations from the C++ oracle:
  * Morton spreading is a plain bit loop (the definition "axis bit i -> origin bit DIM*i + axis",
    reference src/index.rs:155-172, 193-207, 230-250);
  * the scan is the stack-free closed form (each record looks up the equal-key runs of its
    ancestor cells) instead of the literal stack sweep of src/layer.rs:550-573.
All file:line citations are relative to /root/reference.
"""
import numpy as np

# kind -> (key bits, DIM, DEPTH_BITS, AXIS_BITS) -- src/index.rs:293-295
KINDS = {
    0: (32, 2, 4, 14),  # Index32_2D
    1: (64, 2, 5, 29),  # Index64_2D
    2: (64, 3, 5, 19),  # Index64_3D
}
INDEX32_2D, INDEX64_2D, INDEX64_3D = 0, 1, 2

FILTER_NONE, FILTER_ID_PARITY, FILTER_XOR_MASK, FILTER_CATEGORY, FILTER_SPHERES = 0, 1, 2, 3, 4

RANGE = np.float32(4294967040.0)  # 0xffff_ff00 as f32 -- src/geom.rs:152-154


def encode_axis(kind, v):
    """src/index.rs:155-172 / 193-207 as a bit loop. v: uint32 array -> python-int-safe uint64."""
    _, dim, _, axis_bits = KINDS[kind]
    top = (np.asarray(v, dtype=np.uint64) >> np.uint64(32 - axis_bits))
    out = np.zeros_like(top)
    for i in range(axis_bits):
        out |= ((top >> np.uint64(i)) & np.uint64(1)) << np.uint64(dim * i)
    return out


def decode_axis(kind, origin_bits):
    """src/index.rs:134-151 / 176-190 as a bit loop."""
    _, dim, _, axis_bits = KINDS[kind]
    o = np.asarray(origin_bits, dtype=np.uint64)
    out = np.zeros_like(o)
    for i in range(axis_bits):
        out |= ((o >> np.uint64(dim * i)) & np.uint64(1)) << np.uint64(i)
    return (out << np.uint64(32 - axis_bits)).astype(np.uint32)


def level_mask(kind, depth):
    """src/index.rs:82-86 (python ints)."""
    _, dim, depth_bits, axis_bits = KINDS[kind]
    if depth <= 0:
        return 0
    return ((1 << (dim * depth)) - 1) << (dim * axis_bits + depth_bits - dim * depth)


def make_index(kind, depth, origin):
    """Index::default().set_depth(depth).set_origin(origin) -- src/index.rs:106-112, 230-250.
    depth: uint32 array (n,), origin: uint32 array (n, DIM) -> uint64 array."""
    bits, dim, depth_bits, axis_bits = KINDS[kind]
    depth = np.minimum(np.asarray(depth, dtype=np.uint64), np.uint64(axis_bits))
    o = np.zeros(depth.shape, dtype=np.uint64)
    for a in range(dim):
        o |= encode_axis(kind, origin[:, a]) << np.uint64(a)
    origin_mask = np.uint64((((1 << (dim * axis_bits)) - 1) << depth_bits) & ((1 << 64) - 1))
    key = ((o << np.uint64(depth_bits)) & origin_mask) | (depth & np.uint64((1 << depth_bits) - 1))
    if bits == 32:
        key &= np.uint64(0xFFFFFFFF)
    return key


def f32_as_u32(x):
    """Rust `as u32`: truncating, saturating, NaN -> 0."""
    x = np.asarray(x, dtype=np.float32)
    y = np.where(np.isnan(x), np.float32(0), x)
    y = np.clip(y.astype(np.float64), 0.0, 4294967295.0)
    return np.floor(y).astype(np.uint64).astype(np.uint32)


def to_local(sys_bounds, bounds, dim):
    """SystemBounds::to_local -- src/geom.rs:148-163. bounds (n, 2*dim) f32 -> (n, 2*dim) u32."""
    sys_bounds = np.asarray(sys_bounds, dtype=np.float32)
    bounds = np.asarray(bounds, dtype=np.float32).reshape(-1, 2 * dim)
    smin = sys_bounds[:dim]
    size = (sys_bounds[dim:] - smin).astype(np.float32)
    out = np.zeros(bounds.shape, dtype=np.uint32)
    with np.errstate(all="ignore"):
        for half in (0, 1):
            g = bounds[:, half * dim:(half + 1) * dim]
            t = (g - smin).astype(np.float32)
            q = (t / size).astype(np.float32)
            m = (q * RANGE).astype(np.float32)
            r = (m + np.float32(0.0)).astype(np.float32)
            out[:, half * dim:(half + 1) * dim] = f32_as_u32(r)
    return out


def to_global(sys_bounds, local, dim):
    """SystemBounds::to_global -- src/geom.rs:165-181."""
    sys_bounds = np.asarray(sys_bounds, dtype=np.float32)
    local = np.asarray(local, dtype=np.uint32).reshape(-1, 2 * dim)
    smin = sys_bounds[:dim]
    size = (sys_bounds[dim:] - smin).astype(np.float32)
    out = np.zeros(local.shape, dtype=np.float32)
    for half in (0, 1):
        l = local[:, half * dim:(half + 1) * dim].astype(np.float32)
        q = (l / RANGE).astype(np.float32)
        m = (q * size).astype(np.float32)
        out[:, half * dim:(half + 1) * dim] = (smin + m).astype(np.float32)
    return out


def contains(sys_bounds, bounds, dim):
    """Bounds::contains -- src/geom.rs:121-128 (NaN passes)."""
    sys_bounds = np.asarray(sys_bounds, dtype=np.float32)
    bounds = np.asarray(bounds, dtype=np.float32).reshape(-1, 2 * dim)
    with np.errstate(all="ignore"):
        bad = (sys_bounds[:dim] > bounds[:, :dim]) | (sys_bounds[dim:] < bounds[:, dim:])
    return ~bad.any(axis=1)


def extend(kind, min_depth, sys_bounds, bounds, ids):
    """Layer::extend -- src/layer.rs:94-121 + src/geom.rs:189-304.
    Returns (keys uint64, ids) in the reference's record order: object order, z outer, y, x inner."""
    _, dim, _, axis_bits = KINDS[kind]
    bounds = np.asarray(bounds, dtype=np.float32).reshape(-1, 2 * dim)
    ids = np.asarray(ids)
    ok = contains(sys_bounds, bounds, dim)
    bounds, ids = bounds[ok], ids[ok]
    n = bounds.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.uint64), ids[:0].copy()
    loc = to_local(sys_bounds, bounds, dim)
    lmin, lmax = loc[:, :dim], loc[:, dim:]
    sizei = (lmax - lmin + np.uint32(1)).astype(np.uint32)  # wrapping u32
    m = (sizei.max(axis=1) - np.uint32(1)).astype(np.uint32)
    # leading_zeros, as 32 - bit_length
    bit_length = np.zeros(n, dtype=np.uint32)
    t = m.copy()
    for _ in range(32):
        bit_length += (t != 0)
        t >>= np.uint32(1)
    depth = (np.uint32(32) - bit_length).astype(np.uint32)
    depth = np.maximum(depth, np.uint32(min_depth))
    depth = np.minimum(depth, np.uint32(axis_bits))

    d0 = depth == 0
    shift = np.where(d0, 0, 32 - depth.astype(np.int64)).astype(np.uint64)
    step = np.where(d0, np.uint64(1) << np.uint64(32), np.uint64(1) << shift)  # u64 to hold 2^32
    tmask = (~(step - np.uint64(1))) & np.uint64(0xFFFFFFFF)
    tmin = (lmin.astype(np.uint64) & tmask[:, None])
    tmax = (lmax.astype(np.uint64) & tmask[:, None])
    cnt = np.where(tmax > tmin, (tmax - tmin) // step[:, None] + np.uint64(1), np.uint64(1)).astype(np.int64)
    cnt[d0] = 1
    if dim == 2:
        cnt = np.concatenate([cnt, np.ones((n, 1), dtype=np.int64)], axis=1)
        tmin = np.concatenate([tmin, np.zeros((n, 1), dtype=np.uint64)], axis=1)
    per_obj = cnt[:, 0] * cnt[:, 1] * cnt[:, 2]
    total = int(per_obj.sum())
    obj = np.repeat(np.arange(n), per_obj)
    first = np.cumsum(per_obj) - per_obj
    local = np.arange(total) - np.repeat(first, per_obj)  # index within object: x fastest
    cx, cy = cnt[obj, 0], cnt[obj, 1]
    ix = local % cx
    iy = (local // cx) % cy
    iz = local // (cx * cy)
    st = step[obj]
    origin = np.zeros((total, 3), dtype=np.uint64)
    origin[:, 0] = tmin[obj, 0] + ix.astype(np.uint64) * st
    origin[:, 1] = tmin[obj, 1] + iy.astype(np.uint64) * st
    origin[:, 2] = tmin[obj, 2] + iz.astype(np.uint64) * st
    origin[d0[obj]] = 0
    keys = make_index(kind, depth[obj], origin.astype(np.uint32))
    keys[d0[obj]] = 0
    return keys, ids[obj]


def sort_records(keys, ids):
    """Layer::sort -- src/layer.rs:146-165: lexicographic (Index, ID)."""
    order = np.lexsort((ids, keys))
    return keys[order], ids[order]


def _filter(kind, arg, table, a, b):
    if kind == FILTER_NONE:
        return np.ones(a.shape, dtype=bool)
    if kind == FILTER_ID_PARITY:
        return ((a ^ b) & 1) == 1
    if kind == FILTER_XOR_MASK:
        return ((a ^ b) & np.uint64(arg)) != 0
    if kind == FILTER_CATEGORY:
        table = np.asarray(table, dtype=np.uint32).reshape(-1, 2)
        n = table.shape[0]

        def look(x, col):
            out = np.full(x.shape, 0xFFFFFFFF, dtype=np.uint32)
            inr = x < n
            out[inr] = table[x[inr].astype(np.int64), col]
            return out
        return ((look(a, 0) & look(b, 1)) != 0) & ((look(b, 0) & look(a, 1)) != 0)
    if kind == FILTER_SPHERES:  # the narrow phase of examples/main.rs:461-479 in float32, one rounding per operation
        t = np.asarray(table, dtype=np.float32).reshape(-1, 4)
        n = t.shape[0]
        inr = (a < n) & (b < n)
        ia, ib = np.where(inr, a, 0).astype(np.int64), np.where(inr, b, 0).astype(np.int64)
        d = (t[ib, :3] - t[ia, :3]).astype(np.float32)
        sq = (d * d).astype(np.float32)
        m = ((sq[:, 0] + sq[:, 1]).astype(np.float32) + sq[:, 2]).astype(np.float32)
        return ~inr | ~(np.sqrt(m, dtype=np.float32) > (t[ia, 3] + t[ib, 3]).astype(np.float32))
    raise ValueError(kind)


def scan_raw(kind, keys, ids, filter_kind=FILTER_NONE, filter_arg=0, table=None, first_owned=0):
    """The raw (later, earlier) pairs of Layer::scan_filtered -- src/layer.rs:456-477 -- in closed
    form, before the sort + dedup, as two uint64 arrays.

    keys/ids must already be sorted.  For record j the stack of src/layer.rs:550-573 holds exactly
    the pushed earlier records whose cell contains cell(j); those are, for every depth
    d <= depth(j), the run of records with key == (key_j & level_mask(d)) | d that lie before j.
    first_owned > 0: records below that index are halo records of a multi-GPU shard -- they act as
    ancestors but never as the later record of a pair."""
    _, dim, depth_bits, axis_bits = KINDS[kind]
    keys = np.asarray(keys, dtype=np.uint64)
    ids64 = np.asarray(ids).astype(np.uint64)
    n = keys.shape[0]
    empty = np.zeros(0, dtype=np.uint64)
    if n == 0:
        return empty, empty
    depth = (keys & np.uint64((1 << depth_bits) - 1)).astype(np.int64)
    j_all = np.arange(n)
    src_i, src_j = [], []
    for d in range(0, axis_bits + 1):
        sel = j_all[depth >= d]
        if sel.size == 0:
            continue
        anc = (keys[sel] & np.uint64(level_mask(kind, d))) | np.uint64(d)
        lo = np.searchsorted(keys, anc, side="left")
        hi = np.minimum(np.searchsorted(keys, anc, side="right"), sel)  # strictly earlier records
        cnt = np.maximum(hi - lo, 0)
        tot = int(cnt.sum())
        if tot == 0:
            continue
        jj = np.repeat(sel, cnt)
        first = np.cumsum(cnt) - cnt
        ii = np.repeat(lo, cnt) + (np.arange(tot) - np.repeat(first, cnt))
        src_i.append(ii)
        src_j.append(jj)
    if not src_i:
        return empty, empty
    ii = np.concatenate(src_i)
    jj = np.concatenate(src_j)
    inactive = np.zeros(n, dtype=bool)
    inactive[jj[ids64[ii] == ids64[jj]]] = True
    keep = ~inactive[ii] & ~inactive[jj] & (jj >= first_owned)
    ii, jj = ii[keep], jj[keep]
    a, b = ids64[jj], ids64[ii]
    f = _filter(filter_kind, filter_arg, table, a, b)
    return a[f], b[f]


def scan(kind, keys, ids, filter_kind=FILTER_NONE, filter_arg=0, table=None):
    """Layer::scan_filtered: (pairs (P, 2) uint64 sorted + unique, raw pair count)."""
    a, b = scan_raw(kind, keys, ids, filter_kind, filter_arg, table)
    raw = int(a.shape[0])
    pairs = np.stack([a, b], axis=1)
    if raw:
        pairs = np.unique(pairs, axis=0)  # lexicographic sort + dedup
    return pairs, raw


# ---------------------------------------------------------------------------------------------
# Queries -- Layer::test_box / test_ray (src/layer.rs:244-351) in closed form.
#
# test_impl (src/layer.rs:167-242) reports a record exactly when the test geometry, subdivided along
# the record's own path of cells, passes should_test at every level from the root down to
# min(depth, max_depth): the descent reaches the record's cell (or its max_depth ancestor) and
# reports the slice it sits in.  So instead of walking the hierarchy, every record replays its own
# path -- vectorised over all records, one numpy step per level.  The f32 arithmetic is the
# reference's: centre = min + (max - min) / 2 (cgmath 0.17 EuclideanSpace::midpoint behind
# Bounds::center, src/geom.rs:130-132), ray distances = (centre - origin) / direction.
# PARITY UNPINNED: the reference has no test for its queries; this is cross-checked against the
# literal recursion in bp_oracle.cpp only.
# ---------------------------------------------------------------------------------------------
def _query(kind, keys, ids, sys_bounds, max_depth, geom):
    bits, dim, depth_bits, axis_bits = KINDS[kind]
    keys = np.asarray(keys, dtype=np.uint64)
    ids = np.asarray(ids, dtype=np.uint64)
    n = keys.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    total = dim * axis_bits + depth_bits
    depth = np.minimum((keys & np.uint64((1 << depth_bits) - 1)).astype(np.int64), axis_bits)
    eff = depth if max_depth is None else np.minimum(depth, int(max_depth))
    sysb = np.asarray(sys_bounds, dtype=np.float32)
    cmin = np.repeat(sysb[None, :dim], n, axis=0).copy()
    cmax = np.repeat(sysb[None, dim:], n, axis=0).copy()
    state = geom["init"](n, cmin, cmax)
    ok = geom["should_test"](cmin, cmax, state)
    two = np.float32(2.0)
    with np.errstate(all="ignore"):
        for level in range(1, int(eff.max()) + 1 if n else 1):
            active = eff >= level
            digit = ((keys >> np.uint64(total - dim * level)) & np.uint64((1 << dim) - 1)).astype(np.int64)
            center = (cmin + (cmax - cmin) / two).astype(np.float32)
            side = np.stack([((digit >> a) & 1) == 1 for a in range(dim)], axis=1)
            new_state = geom["child"](center, side, state)
            new_min = np.where(side, center, cmin)
            new_max = np.where(side, cmax, center)
            cmin = np.where(active[:, None], new_min, cmin)
            cmax = np.where(active[:, None], new_max, cmax)
            state = [np.where(active, ns, s) for ns, s in zip(new_state, state)]
            ok &= (~active) | geom["should_test"](cmin, cmax, state)
    return np.unique(ids[ok])


def test_box(kind, keys, ids, sys_bounds, test_bounds, max_depth=None):
    """Layer::test_box -- src/layer.rs:293-311 with BoxTestGeometry, src/geom.rs:353-460."""
    dim = KINDS[kind][1]
    tb = np.asarray(test_bounds, dtype=np.float32)
    tmin, tmax = tb[:dim], tb[dim:]
    geom = {
        "init": lambda n, cmin, cmax: [],
        "child": lambda center, side, state: [],
        # cell_bounds.overlaps(test_bounds) -- src/geom.rs:104-111
        "should_test": lambda cmin, cmax, state: ~((cmin > tmax[None, :]) | (cmax < tmin[None, :])).any(axis=1),
    }
    return _query(kind, keys, ids, sys_bounds, max_depth, geom)


def test_ray(kind, keys, ids, sys_bounds, origin, direction, range_min, range_max, max_depth=None):
    """Layer::test_ray -- src/layer.rs:326-351 with RayTestGeometry, src/geom.rs:462-615."""
    dim = KINDS[kind][1]
    org = np.asarray(origin, dtype=np.float32)
    dirn = np.asarray(direction, dtype=np.float32)
    sysb = np.asarray(sys_bounds, dtype=np.float32)
    rmin, rmax = np.float32(range_min), np.float32(range_max)
    with np.errstate(all="ignore"):  # with_system_bounds -- src/geom.rs:512-535
        for axis in range(dim):
            d_lo = np.float32(np.float32(sysb[axis] - org[axis]) / dirn[axis])
            d_hi = np.float32(np.float32(sysb[dim + axis] - org[axis]) / dirn[axis])
            d0, d1 = (d_lo, d_hi) if dirn[axis] > 0 else (d_hi, d_lo)
            if np.isfinite(d0):
                rmin = np.fmax(rmin, d0)
            if np.isfinite(d1):
                rmax = np.fmin(rmax, d1)

    def child(center, side, state):  # subdivide -- src/geom.rs:537-577
        lo, hi = state[0].copy(), state[1].copy()
        for axis in range(dim):
            dist = ((center[:, axis] - org[axis]) / dirn[axis]).astype(np.float32)
            fin = np.isfinite(dist)
            towards = (dirn[axis] > 0) != side[:, axis]
            hi = np.where(fin & towards, np.fmin(hi, dist), hi)
            lo = np.where(fin & ~towards, np.fmax(lo, dist), lo)
            kill = ~fin & ((org[axis] > center[:, axis]) != side[:, axis])
            lo = np.where(kill, np.float32(np.inf), lo)
            hi = np.where(kill, np.float32(-np.inf), hi)
        return [lo.astype(np.float32), hi.astype(np.float32)]

    geom = {
        "init": lambda n, cmin, cmax: [np.full(n, rmin, dtype=np.float32), np.full(n, rmax, dtype=np.float32)],
        "child": child,
        # should_test(nearest = inf) -- src/geom.rs:612-614
        "should_test": lambda cmin, cmax, state: (state[0] < state[1]) & (state[0] < np.float32(np.inf)),
    }
    return _query(kind, keys, ids, sys_bounds, max_depth, geom)
