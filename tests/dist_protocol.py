"""An executable model of the sharded frame's protocol (product: csrc/bp_dist.cu behind bp_dist_frame, which follows this
choreography step by step) -- TEST INFRASTRUCTURE.  `ops` is a numpy test double of the device operations
(tests/dist_cpu_ops.py) and the collectives run over gloo, which lets world_size-2/3 CPU tests check the splitter / halo /
ownership / sort-plan / dedup-decision logic without GPUs.  One frame, on every rank r of g: see broadphase-rs_b200/dist.py.
"""
import numpy as np
import torch
import torch.distributed as dist

# (key bits, DIM, DEPTH_BITS, AXIS_BITS) -- reference src/index.rs:293-295
KIND_PARAMS = {0: (32, 2, 4, 14), 1: (64, 2, 5, 29), 2: (64, 3, 5, 19)}
SAMPLES_PER_RANK = 2048
U64_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)
N_TAGS = 7  # words a rank appends to its row of the record count matrix: id_or | fold bit, then CudaOps.sort_tags()
REBALANCE_AT = 1.15  # recompute cached splitters when the fullest shard exceeds the mean by this factor


def level_mask(kind, depth):
    _, dim, depth_bits, axis_bits = KIND_PARAMS[kind]
    if depth <= 0:
        return 0
    return ((1 << (dim * depth)) - 1) << (dim * axis_bits + depth_bits - dim * depth)


def run_upper_key(kind, key):
    """Largest key a record inside cell(key) can have (csrc/bp_common.cuh run_upper_key)."""
    _, dim, depth_bits, axis_bits = KIND_PARAMS[kind]
    used = (1 << (dim * axis_bits + depth_bits)) - 1
    depth = key & ((1 << depth_bits) - 1)
    return key | (~level_mask(kind, depth) & used)


def ancestor_keys(kind, key):
    """Keys of every cell that contains cell(key), the cell itself included, ascending."""
    _, _, depth_bits, _ = KIND_PARAMS[kind]
    depth = key & ((1 << depth_bits) - 1)
    return [(key & level_mask(kind, d)) | d for d in range(depth + 1)]


def choose_splitters(sample, parts):
    """parts-1 ascending splitters at the quantiles of a (host, uint64) sample; identical on every rank because the
    gathered sample is.  Any value is a valid splitter (it only decides the balance), so each one is moved to the roundest
    value (most trailing zero bits) whose sample rank stays within 1/32 of a shard's size of its quantile: the keys of a
    shard then share their top bits (shard_fixed_bits) and the shard's sort never looks at them."""
    s = np.sort(np.asarray(sample, dtype=np.uint64))
    m = s.shape[0]
    if m == 0:
        return np.full(parts - 1, U64_MAX, dtype=np.uint64)
    slack = m // parts // 32
    q = []
    for i in range(1, parts):
        t = min(m - 1, (i * m) // parts)
        v = int(s[t])
        if slack:
            lo, hi = int(s[t - min(t, slack)]), int(s[min(m - 1, t + slack)])
            if lo < hi:  # in (lo, hi]: hi without the bits below the first one in which the two differ
                v = hi & ~((1 << ((lo ^ hi).bit_length() - 1)) - 1)
        q.append(v)
    return np.asarray(q, dtype=np.uint64)


def shard_fixed_bits(splitters, me, top=0xFFFFFFFFFFFFFFFF):
    """(fixed, value): the bit positions every key of shard `me` is known to share, from its two splitters alone, and the
    bits themselves.  A key v belongs to shard d iff splitters[d-1] <= v < splitters[d]; `top` = the largest possible key."""
    parts = len(splitters) + 1
    full = 0xFFFFFFFFFFFFFFFF
    lo = int(splitters[me - 1]) if me > 0 else 0
    if (me < parts - 1 and int(splitters[me]) <= lo) or lo > top:
        return 0, 0  # an empty shard
    hi = min(int(splitters[me]) - 1, top) if me < parts - 1 else top
    b = (lo ^ hi).bit_length()
    fixed = (full << b) & full
    return fixed, lo & fixed


def sort_plan(tags, id_or, n_halo, shard=None):
    """(key_or, key_and, id_or, id_and, ids_ascending) for the sort of a receive buffer, from the tag words every source
    sent with its counts (N_TAGS per source: id_or | fold bit, key_or, key_and, id_and, first ID, last ID, ascending).
    The buffer holds the sources' chunks in rank order, each a stable partition of the source's records: its IDs ascend
    iff every source's do, the sources' ID ranges follow each other in rank order, and no (unordered) halo copies came.
    shard = (splitters, me) of the receiving shard: without halo copies (they lie below the lower splitter) every key
    carries the bits the shard's two ends share (shard_fixed_bits; the last shard ends at the largest key the sources' OR
    allows), whatever the sources' masks say about their whole key sets."""
    full = 0xFFFFFFFFFFFFFFFF
    key_or, key_and, id_and = 0, full, full
    ascending, prev_last = n_halo == 0, -1
    for t in tags:
        key_or |= t[1]
        key_and &= t[2]
        id_and &= t[3]
        first, last, asc = t[4], t[5], t[6]
        if first > last and asc:  # an empty source
            continue
        ascending = ascending and bool(asc) and first >= prev_last
        prev_last = max(prev_last, last)
    if n_halo == 0 and shard is not None and len(shard[0]):
        fixed, value = shard_fixed_bits(shard[0], shard[1], (1 << key_or.bit_length()) - 1)
        key_or &= ~fixed | value
        key_and |= value
    return key_or, key_and, id_or, id_and, ascending


def chunk_offsets(m_own, m_halo, me):
    """Where this rank's chunks start inside every destination's receive buffer.  The buffer of
    destination d is laid out source by source: [owned from 0 | halo from 0 | owned from 1 | ...]."""
    own_off = (m_own[:me] + m_halo[:me]).sum(axis=0)
    return own_off.tolist(), (own_off + m_own[me]).tolist()


def scatter_destinations(key_ptrs, id_ptrs, m_own, m_halo, me):
    """Device addresses (uint64 arrays, one entry per destination rank) at which this rank's owned chunk and its halo
    chunk start inside every receive buffer: (keys, ids, halo keys, halo ids); the halo arrays are None when no halo
    copy leaves this rank.  key_ptrs / id_ptrs: base addresses of every rank's receive buffers (uint64 arrays).  This runs
    between the host's look at the count matrix and the launch of the scatter, with the GPU idle: array arithmetic, no
    Python loops."""
    own_off = (m_own[:me] + m_halo[:me]).sum(axis=0).astype(np.uint64)
    dk = key_ptrs + np.uint64(8) * own_off
    di = id_ptrs + np.uint64(4) * own_off
    if not m_halo[me].any():
        return dk, di, None, None
    halo_off = own_off + m_own[me].astype(np.uint64)
    return dk, di, key_ptrs + np.uint64(8) * halo_off, id_ptrs + np.uint64(4) * halo_off


class DistLayer:
    """The distributed counterpart of clear -> extend -> par_sort -> par_scan(_filtered) for one frame."""

    def __init__(self, ops, kind, group=None, trace=False, reuse_splitters=True, global_dedup_decision=True):
        self.ops, self.kind, self.group = ops, kind, group
        self.global_dedup_decision = global_dedup_decision  # (False only in a test that shows what the decision prevents)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.trace = trace  # per-phase wall times (device-synchronised) in self.last["phases_ms"]; for tuning only
        self.reuse_splitters = reuse_splitters
        self.fuse_counts = True  # frames with cached splitters: counts taken by the encode kernel (CudaOps.encode_count_matrix)
        self._splitters = self._a_splitters = None
        self._id_mask = 0
        self._static_halo = None  # halo records at the front of the resident static shard (None: no static layer)
        self._static_id_bits = 0
        self.last = {}

    # -- small collectives (device tensors over NCCL in production, CPU tensors over gloo in the tests) --
    def _all_gather(self, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.stack(out)

    def _gather_rows(self, row, device):
        t = torch.tensor([int(c) for c in row], dtype=torch.int64, device=device)
        return self._all_gather(t).cpu().numpy()  # [source, ...]

    @staticmethod
    def _imbalance(col_sums):
        mean = float(np.mean(col_sums))
        return float(np.max(col_sums)) / mean if mean > 0 else 1.0

    def set_static(self, sys_bounds, bounds, ids, n):
        """Shards a static scene once (config 4 at N > 1): its records are range-partitioned with splitters
        sampled from the static keys -- which stay FIXED from then on, so every frame's dynamic records are
        routed to the same owners -- sorted, and kept resident.  frame() then merges them in
        (Layer::merge, src/layer.rs:127-138) before the scan.  Halo copies of static records sit at the
        front of the static shard, exactly like those of the dynamic records."""
        ops, g, me = self.ops, self.world, self.rank
        dev = ops.device
        keys, rids, id_or = ops.encode(sys_bounds, bounds, ids, n)
        m = SAMPLES_PER_RANK
        meta = torch.full((m + 1,), -1, dtype=torch.int64, device=dev)
        if keys.shape[0]:
            ks = keys[::max(1, keys.shape[0] // m)][:m]
            meta[:ks.shape[0]] = ks
        meta[m] = id_or
        gathered = self._all_gather(meta).cpu().numpy()
        sample = gathered[:, :m].reshape(-1)
        self._splitters = choose_splitters(sample[sample >= 0].view(np.uint64), g)
        for v in gathered[:, m]:
            self._static_id_bits |= int(v)
        counts, halo = ops.count_records(keys, self._splitters)
        mat = self._gather_rows(counts + halo, dev)
        m_own, m_halo = mat[:, :g], mat[:, g:2 * g]
        rk, ri = ops.exchange_records(keys, rids, self._splitters, m_own, m_halo)
        self._static_halo = int(m_halo[:, me].sum())
        n_static = ops.keep_static(rk, ri)
        return n_static

    def frame(self, sys_bounds, bounds, ids, n, flt=None):
        """Runs one frame on this rank's objects.  Returns the rank's slice of the globally sorted,
        deduplicated pair list as an (P_r, 2) int32 tensor (bit patterns of the u32 IDs)."""
        ops, g, me = self.ops, self.world, self.rank
        dev = ops.device
        phases = []

        def mark(name):
            if self.trace:
                import time
                if dev.type == "cuda":
                    torch.cuda.synchronize(dev)
                phases.append((name, time.perf_counter()))

        mark("start")
        # 1. encode -- together with step 3's counts when the splitters are already known (cached from the last frame)
        m = SAMPLES_PER_RANK
        product = hasattr(ops, "count_records_matrix")
        need_splitters = self._splitters is None or (not self.reuse_splitters and self._static_halo is None)
        fused = product and not need_splitters and self.fuse_counts and ops.enc.min_depth == 0
        mat = None
        if fused:
            keys, rids, mat = ops.encode_count_matrix(sys_bounds, bounds, ids, n, self._splitters, self._static_halo is None)
            id_or = None
        else:
            keys, rids, id_or = ops.encode(sys_bounds, bounds, ids, n)
        r_loc = keys.shape[0]
        mark("encode")

        # 2. key splitters (+ the ID bits, piggybacked) from an all-gathered sample
        if need_splitters:
            meta = torch.full((m + 1,), -1, dtype=torch.int64, device=dev)
            if r_loc:
                ks = keys[::max(1, r_loc // m)][:m]
                meta[:ks.shape[0]] = ks
            meta[m] = id_or
            gathered = self._all_gather(meta).cpu().numpy()
            sample = gathered[:, :m].reshape(-1)
            self._splitters = choose_splitters(sample[sample >= 0].view(np.uint64), g)
            id_bits = 0
            for v in gathered[:, m]:
                id_bits |= int(v)
            self._id_mask = (1 << max(1, id_bits.bit_length())) - 1
        splitters = self._splitters
        mark("splitters")

        # 3. count, all-gather the count matrix, scatter straight into the owners' receive buffers
        # bit 63 of the tag: "my IDs leave their top 3 bits free" (dedup at the source across the exchange, below)
        can_fold = product and not fused and self._static_halo is None and int(id_or) < (1 << 29)
        if fused:    # (already there)
            pass
        elif product:  # counts stay on the device, the matrix travels over NVLink
            mat = ops.count_records_matrix(keys, splitters, [int(id_or) | ((1 << 63) if can_fold else 0)] + ops.sort_tags())
        else:        # CPU test double: host counts + all_gather (gloo)
            counts, halo = ops.count_records(keys, splitters)
            mat = self._gather_rows(counts + halo + [id_or], dev)
        m_own, m_halo = mat[:, :g], mat[:, g:2 * g]
        n_halo = int(m_halo[:, me].sum())
        tags = np.ascontiguousarray(mat[:, 2 * g:]).view(np.uint64).tolist()  # Python ints, unsigned
        flagged, id_bits, plan = product, 0, None
        for t in tags:
            flagged = flagged and bool(t[0] >> 63)  # every rank can: the cell flags ride in the IDs across the exchange
            id_bits |= t[0] & ~(1 << 63)
        if product:
            plan = sort_plan(tags, id_bits, n_halo, (splitters, me))
        id_bits |= self._static_id_bits
        self._id_mask |= (1 << max(1, id_bits.bit_length())) - 1  # IDs seen since the splitters were cached
        mark("counts")
        if product:
            rk, ri = ops.exchange_records(keys, rids, splitters, m_own, m_halo, fold=flagged)
        else:
            rk, ri = ops.exchange_records(keys, rids, splitters, m_own, m_halo)
        mark("exchange")

        # 4. local sort: the halo records (all < my lower splitter) end up in front
        sk, si = ops.sort_records(rk, ri, flagged, plan) if product else ops.sort_records(rk, ri)
        if self._static_halo is not None:  # Layer::merge of the resident static shard (sorted runs: merge path)
            ops.merge_static()
            n_halo += self._static_halo
        mark("sort")

        # 5. shard-local scan; pairs whose later record is a halo record belong to an earlier shard
        raw = ops.scan_raw(sk, si, n_halo, flt)
        p_raw = raw.shape[0]
        mark("scan")

        # 6. global dedup: range-partition the raw pairs on the later ID, scatter, sort + unique
        if self._a_splitters is None or not self.reuse_splitters:
            ps = torch.full((m,), -1, dtype=torch.int64, device=dev)
            if p_raw:
                a = (raw[::max(1, p_raw // m)][:m] >> 32) & 0xFFFFFFFF
                ps[:a.shape[0]] = a
            gathered = self._all_gather(ps).cpu().numpy().reshape(-1)
            self._a_splitters = choose_splitters(gathered[gathered >= 0].astype(np.uint64), g)
        a_splitters = self._a_splitters
        if hasattr(ops, "count_pairs_matrix"):
            # Dedup at the source (every ID pair emitted from its canonical shared cell only) is valid only while NO record
            # of the whole scene is inactive: the shard holding a pair's canonical cell skips it there if that record's ID
            # owns an enclosing bound (src/layer.rs:562-564), and the reference then reports the pair from another shared
            # cell -- possibly in a neighbouring shard, which must not have suppressed its copy.  A shard knows only its
            # own records, so the flag travels with the pair counts, and when ANY shard saw an inactive record, every
            # shard whose scan ran with the dedup scans again without it (a rare path: IDs owning nested bounds).
            same = bool(getattr(ops, "saw_same_id", False))
            pm, seen = ops.count_pairs_matrix(raw, a_splitters, int(same))
            if flagged and self.global_dedup_decision and bool(np.any(seen != 0)):
                if not same and n_halo == 0:
                    raw = ops.scan_raw(sk, si, n_halo, flt, dedup=False)
                    p_raw = raw.shape[0]
                pm, _ = ops.count_pairs_matrix(raw, a_splitters, int(same))
        else:
            pc = ops.count_pairs(raw, a_splitters)
            pm = self._gather_rows(pc, dev)
        mark("pair_counts")
        rp = ops.exchange_pairs(raw, a_splitters, pm)
        mark("pair_exchange")
        if g > 1 and hasattr(ops, "set_pair_later_fixed"):  # my later IDs lie between two pair splitters: no radix pass on their top bits
            ops.set_pair_later_fixed(*shard_fixed_bits(a_splitters, me, self._id_mask))
        pairs = ops.unique_pairs(rp, self._id_mask)
        mark("unique")

        # cached splitters are recomputed next frame when a shard has drifted too far from the mean
        if self.reuse_splitters:
            if self._static_halo is None and self._imbalance((m_own + m_halo).sum(axis=0)) > REBALANCE_AT:
                self._splitters = None  # (with a static layer the record splitters are fixed)
            if self._imbalance(pm.sum(axis=0)) > REBALANCE_AT:
                self._a_splitters = None
        phases_ms = {b[0]: (b[1] - a_[1]) * 1e3 for a_, b in zip(phases[:-1], phases[1:])}
        self.last = dict(phases_ms=phases_ms, records_local=r_loc, records_owned=int(m_own[:, me].sum()), halo=n_halo,
                         raw_pairs=int(p_raw), pairs=int(pairs.shape[0]), record_matrix=m_own, halo_matrix=m_halo,
                         pair_matrix=pm)
        return pairs

    def gather_pairs(self, pairs):
        """Concatenates every rank's slice in rank order: the reference's scan() vector (host numpy)."""
        cnt = self._all_gather(torch.tensor([pairs.shape[0]], dtype=torch.int64, device=pairs.device)).cpu().numpy().reshape(-1)
        mx = int(cnt.max()) if cnt.size else 0
        buf = torch.zeros((mx, 2), dtype=pairs.dtype, device=pairs.device)
        buf[:pairs.shape[0]] = pairs
        allp = self._all_gather(buf).cpu().numpy()
        return np.concatenate([allp[r, :int(cnt[r])] for r in range(self.world)], axis=0).view(np.uint32)
