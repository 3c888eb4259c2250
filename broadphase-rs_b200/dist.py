"""Multi-GPU broadphase: Morton-prefix range sharding, one process per GPU.

The reference is single-process (Rayon on one host); this is the B200-native way to scale its hot
path over NVLink.  One frame, on every rank r of g:

  1. encode     the rank's own objects -> unsorted (Index, ID) records                       [K1, local]
  2. splitters  g-1 key splitters from an all-gathered regular sample of the local keys (sample sort).
                They are kept across frames while the shards stay balanced (objects move little from
                frame to frame) and recomputed when the imbalance exceeds 15 %.            [all_gather, tiny]
  3. exchange   every rank counts its records per destination shard (+ halo copies, below) and the count
                matrix is all-gathered, so every rank knows where its records go inside every receive
                buffer; then ONE partition pass (the onesweep radix pass with a splitter-search digit)
                writes each record straight into the destination GPU's receive buffer through NVLink
                peer pointers (torch symmetric memory): the pack kernel IS the all-to-all.     [fused, NVLink]
     halos      by the contiguity lemma a record of an earlier shard can only be an ancestor
                of something in shard s if its cell reaches past s's lower splitter S_s, i.e. if
                run_upper_key(key) >= S_s.  Such records (scene-sized objects; usually none) are sent to
                s as well, by the same pass.  All of them sort before S_s, so after the local sort they are
                exactly the first n_halo records of the shard.
  4. sort       the received records                                                         [K2, local]
  5. scan       over [halo | owned]; only pairs whose LATER record is owned are emitted, so every raw
                pair is produced exactly once globally                                       [K3, local]
  6. dedup      the same ID pair can be produced in several shards, and the reference returns one
                globally sorted vector: raw pairs are range-partitioned on the later ID (splitters from a
                sample of the raw pairs) and scattered to their owners the same way, then sorted +
                deduplicated per rank.  Concatenating the ranks' results in rank order is exactly the
                reference's scan() output.                                                   [fused, NVLink; K4]

The choreography below is independent of where the local operations run: `ops` is CudaOps (the
product: every operation is a C-ABI call into libbroadphase_b200.so on device tensors) or, in the CPU
tests only, a numpy test double with gloo -- which lets world_size-2 tests check the splitter / halo /
ownership logic without GPUs.  32-bit IDs only.
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
    """parts-1 ascending splitters at the quantiles of a (host, uint64) sample; identical on every
    rank because the gathered sample is."""
    s = np.sort(np.asarray(sample, dtype=np.uint64))
    if s.shape[0] == 0:
        return np.full(parts - 1, U64_MAX, dtype=np.uint64)
    q = [s[min(s.shape[0] - 1, (i * s.shape[0]) // parts)] for i in range(1, parts)]
    return np.asarray(q, dtype=np.uint64)


def sort_plan(tags, id_or, n_halo):
    """(key_or, key_and, id_or, id_and, ids_ascending) for the sort of a receive buffer, from the tag words every source
    sent with its counts (N_TAGS per source: id_or | fold bit, key_or, key_and, id_and, first ID, last ID, ascending).
    The buffer holds the sources' chunks in rank order, each a stable partition of the source's records: its IDs ascend
    iff every source's do, the sources' ID ranges follow each other in rank order, and no (unordered) halo copies came."""
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


class _CudaView:
    """Wraps a raw device pointer so torch can view it without a copy."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, n, dtype, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=device)


class _SymmBuffer:
    """A receive buffer every rank can store into over NVLink (torch symmetric memory)."""

    def __init__(self, nbytes, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.nbytes = nbytes
        self.t = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm_mem.rendezvous(self.t, group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    def barrier(self):
        self.hdl.barrier()


class _CountMatrix:
    """A g x row matrix of 64-bit counts in symmetric memory.  The kernel that finishes rank r's counts stores them as
    row r of EVERY rank's copy (its own, and the peers' through NVLink: self.row_ptrs); after one barrier every rank
    holds the whole matrix.  Replaces count -> host -> device -> NCCL all_gather -> host (two synchronisations and a
    collective launch per exchange: ~0.2 ms of the 1.5 ms frame at 2^20 objects per GPU), and the g - 1 copy-engine
    transfers per gather of the first symmetric-memory version (serialised in the stream: ~0.1 ms per frame at g = 8)."""

    def __init__(self, g, me, row, device, group):
        self.g, self.me, self.row = g, me, row
        self.buf = _SymmBuffer(g * row * 8, device, group)
        self.local = self.buf.t.view(torch.int64).view(g, row)
        self.row_ptrs = [self.buf.ptrs[p] + me * row * 8 for p in range(g)]

    def gather(self):
        self.buf.barrier()  # every rank's row has landed everywhere
        return self.local.cpu().numpy()


class CudaOps:
    """The shard-local operations on one B200, every one a call through the C ABI; the exchanges are
    partition passes that store directly into the peers' symmetric receive buffers."""

    def __init__(self, bp, kind, min_depth, device, group=None):
        if kind == 0:
            raise NotImplementedError("the distributed path handles the 64-bit index types")
        self.bp, self.kind, self.device = bp, kind, torch.device("cuda", device)
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        mk = lambda: bp.LayerBuilder().with_min_depth(min_depth).with_device(device).build(kind, "u32")
        self.enc, self.shard, self.static = mk(), mk(), mk()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for l in self.layers():
            l.set_stream(stream)
        self.rec_cap = self.pair_cap = 0
        self.rk = self.ri = self.rp = None
        self._cm_rec = self._cm_pair = None

    def layers(self):
        return (self.enc, self.shard, self.static)

    def encode(self, sys_bounds, bounds, ids, n):
        self.enc.clear()
        self.enc.extend_device(sys_bounds, bounds, ids, n)
        kp, ip, r, _ = self.enc.records_device()
        id_or = self.enc.masks()[2]
        return _view(kp, r, torch.int64, self.device), _view(ip, r, torch.int32, self.device), id_or

    def encode_count_matrix(self, sys_bounds, bounds, ids, n, splitters, allow_fold):
        """encode + count_records_matrix in one step for frames whose splitters are cached: the encode kernel counts the
        records per destination shard as it generates them (no counting pass over the keys), the row with its tag words is
        put together on the device, and the first host synchronisation of the frame is the one that fetches the finished
        matrix.  Returns (keys, ids, matrix)."""
        if self._cm_rec is None:
            self._cm_rec = _CountMatrix(self.world, self.rank, 2 * self.world + N_TAGS, self.device, self.group)
        self.enc.extend_count_rows(sys_bounds, bounds, ids, n, splitters, allow_fold, self._cm_rec.row_ptrs)
        mat = self._cm_rec.gather()
        kp, ip, r, _ = self.enc.records_device()
        return _view(kp, r, torch.int64, self.device), _view(ip, r, torch.int32, self.device), mat

    def count_records(self, keys, splitters):
        c, h = self.enc.count_records(keys, keys.shape[0], splitters)
        return [int(x) for x in c], [int(x) for x in h]

    def sort_tags(self):
        """What the receivers of this rank's records need to plan their sort without a pass over the records:
        [key_or, key_and, id_and, first ID, last ID, IDs ascending] of the freshly encoded tree (N_TAGS - 1 words)."""
        key_or, key_and, _, id_and = self.enc.masks()
        first, last, asc = self.enc.id_order()
        return [key_or, key_and, id_and, first, last, int(asc)]

    def count_records_matrix(self, keys, splitters, tags):
        """count_records + the exchange of the count matrix, without leaving the device until the matrix is complete:
        returns the [source, 2 g + N_TAGS] matrix (owned counts | halo counts | tags) as host numpy."""
        if self._cm_rec is None:
            self._cm_rec = _CountMatrix(self.world, self.rank, 2 * self.world + N_TAGS, self.device, self.group)
        assert len(tags) == N_TAGS
        self.enc.count_records_rows(keys, keys.shape[0], splitters, tags, self._cm_rec.row_ptrs)
        return self._cm_rec.gather()

    def count_pairs_matrix(self, raw, splitters, tag=0):
        """-> (count matrix [source, destination], the tag word every source sent along)."""
        if self._cm_pair is None:
            self._cm_pair = _CountMatrix(self.world, self.rank, self.world + 1, self.device, self.group)
        self.shard.count_pairs_rows(raw, raw.shape[0], splitters, [int(tag)], self._cm_pair.row_ptrs)
        mat = self._cm_pair.gather()
        return mat[:, :self.world], mat[:, self.world]

    def exchange_records(self, keys, ids, splitters, m_own, m_halo, fold=False):
        """fold: the cell flags of the encoded records leave in the top 3 bits of their IDs (dedup at the source across
        the exchange: the receiving shard emits every ID pair from its canonical shared cell only)."""
        me = self.rank
        recv = (m_own + m_halo).sum(axis=0)  # records every rank receives
        need = int(recv.max())
        if need > self.rec_cap or self.rk is None:  # the matrices are global, so every rank grows (collectively) at the same time
            self.rec_cap = int(need * 1.25) + 1024
            self.rk = _SymmBuffer(self.rec_cap * 8, self.device, self.group)
            self.ri = _SymmBuffer(self.rec_cap * 4, self.device, self.group)
            self._rk_ptrs = np.asarray(self.rk.ptrs, dtype=np.uint64)
            self._ri_ptrs = np.asarray(self.ri.ptrs, dtype=np.uint64)
        # (no halo arrays = no halo copies leave this rank, the usual case: no second pass over the keys)
        dk, di, hk, hi = scatter_destinations(self._rk_ptrs, self._ri_ptrs, m_own, m_halo, me)
        self.enc.scatter_records(keys, ids, keys.shape[0], splitters, dk, di, hk, hi, fold_cell_flags=fold)
        self.rk.barrier()  # every rank's stores have landed before anybody reads its receive buffer
        n_recv = int(recv[me])
        return self.rk.t.view(torch.int64)[:n_recv], self.ri.t.view(torch.int32)[:n_recv]

    def sort_records(self, keys, ids, flagged=False, plan=None):
        """plan = (key_or, key_and, id_or, id_and, ids_ascending) gathered with the count matrix: the sort then reads the
        receive buffer directly, with neither a staging copy nor a mask pass."""
        if plan is not None:
            self.shard.sort_from_device(keys, ids, keys.shape[0], flagged, *plan)
        else:
            self.shard.set_records(keys, ids, sorted_=False, on_device=True, n=keys.shape[0], flagged=flagged)
            self.shard.sort()
        if flagged:  # (records_device would strip the flags again; the scan below reads the shard layer itself)
            return None, None
        kp, ip, r, _ = self.shard.records_device()
        return _view(kp, r, torch.int64, self.device), _view(ip, r, torch.int32, self.device)

    def keep_static(self, keys, ids):
        """Sorts the received static records once and keeps them resident (Layer::merge's "static scene
        layer", reference README: sorted once, merged into every frame's dynamic layer)."""
        self.static.set_records(keys, ids, sorted_=False, on_device=True, n=keys.shape[0])
        self.static.sort()
        return len(self.static)

    def merge_static(self):
        """Layer::merge of the resident static shard into this frame's sorted dynamic shard; the scan's
        implicit sort is then one merge-path merge."""
        self.shard.merge(self.static)
        return len(self.shard)

    def scan_raw(self, keys, ids, n_halo, flt, dedup=True):
        """The sorted records live in self.shard; its first n_halo records are halo.  dedup = False: every ID pair is
        emitted from every cell the two objects share (bp_layer_set_scan_dedup)."""
        self.shard.set_halo(n_halo)
        self.shard.set_scan_dedup(dedup)
        ptr, n = self.shard.scan_raw_device(flt)
        self.shard.set_scan_dedup(True)
        self.shard.set_halo(0)
        self.saw_same_id = self.shard.stats()["rescans"] != 0   # an ID owns nested bounds here: some record is inactive
        return _view(ptr, n, torch.int64, self.device)

    def count_pairs(self, raw, splitters):
        return [int(x) for x in self.shard.count_pairs(raw, raw.shape[0], splitters)]

    def exchange_pairs(self, raw, splitters, m):
        me = self.rank
        recv = m.sum(axis=0)
        need = int(recv.max())
        if need > self.pair_cap or self.rp is None:
            self.pair_cap = int(need * 1.25) + 1024
            self.rp = _SymmBuffer(self.pair_cap * 8, self.device, self.group)
            self._rp_ptrs = np.asarray(self.rp.ptrs, dtype=np.uint64)
        dst = self._rp_ptrs + np.uint64(8) * m[:me].sum(axis=0).astype(np.uint64)
        self.shard.scatter_pairs(raw, raw.shape[0], splitters, dst)
        self.rp.barrier()
        return self.rp.t.view(torch.int64)[:int(recv[me])]

    def unique_pairs(self, raw, id_mask):
        ptr, n = self.shard.unique_pairs_inplace_device(raw, raw.shape[0], id_mask)  # raw = our receive buffer: sort scratch
        return _view(ptr, 2 * n, torch.int32, self.device).view(-1, 2)


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
            plan = sort_plan(tags, id_bits, n_halo)
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
