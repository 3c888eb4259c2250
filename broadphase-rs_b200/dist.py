"""Multi-GPU broadphase: Morton-prefix range sharding, one process per GPU.

The reference is single-process (Rayon on one host); this is the B200-native way to scale its hot
path over NVLink.  One frame, on every rank r of g:

  1. encode     the rank's own objects -> unsorted (Index, ID) records          [K1, local]
  2. splitters  a regular sample of the local keys is all-gathered; every rank derives the same
                g-1 key splitters (sample sort)                                  [all_gather, tiny]
  3. exchange   records are range-partitioned by splitter (the onesweep pass with a splitter-search
                digit; stable) and exchanged with ONE all-to-all over NVLink     [all_to_all]
  4. sort       the received records (IDs still ascend in record order when every rank's IDs ascend
                and ranks hold ascending ID blocks, so the key-only sort applies) [K2, local]
  5. halo       by the contiguity lemma the only records of earlier shards that can be
                ancestors of anything in shard s are those whose cell contains the cell of s's FIRST
                record: for every depth d <= depth(first) the equal-key run of
                (key_first & level_mask(d)) | d.  Earlier shards look these runs up and send them
                (usually nothing; a few records for scene-sized objects)          [all_gather + all_to_all, tiny]
  6. scan       over [halo | owned]; only pairs whose LATER record is owned are emitted, so every raw
                pair is produced exactly once globally                            [K3, local]
  7. dedup      the same ID pair can be produced in several shards, and the reference returns one
                globally sorted vector: raw pairs are range-partitioned on the later ID and
                exchanged, then sorted + deduplicated per rank.  Concatenating the ranks' results in
                rank order is exactly the reference's scan() output.              [all_to_all, K4]

The choreography below is independent of where the local operations run: `ops` is CudaOps (the
product: every operation is a C-ABI call into libbroadphase_b200.so on device tensors, collectives
over NCCL) or, in the CPU tests only, a numpy test double with gloo -- which lets world_size-2 tests
check the splitter / halo / ownership logic without GPUs.  32-bit IDs only.
"""
import numpy as np
import torch
import torch.distributed as dist

# (key bits, DIM, DEPTH_BITS, AXIS_BITS) -- reference src/index.rs:293-295
KIND_PARAMS = {0: (32, 2, 4, 14), 1: (64, 2, 5, 29), 2: (64, 3, 5, 19)}
SAMPLES_PER_RANK = 2048
U64_MAX = np.uint64(0xFFFFFFFFFFFFFFFF)


def level_mask(kind, depth):
    _, dim, depth_bits, axis_bits = KIND_PARAMS[kind]
    if depth <= 0:
        return 0
    return ((1 << (dim * depth)) - 1) << (dim * axis_bits + depth_bits - dim * depth)


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
        return np.zeros(0, dtype=np.uint64)
    q = [s[min(s.shape[0] - 1, (i * s.shape[0]) // parts)] for i in range(1, parts)]
    return np.asarray(q, dtype=np.uint64)


class _CudaView:
    """Wraps a raw device pointer so torch can view it without a copy."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, n, dtype, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=device)


class CudaOps:
    """The shard-local operations on one B200, every one a call through the C ABI."""

    def __init__(self, bp, kind, min_depth, device):
        if kind == 0:
            raise NotImplementedError("the distributed path handles the 64-bit index types")
        self.bp, self.kind, self.device = bp, kind, torch.device("cuda", device)
        mk = lambda: bp.LayerBuilder().with_min_depth(min_depth).with_device(device).build(kind, "u32")
        self.enc, self.shard, self.scanl = mk(), mk(), mk()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for l in (self.enc, self.shard, self.scanl):
            l.set_stream(stream)

    def layers(self):
        return (self.enc, self.shard, self.scanl)

    def encode(self, sys_bounds, bounds, ids, n):
        self.enc.clear()
        self.enc.extend_device(sys_bounds, bounds, ids, n)
        kp, ip, r, _ = self.enc.records_device()
        id_or = self.enc.masks()[2]
        return _view(kp, r, torch.int64, self.device), _view(ip, r, torch.int32, self.device), id_or

    def partition_records(self, keys, ids, splitters):
        n = keys.shape[0]
        ok, oi = torch.empty_like(keys), torch.empty_like(ids)
        counts = self.enc.partition_records(keys, ids, n, splitters, ok, oi)
        return ok, oi, [int(c) for c in counts]

    def sort_records(self, keys, ids):
        self.shard.set_records(keys, ids, sorted_=False, on_device=True, n=keys.shape[0])
        self.shard.sort()
        kp, ip, r, _ = self.shard.records_device()
        return _view(kp, r, torch.int64, self.device), _view(ip, r, torch.int32, self.device)

    def lookup_ranges(self, sorted_keys, queries):
        return self.shard.lookup_ranges(sorted_keys, sorted_keys.shape[0], queries)

    def scan_raw(self, keys, ids, n_halo, flt):
        """keys/ids: the shard's sorted records with n_halo halo records in front."""
        if n_halo == 0:
            layer = self.shard  # the sorted records already live in this layer
        else:
            layer = self.scanl
            layer.set_records(keys, ids, sorted_=True, on_device=True, n=keys.shape[0])
        layer.set_halo(n_halo)
        ptr, n = layer.scan_raw_device(flt)
        layer.set_halo(0)
        return _view(ptr, n, torch.int64, self.device)

    def partition_pairs(self, raw, splitters):
        out = torch.empty_like(raw)
        counts = self.scanl.partition_pairs(raw, raw.shape[0], splitters, out)
        return out, [int(c) for c in counts]

    def unique_pairs(self, raw, id_mask):
        ptr, n = self.scanl.unique_pairs_device(raw, raw.shape[0], id_mask)
        return _view(ptr, 2 * n, torch.int32, self.device).view(-1, 2)


class DistLayer:
    """The distributed counterpart of clear -> extend -> par_sort -> par_scan(_filtered) for one frame."""

    def __init__(self, ops, kind, group=None):
        self.ops, self.kind, self.group = ops, kind, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.last = {}

    # -- collectives (device tensors over NCCL in production, CPU tensors over gloo in the tests) --
    def _all_gather(self, t):
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.stack(out)

    def _all_to_all(self, send, send_counts, recv_counts):
        recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send, [int(c) for c in recv_counts], [int(c) for c in send_counts], group=self.group)
        return recv

    def _count_matrix(self, my_counts, device):
        row = torch.tensor([int(c) for c in my_counts], dtype=torch.int64, device=device)
        return self._all_gather(row).cpu().numpy()  # [source, destination]

    def frame(self, sys_bounds, bounds, ids, n, flt=None):
        """Runs one frame on this rank's objects.  Returns the rank's slice of the globally sorted,
        deduplicated pair list as an (P_r, 2) int32 tensor (bit patterns of the u32 IDs)."""
        ops, g, me = self.ops, self.world, self.rank
        dev = ops.device

        # 1. encode
        keys, rids, id_or = ops.encode(sys_bounds, bounds, ids, n)
        r_loc = keys.shape[0]

        # 2. splitters from an all-gathered sample (+ the ID bits and a sample of IDs, piggybacked)
        m = SAMPLES_PER_RANK
        meta = torch.full((2 * m + 2,), -1, dtype=torch.int64, device=dev)
        if r_loc:
            step = max(1, r_loc // m)
            ks = keys[::step][:m]
            meta[:ks.shape[0]] = ks
            meta[m:m + ks.shape[0]] = rids[::step][:m].to(torch.int64) & 0xFFFFFFFF
        meta[2 * m] = id_or
        meta[2 * m + 1] = r_loc
        gathered = self._all_gather(meta).cpu().numpy()
        key_sample = gathered[:, :m].reshape(-1)
        key_sample = key_sample[key_sample >= 0].view(np.uint64) if key_sample.size else key_sample.view(np.uint64)
        id_sample = gathered[:, m:2 * m].reshape(-1)
        id_sample = id_sample[id_sample >= 0].astype(np.uint64)
        id_bits = 0
        for v in gathered[:, 2 * m]:
            id_bits |= int(v)
        id_mask = (1 << max(1, id_bits.bit_length())) - 1
        splitters = choose_splitters(key_sample, g)
        if splitters.shape[0] < g - 1:  # no records anywhere
            splitters = np.full(g - 1, U64_MAX, dtype=np.uint64)

        # 3. partition + one all-to-all of the records
        pk, pi, send_counts = ops.partition_records(keys, rids, splitters)
        cm = self._count_matrix(send_counts, dev)
        recv_counts = cm[:, me]
        rk = self._all_to_all(pk, send_counts, recv_counts)
        ri = self._all_to_all(pi, send_counts, recv_counts)

        # 4. local sort
        sk, si = ops.sort_records(rk, ri)
        r_own = sk.shape[0]

        # 5. halos: earlier shards send the records whose cell contains the cell of my first record
        first = torch.full((2,), -1, dtype=torch.int64, device=dev)
        if r_own:
            first[0] = sk[0]
            first[1] = 1
        firsts = self._all_gather(first).cpu().numpy()
        queries, owner = [], []
        for s in range(me + 1, g):
            if firsts[s, 1] == 1:
                for q in ancestor_keys(self.kind, int(np.uint64(firsts[s, 0]))):
                    queries.append(q)
                    owner.append(s)
        halo_send = [0] * g
        ranges = []
        if queries and r_own:
            lo, hi = ops.lookup_ranges(sk, np.asarray(queries, dtype=np.uint64))
            for s, a, b in zip(owner, lo, hi):
                if b > a:
                    halo_send[s] += int(b - a)
                    ranges.append((s, int(a), int(b)))
        hm = self._count_matrix(halo_send, dev)
        n_halo = int(hm[:, me].sum())
        if hm.sum() > 0:  # the matrix is identical everywhere, so every rank takes the same branch
            if ranges:  # ranges are grouped by destination and ascend within it
                idx = torch.cat([torch.arange(a, b, device=dev) for _, a, b in ranges])
                hk_send, hi_send = sk[idx], si[idx]
            else:
                hk_send, hi_send = sk[:0], si[:0]
            hk = self._all_to_all(hk_send.contiguous(), halo_send, hm[:, me])
            hi_ = self._all_to_all(hi_send.contiguous(), halo_send, hm[:, me])
            if n_halo:
                sk = torch.cat([hk, sk])
                si = torch.cat([hi_, si])

        # 6. shard-local scan; pairs whose later record is a halo record belong to an earlier shard
        raw = ops.scan_raw(sk, si, n_halo, flt)

        # 7. global dedup: range-partition the raw pairs on the later ID, exchange, sort + unique
        a_splitters = choose_splitters(id_sample, g)
        if a_splitters.shape[0] < g - 1:
            a_splitters = np.full(g - 1, U64_MAX, dtype=np.uint64)
        pp, psend = ops.partition_pairs(raw, a_splitters)
        pm = self._count_matrix(psend, dev)
        rp = self._all_to_all(pp, psend, pm[:, me])
        pairs = ops.unique_pairs(rp, id_mask)
        self.last = dict(records_local=r_loc, records_owned=r_own, halo=n_halo, raw_pairs=int(raw.shape[0]),
                         pairs=int(pairs.shape[0]), record_matrix=cm, pair_matrix=pm)
        return pairs

    def gather_pairs(self, pairs):
        """Concatenates every rank's slice in rank order: the reference's scan() vector (host numpy)."""
        cnt = self._all_gather(torch.tensor([pairs.shape[0]], dtype=torch.int64, device=pairs.device)).cpu().numpy().reshape(-1)
        mx = int(cnt.max()) if cnt.size else 0
        buf = torch.zeros((mx, 2), dtype=pairs.dtype, device=pairs.device)
        buf[:pairs.shape[0]] = pairs
        allp = self._all_gather(buf).cpu().numpy()
        return np.concatenate([allp[r, :int(cnt[r])] for r in range(self.world)], axis=0).view(np.uint32)
