"""CPU test double of the device operations (numpy + the CPU oracle) and the gloo worker that drives the protocol
model tests/dist_protocol.DistLayer with it (the product runs the same protocol in C++: csrc/bp_dist.cu).  TEST INFRASTRUCTURE ONLY: lets world_size > 1 tests check the splitter /
halo / ownership / global-dedup choreography of the multi-GPU path on a CPU-only box.  The product
path (bp_dist_frame) never touches any of this."""
import os
import sys
import traceback

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cpu_oracle as co  # noqa: E402
from oracle import pyref  # noqa: E402


def _i64(u):
    return torch.from_numpy(np.ascontiguousarray(u, dtype=np.uint64).view(np.int64).copy())


def _i32(u):
    return torch.from_numpy(np.ascontiguousarray(u, dtype=np.uint32).view(np.int32).copy())


def _u64(t):
    return t.numpy().view(np.uint64)


def _u32(t):
    return t.numpy().view(np.uint32)


class CpuOps:
    device = torch.device("cpu")

    def __init__(self, kind, min_depth):
        self.kind, self.min_depth = kind, min_depth

    def encode(self, sys_bounds, bounds, ids, n):
        L = co.OracleLayer(self.kind, 4, self.min_depth)
        L.extend(sys_bounds, bounds[:n], ids[:n])
        k, i = L.records()
        id_or = int(np.bitwise_or.reduce(i)) if i.size else 0
        return _i64(k), _i32(i.astype(np.uint32)), id_or

    # -- helpers ---------------------------------------------------------------------------------
    def _hi(self, k):
        """run_upper_key for an array of keys (numpy)."""
        _, dim, depth_bits, axis_bits = pyref.KINDS[self.kind]
        used = (1 << (dim * axis_bits + depth_bits)) - 1
        low = np.array([(~pyref.level_mask(self.kind, d)) & used for d in range(axis_bits + 1)], dtype=np.uint64)
        depth = (k & np.uint64((1 << depth_bits) - 1)).astype(np.int64)
        return k | low[depth]

    def _homes(self, k, splitters):
        spl = np.asarray(splitters, dtype=np.uint64)
        return np.searchsorted(spl, k, side="right"), np.searchsorted(spl, self._hi(k), side="right")

    def count_records(self, keys, splitters):
        k = _u64(keys)
        g = len(splitters) + 1
        home, last = self._homes(k, splitters)
        counts = np.bincount(home, minlength=g)
        halo = np.zeros(g, dtype=np.int64)
        for s in range(1, g):
            halo[s] = int(((home < s) & (last >= s)).sum())
        return [int(c) for c in counts], [int(h) for h in halo]

    def _a2a(self, send, send_counts, recv_counts):
        recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype)
        dist.all_to_all_single(recv, send, [int(c) for c in recv_counts], [int(c) for c in send_counts])
        return recv

    def exchange_records(self, keys, ids, splitters, m_own, m_halo):
        """Test-double transport: per-destination chunks [owned | halo copies] through a gloo all-to-all
        (the product scatters the same chunks into the peers' symmetric memory instead)."""
        me = dist.get_rank()
        k, i = _u64(keys), _u32(ids)
        g = len(splitters) + 1
        home, last = self._homes(k, splitters)
        ck, ci = [], []
        for d in range(g):
            own = home == d
            hal = (home < d) & (last >= d)
            assert int(own.sum()) == int(m_own[me, d]) and int(hal.sum()) == int(m_halo[me, d])
            ck += [k[own], k[hal]]
            ci += [i[own], i[hal]]
        both = m_own + m_halo
        rk = self._a2a(_i64(np.concatenate(ck)), both[me, :], both[:, me])
        ri = self._a2a(_i32(np.concatenate(ci)), both[me, :], both[:, me])
        return rk, ri

    def sort_records(self, keys, ids):
        k, i = pyref.sort_records(_u64(keys), _u32(ids))
        return _i64(k), _i32(i)

    def keep_static(self, keys, ids):
        self._static = pyref.sort_records(_u64(keys), _u32(ids))
        return self._static[0].shape[0]

    def merge_static(self):
        self._merge = True
        return 0

    def scan_raw(self, keys, ids, n_halo, flt):
        fk, arg = flt if flt else (0, 0)
        if getattr(self, "_merge", False):  # the double simply re-sorts the concatenation, like the reference does
            k = np.concatenate([_u64(keys), self._static[0]])
            i = np.concatenate([_u32(ids), self._static[1]])
            k, i = pyref.sort_records(k, i)
            keys, ids = _i64(k), _i32(i)
            self._merge = False
        a, b = pyref.scan_raw(self.kind, _u64(keys), _u32(ids), fk, arg, None, first_owned=n_halo)
        return _i64((a << np.uint64(32)) | b)

    def count_pairs(self, raw, splitters):
        r = _u64(raw)
        bucket = np.searchsorted(np.asarray(splitters, dtype=np.uint64), r >> np.uint64(32), side="right")
        return [int(c) for c in np.bincount(bucket, minlength=len(splitters) + 1)]

    def exchange_pairs(self, raw, splitters, m):
        me = dist.get_rank()
        r = _u64(raw)
        bucket = np.searchsorted(np.asarray(splitters, dtype=np.uint64), r >> np.uint64(32), side="right")
        order = np.argsort(bucket, kind="stable")
        return self._a2a(_i64(r[order]), m[me, :], m[:, me])

    def unique_pairs(self, raw, id_mask):
        r = np.unique(_u64(raw))
        assert ((r >> np.uint64(32)) <= np.uint64(id_mask)).all() and ((r & np.uint64(0xFFFFFFFF)) <= np.uint64(id_mask)).all()
        out = np.stack([(r >> np.uint64(32)).astype(np.uint32), (r & np.uint64(0xFFFFFFFF)).astype(np.uint32)], axis=1)
        return torch.from_numpy(out.view(np.int32).copy())


def _s64(v):
    v = int(v) & 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >> 63 else v


class _EncInfo:
    def __init__(self, min_depth):
        self.min_depth = min_depth


class CpuOpsProduct(CpuOps):
    """The same double behind the PRODUCT-side protocol of DistLayer.frame (what bp_dist_frame speaks): the count matrix
    carries every sender's tag words, the receivers plan their sort from them (dist.sort_plan), frames with cached splitters
    take the counts together with the encode.  sort_records CHECKS the plan it is handed against the records that actually
    arrived -- the masks must cover them, and "IDs ascending" must be true of the receive buffer, because the GPU sort then
    skips the ID digits and relies on stability.  (The cell flags of the fold are a device-side matter: ignored here.)"""

    def __init__(self, kind, min_depth):
        super().__init__(kind, min_depth)
        self.enc = _EncInfo(min_depth)
        self.fused_frames = 0
        self.plans = []  # (ids_ascending, n_records) of every planned sort

    def encode(self, sys_bounds, bounds, ids, n):
        k, i, id_or = super().encode(sys_bounds, bounds, ids, n)
        ku, iu = _u64(k), _u32(i)
        full = (1 << 64) - 1
        obj = np.asarray(ids[:n]).astype(np.int64)
        empty = ku.size == 0
        self._tags = [int(np.bitwise_or.reduce(ku)) if not empty else 0, int(np.bitwise_and.reduce(ku)) if not empty else full,
                      int(np.bitwise_and.reduce(iu)) if not empty else full,
                      full if empty else int(obj[0]), 0 if empty else int(obj[-1]),
                      1 if empty else int(bool((np.diff(obj) >= 0).all()))]
        return k, i, id_or

    def sort_tags(self):
        return list(self._tags)

    def set_pair_later_fixed(self, fixed, value):
        self._later_fixed = (fixed & 0xFFFFFFFF, value & 0xFFFFFFFF)

    def unique_pairs(self, raw, id_mask):
        fixed, value = getattr(self, "_later_fixed", (0, 0))  # CHECKED against the pairs that arrived, like the sort plan
        self._later_fixed = (0, 0)
        later = _u64(raw) >> np.uint64(32)
        assert ((later & np.uint64(fixed)) == np.uint64(value)).all(), "a later ID outside the bits its shard is said to share"
        self.later_fixed_bits = getattr(self, "later_fixed_bits", []) + [bin(fixed).count("1")]
        return super().unique_pairs(raw, id_mask)

    def count_records_matrix(self, keys, splitters, tags):
        counts, halo = self.count_records(keys, splitters)
        row = torch.tensor([_s64(v) for v in counts + halo + list(tags)], dtype=torch.int64)
        out = [torch.empty_like(row) for _ in range(dist.get_world_size())]
        dist.all_gather(out, row)
        return torch.stack(out).numpy()

    def encode_count_matrix(self, sys_bounds, bounds, ids, n, splitters, allow_fold):
        k, i, id_or = self.encode(sys_bounds, bounds, ids, n)
        tag0 = id_or | ((1 << 63) if (allow_fold and id_or < (1 << 29)) else 0)
        self.fused_frames += 1
        return k, i, self.count_records_matrix(k, splitters, [tag0] + self.sort_tags())

    def exchange_records(self, keys, ids, splitters, m_own, m_halo, fold=False):
        return super().exchange_records(keys, ids, splitters, m_own, m_halo)

    def sort_records(self, keys, ids, flagged=False, plan=None):
        assert plan is not None, "the product path always plans the shard sort from the tag words"
        k, i = _u64(keys), _u32(ids)
        key_or, key_and, id_or, id_and, asc = plan
        if k.size:
            assert int(np.bitwise_or.reduce(k)) & ~key_or == 0 and key_and & ~int(np.bitwise_and.reduce(k)) == 0
            assert int(np.bitwise_or.reduce(i)) & ~id_or == 0 and id_and & ~int(np.bitwise_and.reduce(i)) == 0
        want_k, want_i = pyref.sort_records(k, i)
        if asc:  # what the GPU does with this plan: a stable sort on the key alone
            assert (np.diff(i.astype(np.int64)) >= 0).all(), "plan says the IDs ascend in the receive buffer, they do not"
            order = np.argsort(k, kind="stable")
            assert (k[order] == want_k).all() and (i[order] == want_i).all()
        self.plans.append((bool(asc), int(k.size)))
        return _i64(want_k), _i32(want_i)


def make_case(name):
    """Deterministic scenes that stress the distributed logic.  Returns (kind, min_depth, sys, bounds, ids, filter)."""
    rng = np.random.Generator(np.random.Philox(abs(hash(name)) % (1 << 31) if False else sum(map(ord, name))))
    if name == "uniform3d":
        n, kind, dim = 6000, 2, 3
        size = (0.03 * rng.random((n, dim))).astype(np.float32)
    elif name == "big_objects3d":  # scene-sized boxes: halos at depth 0..2 reach every shard
        n, kind, dim = 4000, 2, 3
        size = (0.02 * rng.random((n, dim))).astype(np.float32)
        size[::50] = (0.3 + 0.6 * rng.random((len(size[::50]), dim))).astype(np.float32)
    elif name == "multibounds2d":  # several (nested) bounds per ID: inactive records across shard borders
        n, kind, dim = 5000, 1, 2
        size = (0.2 * rng.random((n, dim)) ** 3).astype(np.float32)
    elif name == "multibounds3d":  # Index64_3D, several nested bounds per ID, multi-cell objects: with the cell flags riding
        n, kind, dim = 6000, 2, 3  # across the exchange (dedup at the source) a skipped inactive record in one shard must
        size = (0.01 + 0.02 * rng.random((n, dim))).astype(np.float32)  # switch the dedup off in every other shard too
    elif name == "skewed3d":  # almost everything in one corner: very uneven key distribution
        n, kind, dim = 5000, 2, 3
        size = (0.01 * rng.random((n, dim))).astype(np.float32)
    else:
        raise ValueError(name)
    sysb = np.concatenate([np.zeros(dim), np.ones(dim)]).astype(np.float32)
    mn = (rng.random((n, dim)) * (1.0 - size)).astype(np.float32)
    if name == "skewed3d":
        mn = (mn ** 4).astype(np.float32)
    bounds = np.concatenate([mn, np.minimum(mn + size, 1.0)], axis=1).astype(np.float32)
    if name == "multibounds2d":
        ids = np.sort(rng.integers(0, n // 4, size=n)).astype(np.uint32)
    elif name == "multibounds3d":   # unsorted: an ID's bounds sit on different ranks
        ids = rng.integers(0, n // 3, size=n).astype(np.uint32)
        # a few IDs get a tiny box nested inside one of their boxes (an inactive record); boxes this small rarely reach
        # past a splitter, so most shards see neither a halo nor a same-ID item and keep the dedup at the source on
        for j in range(0, n, 97):
            src = (j * 31 + 7) % n
            c = (bounds[src, :dim] + bounds[src, dim:]) / 2
            bounds[j, :dim], bounds[j, dim:] = c, c + np.float32(1e-4)
            ids[j] = ids[src]
    else:
        ids = np.arange(n, dtype=np.uint32)
    flt = (1, 0) if name == "uniform3d" else None  # ID-parity filter on one case
    return kind, 0, sysb, bounds, ids, flt


def reference_pairs(case):
    kind, md, sysb, bounds, ids, flt = make_case(case)
    L = co.OracleLayer(kind, 4, md)
    L.extend(sysb, bounds, ids)
    fk, arg = flt if flt else (0, 0)
    return L.scan(fk, arg).astype(np.uint32)


def reference_pairs_unfiltered(case):
    kind, md, sysb, bounds, ids, _ = make_case(case)
    L = co.OracleLayer(kind, 4, md)
    L.extend(sysb, bounds, ids)
    return L.scan().astype(np.uint32)


def reference_static_dynamic(frame):
    kind, md, sysb, sb, sids, _ = make_case("big_objects3d")
    _, _, _, db, dids, _ = make_case("uniform3d")
    dids = (dids + np.uint32(100_000)).astype(np.uint32)
    moved = np.clip(db + np.float32(0.001 * frame), 0.0, 1.0).astype(np.float32)
    S = co.OracleLayer(kind, 4, md)
    S.extend(sysb, sb, sids)
    S.sort()
    D = co.OracleLayer(kind, 4, md)
    D.extend(sysb, moved, dids)
    D.sort()
    D.merge(S)
    return D.scan().astype(np.uint32)


def worker(rank, world, port, cases, empty_rank, out_dir, product=False):
    Ops = CpuOpsProduct if product else CpuOps
    stats = []
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import _loadpkg
        bp = _loadpkg.load()
        from tests import dist_protocol as bpd
        for case in cases:
            kind, md, sysb, bounds, ids, flt = make_case(case)
            n = bounds.shape[0]
            # block distribution by object order; optionally one rank gets nothing
            holders = [r for r in range(world) if r != empty_rank]
            cuts = np.linspace(0, n, len(holders) + 1).astype(int)
            if rank in holders:
                h = holders.index(rank)
                lo, hi = cuts[h], cuts[h + 1]
            else:
                lo, hi = 0, 0
            dl = bpd.DistLayer(Ops(kind, md), kind)
            stats.append(dl.ops)
            dl.frame(sysb, bounds[lo:hi], ids[lo:hi], hi - lo, flt)  # first frame computes the splitters
            pairs = dl.frame(sysb, bounds[lo:hi], ids[lo:hi], hi - lo, flt)  # second frame reuses them
            allp = dl.gather_pairs(pairs)
            if rank == 0:
                np.save(os.path.join(out_dir, "%s.npy" % case), allp)
                np.save(os.path.join(out_dir, "%s.halo.npy" % case), np.array([dl.last["halo"]]))
            halos = [None] * world
            dist.all_gather_object(halos, dl.last["halo"])
            if rank == 0:
                np.save(os.path.join(out_dir, "%s.halos.npy" % case), np.array(halos))
        # config-4 shape: a static scene sharded once, a fresh dynamic layer merged in every frame
        kind, md, sysb, sb, sids, _ = make_case("big_objects3d")
        _, _, _, db, dids, _ = make_case("uniform3d")
        dids = (dids + np.uint32(100_000)).astype(np.uint32)
        cs = np.linspace(0, sb.shape[0], world + 1).astype(int)
        cd = np.linspace(0, db.shape[0], world + 1).astype(int)
        dl = bpd.DistLayer(Ops(kind, md), kind)
        stats.append(dl.ops)
        dl.set_static(sysb, sb[cs[rank]:cs[rank + 1]], sids[cs[rank]:cs[rank + 1]], cs[rank + 1] - cs[rank])
        for frame in range(2):
            shift = np.float32(0.001 * frame)
            moved = np.clip(db + shift, 0.0, 1.0).astype(np.float32)
            pairs = dl.frame(sysb, moved[cd[rank]:cd[rank + 1]], dids[cd[rank]:cd[rank + 1]], cd[rank + 1] - cd[rank], None)
            allp = dl.gather_pairs(pairs)
            if rank == 0:
                np.save(os.path.join(out_dir, "static_dynamic_%d.npy" % frame), allp)
        if product:
            # the scene changes under cached splitters: the frame with the stale splitters is still exact, notices the
            # imbalance and drops them; the next frame samples new ones (plain counts), the one after is fused again
            kind, md, sysb, ba, ia, _ = make_case("uniform3d")
            _, _, _, bb, ib, _ = make_case("skewed3d")
            ca = np.linspace(0, ba.shape[0], world + 1).astype(int)
            cb = np.linspace(0, bb.shape[0], world + 1).astype(int)
            dl = bpd.DistLayer(Ops(kind, md), kind)
            for f, (b, i, c) in enumerate([(ba, ia, ca), (ba, ia, ca), (bb, ib, cb), (bb, ib, cb), (bb, ib, cb)]):
                pairs = dl.frame(sysb, b[c[rank]:c[rank + 1]], i[c[rank]:c[rank + 1]], c[rank + 1] - c[rank], None)
                allp = dl.gather_pairs(pairs)
                if rank == 0:
                    np.save(os.path.join(out_dir, "rebalance_%d.npy" % f), allp)
            if rank == 0:
                np.save(os.path.join(out_dir, "rebalance_fused.npy"), np.array([dl.ops.fused_frames]))
        if product:  # every DistLayer ran frames with cached splitters (fused counts) and planned sorts of both kinds
            mine = [(o.fused_frames, o.plans) for o in stats]
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            if rank == 0:
                import json
                with open(os.path.join(out_dir, "product_stats.json"), "w") as f:
                    json.dump(everyone, f)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        traceback.print_exc()
        raise
