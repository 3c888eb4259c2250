"""Multi-GPU broadphase: Morton-prefix range sharding, one process per GPU -- Python binding.

The reference is single-process (Rayon on one host); this is the B200-native way to scale its hot path over NVLink.
The sharded frame itself lives behind the C ABI: a `bp_dist` context (include/bp.h, csrc/bp_dist.cu) runs

  1. encode     the rank's own objects -> unsorted (Index, ID) records                       [K1, local]
  2. splitters  g-1 key splitters from a sample of every rank's keys (sample sort), kept across frames while the shards
                stay balanced and recomputed when the fullest shard exceeds the mean by 15 %
  3. counts     records per destination shard (+ halo copies) -- taken by the encode kernel itself when the splitters
                are cached -- as this rank's row of a count matrix that the counting kernel stores into EVERY rank's copy
                over NVLink; one device barrier later everybody knows where its records go
  4. exchange   ONE partition pass (csrc/bp_exchange.cuh) writes every (tile, shard) run straight into the destination
                GPU's receive buffer through its peer mapping, the aligned body of a run as one cp.async.bulk copy: the
                pack kernel IS the all-to-all.  Records whose cell reaches past a shard's lower splitter are ancestors of
                records that shard owns and are sent to it as well (halo)
  5. sort       straight out of the receive buffer, planned from tag words that travelled with the counts    [K2]
  6. scan       over [halo | owned]; only pairs whose LATER record is owned are emitted                      [K3]
  7. dedup      raw pairs range-partitioned on the later ID, exchanged the same way, sorted + deduplicated  [K4]

in one call, `bp_dist_frame`; the slices it returns on ranks 0..g-1, concatenated, are exactly the reference's scan()
vector.  What is left here is what any host language would do with its own transport: all-gather the ranks' IPC blobs
once at start-up, and (for tests) the pair slices.  32-bit IDs, 64-bit index types.

An executable model of the same protocol in numpy (used by the gloo CPU tests with test doubles for the device
operations) lives in tests/dist_protocol.py.
"""
import numpy as np
import torch
import torch.distributed as dist


class _CudaView:
    """Wraps a raw device pointer so torch can view it without a copy."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _view(ptr, n, dtype, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=dtype, device=device)
    typestr = {torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_CudaView(ptr, n, typestr), device=device)


class DistContext:
    """The product path: one bp_dist context per rank (include/bp.h, csrc/bp_dist.cu) -- the whole sharded frame is ONE C-ABI
    call, bp_dist_frame.  The only thing this binding does with torch.distributed is what any host language would do with its
    own transport: all-gather the ranks' IPC blobs once at start-up (and the pair slices in gather_pairs, for tests)."""

    def __init__(self, bp, kind, min_depth, device, record_capacity, pair_capacity, group=None):
        import ctypes
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.bp, self.kind, self.min_depth = bp, kind, min_depth
        self.device = torch.device("cuda", device)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._stream = None
        self._options = {}
        self._h = None
        self._create(int(record_capacity), int(pair_capacity))

    def _create(self, record_capacity, pair_capacity):
        _lib, ct = self._lib, self._ct
        L = _lib.lib()
        self.record_capacity, self.pair_capacity = record_capacity, pair_capacity
        cfg = _lib.DistConfig(self.kind, self.min_depth, self.device.index, self.rank, self.world, record_capacity, pair_capacity)
        h = ct.c_void_p()
        st = L.bp_dist_create(ct.byref(cfg), ct.byref(h))
        if st != 0:
            raise _lib.BpError(st, "bp_dist_create")
        self._h = h
        if self.world > 1:
            nb = L.bp_dist_handle_bytes()
            blob = (ct.c_ubyte * nb)()
            self._ck(L.bp_dist_export(self._h, blob))
            mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(self.device)
            every = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=self.group)          # the caller's transport: any all-gather of 80 bytes would do
            allb = torch.cat(every).cpu().numpy().tobytes()
            self._ck(L.bp_dist_connect(self._h, allb))
        if self._stream is not None:
            self._ck(L.bp_dist_set_stream(self._h, ct.c_void_p(self._stream)))
        for k, v in self._options.items():
            self._ck(L.bp_dist_set_option(self._h, k, v))
        self._layers = [self.bp.Layer.borrow(L.bp_dist_layer(self._h, i), self.kind, "u32") for i in range(3)]

    def _ck(self, st):
        if st != 0:
            m = self._lib.lib().bp_dist_last_error(self._h)
            raise self._lib.BpError(st, m.decode() if m else "")

    def close(self):
        if self._h:
            self._lib.lib().bp_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def layers(self):
        """(encode, shard, static) layers of the context, for stats() / set_profiling()."""
        return self._layers

    def set_stream(self, cuda_stream):
        self._stream = cuda_stream
        self._ck(self._lib.lib().bp_dist_set_stream(self._h, self._ct.c_void_p(cuda_stream)))

    def set_option(self, option, value):
        self._options[option] = int(value)
        self._ck(self._lib.lib().bp_dist_set_option(self._h, option, int(value)))

    def set_static(self, sys_bounds, d_bounds, d_ids, n):
        sb = np.ascontiguousarray(sys_bounds, dtype=np.float32)
        self._ck(self._lib.lib().bp_dist_set_static(self._h, sb.ctypes.data, d_bounds.data_ptr(), d_ids.data_ptr(), n))
        self._has_static = True
        return self.last

    def frame(self, sys_bounds, d_bounds, d_ids, n, flt=None):
        """One frame on this rank's objects -> the rank's slice of the globally sorted, duplicate-free pair list as a
        (P_r, 2) int32 tensor (bit patterns of the u32 IDs; a view of the context's buffer, valid until the next call)."""
        ct = self._ct
        sb = np.ascontiguousarray(sys_bounds, dtype=np.float32)
        f = None if flt is None else flt._c()
        for attempt in range(3):   # (the record buffers and the pair buffers may each have to grow once)
            out, cnt = ct.c_void_p(), ct.c_size_t()
            st = self._lib.lib().bp_dist_frame(self._h, sb.ctypes.data, d_bounds.data_ptr() if n else None, d_ids.data_ptr() if n else None,
                                               n, None if f is None else ct.byref(f), ct.byref(out), ct.byref(cnt))
            if st == 4 and attempt < 2 and not self.last["have_static"]:
                # a receive buffer is too small -- on every rank alike (the count matrices are global): grow together, once
                need = self.last
                self.close()
                self._create(max(self.record_capacity, int(need["records_needed"] * 1.25) + 1024),
                             max(self.pair_capacity, int(need["pairs_needed"] * 1.25) + 1024))
                continue
            self._ck(st)
            break
        return _view(out.value, 2 * cnt.value, torch.int32, self.device).view(-1, 2)

    @property
    def last(self):
        info = self._lib.DistInfo()
        self._lib.lib().bp_dist_last_info(self._h, self._ct.byref(info))
        d = {k: int(getattr(info, k)) for k in ("records_local", "records_owned", "n_halo", "raw_pairs", "pairs", "records_needed",
                                                "pairs_needed", "fused", "rescanned", "rebalance_records", "rebalance_pairs")}
        d["halo"] = d["n_halo"]
        d["phases_ms"] = {name: float(info.phase_ms[i]) for i, name in enumerate(self._lib.DIST_PHASE_NAMES)}
        d["have_static"] = bool(getattr(self, "_has_static", False))
        return d

    def gather_pairs(self, pairs):
        """Concatenates every rank's slice in rank order: the reference's scan() vector (host numpy; tests and self-checks)."""
        if self.world == 1:
            return pairs.cpu().numpy().view(np.uint32)
        cnt = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=pairs.device)
        cnts = [torch.empty_like(cnt) for _ in range(self.world)]
        dist.all_gather(cnts, cnt, group=self.group)
        cnts = [int(c.item()) for c in cnts]
        mx = max(cnts) if cnts else 0
        buf = torch.zeros((mx, 2), dtype=pairs.dtype, device=pairs.device)
        buf[:pairs.shape[0]] = pairs
        allp = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(allp, buf, group=self.group)
        return np.concatenate([allp[r][:cnts[r]].cpu().numpy() for r in range(self.world)], axis=0).view(np.uint32)
