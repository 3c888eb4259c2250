"""Python mirror of the reference's `Layer<Index, ID>` / `LayerBuilder` API over the C ABI.

Same method names, argument meaning and implicit behaviour as the crate (src/layer.rs):
    LayerBuilder().with_min_depth(4).build(Index32_2D, "u32")   # src/layer.rs:620-696
    layer.clear(); layer.extend(system_bounds, bounds, ids)     # :84-88, :94-121
    layer.par_sort(); layer.par_scan(); layer.scan_filtered(f)  # :146-165, :449-520
    layer.merge(other); layer.iter()                            # :127-138, :79-81
Host arrays are numpy; device arrays are anything exposing `data_ptr()` (torch tensors) or raw
integer device pointers.  Compute never happens on the host: every method calls the CUDA library.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import BpError, Filter, LayerConfig, Stats, check, lib

Index32_2D, Index64_2D, Index64_3D = 0, 1, 2
INDEX_DIM = {Index32_2D: 2, Index64_2D: 2, Index64_3D: 3}
INDEX_KEY_DTYPE = {Index32_2D: np.uint32, Index64_2D: np.uint64, Index64_3D: np.uint64}

FILTER_NONE, FILTER_ID_PARITY, FILTER_XOR_MASK, FILTER_CATEGORY, FILTER_SPHERES = 0, 1, 2, 3, 4
PICK_SPHERE, PICK_AABB = 0, 1


class ScanFilter:
    """A device functor standing in for scan_filtered's closure (include/bp.h, bp_filter_kind)."""

    def __init__(self, kind=FILTER_NONE, arg=0, table=None):
        self.kind, self.arg = kind, arg
        if table is None:
            self.table = None
        elif kind == FILTER_SPHERES:   # rows of {x, y, z, r}
            self.table = np.ascontiguousarray(table, dtype=np.float32).reshape(-1, 4)
        else:                          # rows of {cat, msk}
            self.table = np.ascontiguousarray(table, dtype=np.uint32).reshape(-1, 2)

    @staticmethod
    def none():
        return ScanFilter(FILTER_NONE)

    @staticmethod
    def id_parity():
        return ScanFilter(FILTER_ID_PARITY)

    @staticmethod
    def xor_mask(mask):
        return ScanFilter(FILTER_XOR_MASK, mask)

    @staticmethod
    def category(table):
        return ScanFilter(FILTER_CATEGORY, 0, table)

    @staticmethod
    def spheres(table):
        """Fused narrow phase: only pairs whose spheres (rows {x, y, z, r} indexed by ID) touch pass."""
        return ScanFilter(FILTER_SPHERES, 0, table)

    def _c(self):
        f = Filter()
        f.kind, f.arg, f.table_on_device = self.kind, self.arg, 0
        if self.table is not None:
            f.table = self.table.ctypes.data
            f.n_table = self.table.shape[0]
        return f


def _dev_ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    return x.data_ptr()


class Layer:
    """A group of collision data resident on one B200 (reference: src/layer.rs:42-68)."""

    def __init__(self, index=Index64_3D, id_type="u32", min_depth=0, index_capacity=0, collision_capacity=0,
                 test_capacity=0, device=-1):
        self.index = index
        self.dim = INDEX_DIM[index]
        self.id_bytes = {"u32": 4, "u64": 8, 4: 4, 8: 8}[id_type]
        self.id_dtype = np.uint32 if self.id_bytes == 4 else np.uint64
        self.key_dtype = INDEX_KEY_DTYPE[index]
        cfg = LayerConfig(index, self.id_bytes, min_depth, device, index_capacity, collision_capacity, test_capacity)
        h = ctypes.c_void_p()
        check(lib().bp_layer_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self._device = device

    @classmethod
    def borrow(cls, handle, index, id_type="u32"):
        """A view of a layer somebody else owns (the layers inside a bp_dist context): never destroyed from here."""
        self = cls.__new__(cls)
        self.index = index
        self.dim = INDEX_DIM[index]
        self.id_bytes = {"u32": 4, "u64": 8, 4: 4, 8: 8}[id_type]
        self.id_dtype = np.uint32 if self.id_bytes == 4 else np.uint64
        self.key_dtype = INDEX_KEY_DTYPE[index]
        self._h = ctypes.c_void_p(handle)
        self._borrowed = True
        return self

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                lib().bp_layer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, status):
        check(status, self._h)

    # ---- Layer API ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._ck(lib().bp_layer_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def clear(self):
        self._ck(lib().bp_layer_clear(self._h))

    def extend(self, system_bounds, bounds, ids):
        """Layer::extend from host (numpy) arrays: bounds (n, 2*D) f32 as (min.., max..), ids (n,)."""
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32).reshape(2 * self.dim)
        b = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 2 * self.dim)
        i = np.ascontiguousarray(ids, dtype=self.id_dtype).reshape(-1)
        if i.shape[0] != b.shape[0]:
            raise ValueError("bounds and ids differ in length")
        self._ck(lib().bp_layer_extend_host(self._h, sysb.ctypes.data, b.ctypes.data, i.ctypes.data, b.shape[0]))

    def extend_device(self, system_bounds, d_bounds, d_ids, n):
        """Layer::extend from device-resident arrays (torch tensors or raw pointers)."""
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32).reshape(2 * self.dim)
        self._ck(lib().bp_layer_extend_device(self._h, sysb.ctypes.data, _dev_ptr(d_bounds), _dev_ptr(d_ids), n))

    def extend_count_rows(self, system_bounds, d_bounds, d_ids, n, splitters, allow_fold, out_rows):
        """clear + extend_device with the per-shard counts taken by the encode kernel, and the count row (counts, halo
        counts, 7 tag words) stored to every device address in out_rows -- asynchronous (bp_dist_extend_count_rows)."""
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32).reshape(2 * self.dim)
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        rows = np.asarray(out_rows, dtype=np.uint64)
        self._ck(lib().bp_dist_extend_count_rows(self._h, sysb.ctypes.data, _dev_ptr(d_bounds), _dev_ptr(d_ids), n, spl.ctypes.data,
                                                 spl.shape[0], int(allow_fold), rows.ctypes.data, rows.shape[0]))

    def merge(self, other):
        self._ck(lib().bp_layer_merge(self._h, other._h))

    def sort(self):
        self._ck(lib().bp_layer_sort(self._h))

    par_sort = sort  # src/layer.rs:146-151: same result, the device sort is always parallel

    def _scan(self, flt, device):
        f = None if flt is None else flt._c()
        out, cnt = ctypes.c_void_p(), ctypes.c_size_t()
        fn = lib().bp_layer_scan_device if device else lib().bp_layer_scan
        self._ck(fn(self._h, None if f is None else ctypes.byref(f), ctypes.byref(out), ctypes.byref(cnt)))
        return out.value, cnt.value

    def scan_filtered(self, flt=None):
        """Returns the sorted, unique (later_id, earlier_id) pairs as an (P, 2) numpy array (a view of
        the layer's pinned result buffer, valid until the next call -- like the reference's borrow)."""
        p, n = self._scan(flt, device=False)
        if n == 0:
            return np.zeros((0, 2), dtype=self.id_dtype)
        buf = (ctypes.c_char * (n * 2 * self.id_bytes)).from_address(p)
        return np.frombuffer(buf, dtype=self.id_dtype).reshape(n, 2)

    def scan(self):
        return self.scan_filtered(None)

    par_scan = scan
    par_scan_filtered = scan_filtered

    def scan_device(self, flt=None):
        """Like scan_filtered but leaves the pairs on the device: returns (device pointer, count)."""
        return self._scan(flt, device=True)

    # ---- queries (Layer::test_box / test_ray, src/layer.rs:278-351), batched ----------------------
    def _query_batch(self, fn, system_bounds, params, width, max_depth):
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32).reshape(2 * self.dim)
        q = np.ascontiguousarray(params, dtype=np.float32).reshape(-1, width)
        pairs, offs, cnt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t()
        md = -1 if max_depth is None else int(max_depth)
        self._ck(fn(self._h, sysb.ctypes.data, q.ctypes.data, q.shape[0], md, 0, ctypes.byref(pairs), ctypes.byref(offs),
                    ctypes.byref(cnt)))
        nq, n = q.shape[0], cnt.value
        ob = (ctypes.c_char * (4 * (nq + 1))).from_address(offs.value)
        offsets = np.frombuffer(ob, dtype=np.uint32).copy()
        if n == 0:
            return offsets, np.zeros(0, dtype=self.id_dtype)
        pb = (ctypes.c_char * (n * 2 * self.id_bytes)).from_address(pairs.value)
        return offsets, np.frombuffer(pb, dtype=self.id_dtype).reshape(n, 2)[:, 1].copy()

    def test_box_batch(self, system_bounds, boxes, max_depth=None):
        """Layer::test_box for every row of `boxes` ((n, 2*D): min.., max..) in one call.  Returns (offsets, ids):
        ids[offsets[q]:offsets[q + 1]] is the sorted, duplicate-free ID list the reference returns for box q."""
        return self._query_batch(lib().bp_layer_test_box_batch, system_bounds, boxes, 2 * self.dim, max_depth)

    def test_ray_batch(self, system_bounds, rays, max_depth=None):
        """Layer::test_ray for every row of `rays` ((n, 2*D + 2): origin.., direction.., range_min, range_max)."""
        return self._query_batch(lib().bp_layer_test_ray_batch, system_bounds, rays, 2 * self.dim + 2, max_depth)

    def test_box(self, system_bounds, test_bounds, max_depth=None):
        """Layer::test_box -- src/layer.rs:278-299: the IDs whose cells overlap test_bounds, sorted, unique."""
        return self.test_box_batch(system_bounds, np.asarray(test_bounds, dtype=np.float32).reshape(1, -1), max_depth)[1]

    def test_ray(self, system_bounds, origin, direction, range_min, range_max, max_depth=None):
        """Layer::test_ray -- src/layer.rs:313-351."""
        ray = np.concatenate([np.asarray(origin, dtype=np.float32).reshape(-1), np.asarray(direction, dtype=np.float32).reshape(-1),
                              np.asarray([range_min, range_max], dtype=np.float32)])
        return self.test_ray_batch(system_bounds, ray.reshape(1, -1), max_depth)[1]

    PICK_DTYPE = np.dtype([("dist", np.float32), ("hit", np.uint32), ("id", np.uint64), ("point", np.float32, (3,)),
                           ("pad", np.uint32)])

    def pick_ray_batch(self, system_bounds, rays, max_dist, shape_kind, shapes, max_depth=None):
        """Layer::pick_ray (src/layer.rs:424-446) for every row of `rays` ((n, 2*D): origin.., direction..): the nearest
        object hit within max_dist, with the reference's get_dist closure replaced by a shape functor over `shapes`
        (rows indexed by ID: PICK_SPHERE centre.., radius; PICK_AABB min.., max..).  Returns a structured array
        (dist, hit, id, point[3]) with one entry per ray."""
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32).reshape(2 * self.dim)
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 2 * self.dim)
        width = self.dim + 1 if shape_kind == PICK_SPHERE else 2 * self.dim
        sh = np.ascontiguousarray(shapes, dtype=np.float32).reshape(-1, width)
        out = ctypes.c_void_p()
        self._ck(lib().bp_layer_pick_ray_batch(self._h, sysb.ctypes.data, r.ctypes.data, r.shape[0], float(max_dist),
                                               -1 if max_depth is None else int(max_depth), shape_kind, sh.ctypes.data, sh.shape[0], 0,
                                               ctypes.byref(out)))
        if r.shape[0] == 0:
            return np.zeros(0, dtype=self.PICK_DTYPE)
        buf = (ctypes.c_char * (r.shape[0] * self.PICK_DTYPE.itemsize)).from_address(out.value)
        return np.frombuffer(buf, dtype=self.PICK_DTYPE).copy()

    def pick_ray(self, system_bounds, origin, direction, max_dist, shape_kind, shapes, max_depth=None):
        """Layer::pick_ray: None, or (dist, id, point) of the nearest hit."""
        ray = np.concatenate([np.asarray(origin, dtype=np.float32).reshape(-1), np.asarray(direction, dtype=np.float32).reshape(-1)])
        res = self.pick_ray_batch(system_bounds, ray.reshape(1, -1), max_dist, shape_kind, shapes, max_depth)[0]
        if not res["hit"]:
            return None
        return float(res["dist"]), int(res["id"]), res["point"][:self.dim].copy()

    def iter(self):
        """Layer::iter: (keys, ids) numpy copies of the tree in its current order."""
        k, i, n, s = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int()
        self._ck(lib().bp_layer_records(self._h, ctypes.byref(k), ctypes.byref(i), ctypes.byref(n), ctypes.byref(s)))
        if n.value == 0:
            return np.zeros(0, dtype=self.key_dtype), np.zeros(0, dtype=self.id_dtype)
        kb = (ctypes.c_char * (n.value * np.dtype(self.key_dtype).itemsize)).from_address(k.value)
        ib = (ctypes.c_char * (n.value * self.id_bytes)).from_address(i.value)
        return np.frombuffer(kb, dtype=self.key_dtype).copy(), np.frombuffer(ib, dtype=self.id_dtype).copy()

    records = iter

    def clone(self):
        """Clone for Layer (src/layer.rs:597-617): min_depth and the tree with its sorted flag; the copy's result buffers
        start empty.  Goes through the host mirror of bp_layer_records -- cloning is not on the hot path."""
        keys, ids = self.iter()
        c = Layer(self.index, self.id_bytes, self.min_depth, device=getattr(self, "_device", -1))
        c.set_records(keys, ids, sorted_=self.sorted)
        return c

    def equals(self, other):
        """PartialEq for Layer (src/layer.rs:576-587): min_depth and `tree` = the (Index, ID) sequence AND its sorted flag."""
        if (self.index, self.id_bytes, self.min_depth, self.sorted) != (other.index, other.id_bytes, other.min_depth, other.sorted):
            return False
        (ka, ia), (kb, ib) = self.iter(), other.iter()
        return ka.shape == kb.shape and bool((ka == kb).all()) and bool((ia == ib).all())

    def records_device(self):
        k, i, n, s = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_int()
        self._ck(lib().bp_layer_records_device(self._h, ctypes.byref(k), ctypes.byref(i), ctypes.byref(n), ctypes.byref(s)))
        return k.value, i.value, n.value, bool(s.value)

    def set_records(self, keys, ids, sorted_=False, on_device=False, n=None, flagged=False):
        if on_device:
            self._ck(lib().bp_layer_set_records_flagged(self._h, _dev_ptr(keys), _dev_ptr(ids), n, int(sorted_), 1, int(flagged)))
        else:
            k = np.ascontiguousarray(keys, dtype=self.key_dtype)
            i = np.ascontiguousarray(ids, dtype=self.id_dtype)
            self._ck(lib().bp_layer_set_records(self._h, k.ctypes.data, i.ctypes.data, k.shape[0], int(sorted_), 0))

    # ---- multi-GPU building blocks (include/bp.h) ------------------------------------------------
    def set_halo(self, n_halo):
        self._ck(lib().bp_layer_set_halo(self._h, n_halo))

    def set_scan_dedup(self, enabled):
        self._ck(lib().bp_layer_set_scan_dedup(self._h, int(enabled)))

    def set_pair_later_fixed(self, fixed_bits):
        self._ck(lib().bp_layer_set_pair_later_fixed(self._h, int(fixed_bits)))

    def scan_raw_device(self, flt=None):
        """Raw (filtered, unsorted, duplicate-carrying) packed pairs: (device pointer, count)."""
        f = None if flt is None else flt._c()
        out, cnt = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(lib().bp_layer_scan_raw_device(self._h, None if f is None else ctypes.byref(f), ctypes.byref(out),
                                                ctypes.byref(cnt)))
        return out.value, cnt.value

    def unique_pairs_device(self, d_raw, n, id_mask=0):
        out, cnt = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(lib().bp_layer_unique_pairs_device(self._h, _dev_ptr(d_raw), n, id_mask, ctypes.byref(out), ctypes.byref(cnt)))
        return out.value, cnt.value

    def unique_pairs_inplace_device(self, d_raw, n, id_mask=0):
        """unique_pairs_device without the staging copy: d_raw is sort scratch afterwards."""
        out, cnt = ctypes.c_void_p(), ctypes.c_size_t()
        self._ck(lib().bp_layer_unique_pairs_inplace_device(self._h, _dev_ptr(d_raw), n, id_mask, ctypes.byref(out), ctypes.byref(cnt)))
        return out.value, cnt.value

    def sort_from_device(self, d_keys, d_ids, n, flagged, key_or, key_and, id_or, id_and, ids_ascending):
        """set_records + sort straight out of d_keys / d_ids with a caller-supplied digit plan (include/bp.h)."""
        m = 0xFFFFFFFFFFFFFFFF
        self._ck(lib().bp_layer_sort_from_device(self._h, _dev_ptr(d_keys), _dev_ptr(d_ids), n, int(flagged), int(key_or) & m,
                                                 int(key_and) & m, int(id_or) & m, int(id_and) & m, int(ids_ascending)))

    def id_order(self):
        """(first ID, last ID, ascending) of a tree built by extend calls since the last clear (include/bp.h)."""
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
        self._ck(lib().bp_layer_id_order(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, bool(c.value)

    def partition_records(self, d_keys, d_ids, n, splitters, d_out_keys, d_out_ids):
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        counts = np.zeros(spl.shape[0] + 1, dtype=np.uint64)
        self._ck(lib().bp_dist_partition_records(self._h, _dev_ptr(d_keys), _dev_ptr(d_ids), n, spl.ctypes.data, spl.shape[0],
                                                 _dev_ptr(d_out_keys), _dev_ptr(d_out_ids), counts.ctypes.data))
        return counts

    def partition_pairs(self, d_pairs, n, splitters, d_out_pairs):
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        counts = np.zeros(spl.shape[0] + 1, dtype=np.uint64)
        self._ck(lib().bp_dist_partition_pairs(self._h, _dev_ptr(d_pairs), n, spl.ctypes.data, spl.shape[0],
                                               _dev_ptr(d_out_pairs), counts.ctypes.data))
        return counts

    def count_records(self, d_keys, n, splitters):
        """(bucket sizes, halo copies per bucket) of a splitter partition of n records."""
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        counts = np.zeros(spl.shape[0] + 1, dtype=np.uint64)
        halo = np.zeros(spl.shape[0] + 1, dtype=np.uint64)
        self._ck(lib().bp_dist_count_records(self._h, _dev_ptr(d_keys), n, spl.ctypes.data, spl.shape[0], counts.ctypes.data,
                                             halo.ctypes.data))
        return counts, halo

    def count_records_rows(self, d_keys, n, splitters, tags, out_rows):
        """count_records with the result left on the device: row = [counts | halo counts | up to 8 tag words] (u64), stored
        to every device address in out_rows (this rank's row in every rank's count matrix); asynchronous."""
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        t = np.asarray([int(x) & 0xFFFFFFFFFFFFFFFF for x in tags], dtype=np.uint64)
        rows = np.asarray(out_rows, dtype=np.uint64)
        self._ck(lib().bp_dist_count_records_rows(self._h, _dev_ptr(d_keys), n, spl.ctypes.data, spl.shape[0],
                                                  t.ctypes.data, t.shape[0], rows.ctypes.data, rows.shape[0]))

    def count_pairs_rows(self, d_pairs, n, splitters, tags, out_rows):
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        t = np.asarray([int(x) & 0xFFFFFFFFFFFFFFFF for x in tags], dtype=np.uint64)
        rows = np.asarray(out_rows, dtype=np.uint64)
        self._ck(lib().bp_dist_count_pairs_rows(self._h, _dev_ptr(d_pairs), n, spl.ctypes.data, spl.shape[0],
                                                t.ctypes.data, t.shape[0], rows.ctypes.data, rows.shape[0]))

    def scatter_records(self, d_keys, d_ids, n, splitters, dst_keys, dst_ids, halo_dst_keys=None, halo_dst_ids=None,
                        fold_cell_flags=False):
        """Partition pass writing bucket b to the device addresses dst_keys[b] / dst_ids[b]; fold_cell_flags: the layer's
        cell flags leave in the top 3 bits of the IDs (bp_dist_scatter_records_flagged)."""
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (dst_keys, dst_ids)]
        h = [None if x is None else np.ascontiguousarray(x, dtype=np.uint64) for x in (halo_dst_keys, halo_dst_ids)]
        self._ck(lib().bp_dist_scatter_records_flagged(self._h, _dev_ptr(d_keys), _dev_ptr(d_ids), n, spl.ctypes.data, spl.shape[0],
                                                       arrs[0].ctypes.data, arrs[1].ctypes.data,
                                                       None if h[0] is None else h[0].ctypes.data,
                                                       None if h[1] is None else h[1].ctypes.data, int(fold_cell_flags)))

    def count_pairs(self, d_pairs, n, splitters):
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        counts = np.zeros(spl.shape[0] + 1, dtype=np.uint64)
        self._ck(lib().bp_dist_count_pairs(self._h, _dev_ptr(d_pairs), n, spl.ctypes.data, spl.shape[0], counts.ctypes.data))
        return counts

    def scatter_pairs(self, d_pairs, n, splitters, dst_pairs):
        spl = np.ascontiguousarray(splitters, dtype=np.uint64)
        dst = np.ascontiguousarray(dst_pairs, dtype=np.uint64)
        self._ck(lib().bp_dist_scatter_pairs(self._h, _dev_ptr(d_pairs), n, spl.ctypes.data, spl.shape[0], dst.ctypes.data))

    def lookup_ranges(self, d_sorted_keys, n, queries):
        q = np.ascontiguousarray(queries, dtype=np.uint64)
        lo = np.zeros(q.shape[0], dtype=np.uint64)
        hi = np.zeros(q.shape[0], dtype=np.uint64)
        self._ck(lib().bp_dist_lookup_ranges(self._h, _dev_ptr(d_sorted_keys), n, q.ctypes.data, q.shape[0], lo.ctypes.data,
                                             hi.ctypes.data))
        return lo, hi

    def __len__(self):
        n = ctypes.c_size_t()
        self._ck(lib().bp_layer_len(self._h, ctypes.byref(n)))
        return n.value

    @property
    def sorted(self):
        s = ctypes.c_int()
        self._ck(lib().bp_layer_is_sorted(self._h, ctypes.byref(s)))
        return bool(s.value)

    @property
    def min_depth(self):
        d = ctypes.c_uint32()
        self._ck(lib().bp_layer_min_depth(self._h, ctypes.byref(d)))
        return d.value

    def masks(self):
        """(key_or, key_and, id_or, id_and) over the tree."""
        v = [ctypes.c_uint64() for _ in range(4)]
        self._ck(lib().bp_layer_masks(self._h, *[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)

    # ---- instrumentation ------------------------------------------------------------------------
    def set_profiling(self, enabled):
        self._ck(lib().bp_layer_set_profiling(self._h, int(enabled)))

    def reset_stats(self):
        self._ck(lib().bp_layer_reset_stats(self._h))

    def stats(self):
        s = Stats()
        self._ck(lib().bp_layer_stats(self._h, ctypes.byref(s)))
        d = {k: getattr(s, k) for k in ("n_records", "n_invalid", "n_work_items", "n_raw_pairs", "n_pairs",
                                        "sort_passes", "pair_sort_passes", "merged", "rescans", "launches_total")}
        d["launches"] = {c: s.launches[i] for i, c in enumerate(_lib.KERNEL_CLASSES)}
        d["kernel_ms"] = {c: s.kernel_ms[i] for i, c in enumerate(_lib.KERNEL_CLASSES)}
        d["algo_bytes"] = {c: s.algo_bytes[i] for i, c in enumerate(_lib.KERNEL_CLASSES)}
        return d


class LayerBuilder:
    """src/layer.rs:620-696."""

    def __init__(self):
        self._min_depth = 0
        self._index_capacity = 0
        self._collision_capacity = 0
        self._test_capacity = 0
        self._device = -1

    @staticmethod
    def new():
        return LayerBuilder()

    def with_min_depth(self, depth):
        self._min_depth = depth
        return self

    def with_index_capacity(self, capacity):
        self._index_capacity = capacity
        return self

    def with_collision_capacity(self, capacity):
        self._collision_capacity = capacity
        return self

    def with_test_capacity(self, capacity):
        self._test_capacity = capacity
        return self

    def with_device(self, device):
        self._device = device
        return self

    def build(self, index=Index64_3D, id_type="u32"):
        return Layer(index, id_type, self._min_depth, self._index_capacity, self._collision_capacity,
                     self._test_capacity, self._device)


def plan_radix_passes(mask):
    sh = (ctypes.c_uint32 * 16)()
    bt = (ctypes.c_uint32 * 16)()
    n = lib().bp_plan_radix_passes(mask, sh, bt, 16)
    out = []
    for i in range(n):  # (shift, bits) or, with a second bit-field, (shift, bits, shift2, bits2)
        f = (sh[i] & 0xFFFF, bt[i] & 0xFFFF)
        if bt[i] >> 16:
            f += (sh[i] >> 16, bt[i] >> 16)
        out.append(f)
    return out


def plan_sort_finish(mask, n_records):
    """The "top bits + finish" plan of a record sort (bp_plan_sort_finish): None when the plain radix plan is kept, else
    (top_mask, group_shift): radix passes over top_mask, then one pass that orders the groups of equal key >> group_shift."""
    top = ctypes.c_uint64()
    gs = ctypes.c_uint32()
    if not lib().bp_plan_sort_finish(mask, n_records, ctypes.byref(top), ctypes.byref(gs)):
        return None
    return top.value, gs.value


def plan_dist_splitters(sample, parts):
    """parts - 1 splitters of the sharded frame from a gathered key sample (bp_dist_plan_splitters; host only)."""
    s = np.ascontiguousarray(np.asarray(sample, dtype=np.uint64)).copy()  # (sorted in place by the library)
    out = np.zeros(max(parts - 1, 1), dtype=np.uint64)
    st = lib().bp_dist_plan_splitters(s.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), s.shape[0], parts,
                                      out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    if st != 0:
        raise BpError(st, "bp_dist_plan_splitters")
    return out[:parts - 1]


def plan_dist_shard_bits(splitters, shard, top=0xFFFFFFFFFFFFFFFF):
    """(fixed, value): the key bits every record of `shard` shares, from its two splitters (bp_dist_plan_shard_bits)."""
    spl = np.ascontiguousarray(np.asarray(splitters, dtype=np.uint64))
    fixed, value = ctypes.c_uint64(), ctypes.c_uint64()
    st = lib().bp_dist_plan_shard_bits(spl.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), spl.shape[0] + 1, shard, top,
                                       ctypes.byref(fixed), ctypes.byref(value))
    if st != 0:
        raise BpError(st, "bp_dist_plan_shard_bits")
    return fixed.value, value.value


def device_count():
    n = ctypes.c_int()
    st = lib().bp_device_count(ctypes.byref(n))
    return n.value if st == 0 else 0
