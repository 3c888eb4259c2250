"""cpu_oracle.py -- ctypes loader for the C++ CPU oracle (oracle/libbp_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbp_oracle.so")

INDEX32_2D, INDEX64_2D, INDEX64_3D = 0, 1, 2
PICK_SPHERE, PICK_AABB = 0, 1
FILTER_NONE, FILTER_ID_PARITY, FILTER_XOR_MASK, FILTER_CATEGORY, FILTER_SPHERES = 0, 1, 2, 3, 4
DIM = {INDEX32_2D: 2, INDEX64_2D: 2, INDEX64_3D: 3}

_lib = None


def build(force=False):
    """Compile the oracle with its Makefile (g++, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("bp_oracle.cpp", "bp_oracle.h", "Makefile")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B", "libbp_oracle.so"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_SO)
    vp, sz, u32, u64, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    sigs = {
        "bpo_layer_new": (vp, [i32, i32, u32]),
        "bpo_layer_free": (None, [vp]),
        "bpo_layer_clear": (None, [vp]),
        "bpo_layer_extend": (None, [vp, vp, vp, vp, sz]),
        "bpo_layer_merge": (None, [vp, vp]),
        "bpo_layer_sort": (None, [vp]),
        "bpo_layer_par_sort": (None, [vp]),
        "bpo_layer_scan": (sz, [vp, i32, u64, vp, sz]),
        "bpo_layer_par_scan": (sz, [vp, i32, u64, vp, sz]),
        "bpo_layer_len": (sz, [vp]),
        "bpo_layer_sorted": (i32, [vp]),
        "bpo_layer_min_depth": (u32, [vp]),
        "bpo_layer_num_collisions": (sz, [vp]),
        "bpo_layer_num_raw_collisions": (sz, [vp]),
        "bpo_layer_records": (None, [vp, vp, vp]),
        "bpo_layer_collisions": (None, [vp, vp, vp]),
        "bpo_layer_set_records": (None, [vp, vp, vp, sz, i32]),
        "bpo_layer_test_box": (sz, [vp, vp, vp, i32]),
        "bpo_layer_test_ray": (sz, [vp, vp, vp, i32]),
        "bpo_layer_test_results": (None, [vp, vp]),
        "bpo_layer_pick_ray": (i32, [vp, vp, vp, ctypes.c_float, i32, i32, vp, sz, vp, vp]),
        "bpo_encode_axis": (u64, [i32, u32]),
        "bpo_decode_axis": (u32, [i32, u64]),
        "bpo_make_index": (u64, [i32, u32, vp]),
        "bpo_level_mask": (u64, [i32, u32]),
        "bpo_overlaps": (i32, [i32, u64, u64]),
        "bpo_to_local": (None, [i32, vp, vp, vp]),
        "bpo_to_global": (None, [i32, vp, vp, vp]),
        "bpo_max_threads": (i32, []),
        "bpo_set_threads": (None, [i32]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleLayer:
    """Mirror of the reference's Layer<Index, ID> (src/layer.rs:42-68) on the CPU oracle."""

    def __init__(self, kind, id_bytes=4, min_depth=0):
        self.kind, self.id_bytes, self.dim = kind, id_bytes, DIM[kind]
        self.id_dtype = np.uint32 if id_bytes == 4 else np.uint64
        self._h = lib().bpo_layer_new(kind, id_bytes, min_depth)
        if not self._h:
            raise ValueError("bad kind / id_bytes")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bpo_layer_free(self._h)
            self._h = None

    def clear(self):
        lib().bpo_layer_clear(self._h)

    def extend(self, system_bounds, bounds, ids):
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32).reshape(2 * self.dim)
        b = np.ascontiguousarray(bounds, dtype=np.float32).reshape(-1, 2 * self.dim)
        i = np.ascontiguousarray(ids, dtype=self.id_dtype)
        assert i.shape[0] == b.shape[0]
        lib().bpo_layer_extend(self._h, _ptr(sysb), _ptr(b), _ptr(i), b.shape[0])

    def merge(self, other):
        lib().bpo_layer_merge(self._h, other._h)

    def sort(self):
        lib().bpo_layer_sort(self._h)

    def par_sort(self):
        lib().bpo_layer_par_sort(self._h)

    def _scan(self, fn, filter_kind, filter_arg, table):
        if table is None:
            t = None
        elif filter_kind == FILTER_SPHERES:
            t = np.ascontiguousarray(table, dtype=np.float32).reshape(-1, 4)
        else:
            t = np.ascontiguousarray(table, dtype=np.uint32).reshape(-1, 2)
        fn(self._h, filter_kind, filter_arg, None if t is None else _ptr(t), 0 if t is None else t.shape[0])
        return self.collisions()

    def scan(self, filter_kind=FILTER_NONE, filter_arg=0, table=None):
        return self._scan(lib().bpo_layer_scan, filter_kind, filter_arg, table)

    def par_scan(self, filter_kind=FILTER_NONE, filter_arg=0, table=None):
        return self._scan(lib().bpo_layer_par_scan, filter_kind, filter_arg, table)

    def __len__(self):
        return lib().bpo_layer_len(self._h)

    @property
    def sorted(self):
        return bool(lib().bpo_layer_sorted(self._h))

    @property
    def min_depth(self):
        return lib().bpo_layer_min_depth(self._h)

    @property
    def num_raw_collisions(self):
        return lib().bpo_layer_num_raw_collisions(self._h)

    def records(self):
        n = len(self)
        k = np.zeros(n, dtype=np.uint64)
        i = np.zeros(n, dtype=np.uint64)
        if n:
            lib().bpo_layer_records(self._h, _ptr(k), _ptr(i))
        return k, i

    def set_records(self, keys, ids, sorted_=False):
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        i = np.ascontiguousarray(ids, dtype=np.uint64)
        lib().bpo_layer_set_records(self._h, _ptr(k), _ptr(i), k.shape[0], int(sorted_))

    def _test(self, fn, system_bounds, params, max_depth):
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32)
        q = np.ascontiguousarray(params, dtype=np.float32)
        n = fn(self._h, _ptr(sysb), _ptr(q), -1 if max_depth is None else int(max_depth))
        out = np.zeros(n, dtype=np.uint64)
        if n:
            lib().bpo_layer_test_results(self._h, _ptr(out))
        return out

    def test_box(self, system_bounds, test_bounds, max_depth=None):
        """Layer::test_box (src/layer.rs:293-311): sorted unique IDs, widened to u64."""
        return self._test(lib().bpo_layer_test_box, system_bounds, test_bounds, max_depth)

    def test_ray(self, system_bounds, origin, direction, range_min, range_max, max_depth=None):
        """Layer::test_ray (src/layer.rs:326-351)."""
        ray = np.concatenate([np.asarray(origin, dtype=np.float32).reshape(-1), np.asarray(direction, dtype=np.float32).reshape(-1),
                              np.asarray([range_min, range_max], dtype=np.float32)])
        return self._test(lib().bpo_layer_test_ray, system_bounds, ray, max_depth)

    def pick_ray(self, system_bounds, origin, direction, max_dist, shape_kind, shapes, max_depth=None):
        """Layer::pick_ray (src/layer.rs:424-446) with a shape functor for get_dist: None or (dist, id, point)."""
        sysb = np.ascontiguousarray(system_bounds, dtype=np.float32)
        ray = np.concatenate([np.asarray(origin, dtype=np.float32).reshape(-1), np.asarray(direction, dtype=np.float32).reshape(-1)])
        sh = np.ascontiguousarray(shapes, dtype=np.float32)
        dim = ray.shape[0] // 2
        out = np.zeros(1 + dim, dtype=np.float32)
        oid = np.zeros(1, dtype=np.uint64)
        hit = lib().bpo_layer_pick_ray(self._h, _ptr(sysb), _ptr(ray), float(max_dist), -1 if max_depth is None else int(max_depth),
                                       int(shape_kind), _ptr(sh), sh.shape[0], _ptr(out), _ptr(oid))
        return (float(out[0]), int(oid[0]), out[1:].copy()) if hit else None

    def collisions(self):
        n = lib().bpo_layer_num_collisions(self._h)
        a = np.zeros(n, dtype=np.uint64)
        b = np.zeros(n, dtype=np.uint64)
        if n:
            lib().bpo_layer_collisions(self._h, _ptr(a), _ptr(b))
        return np.stack([a, b], axis=1)


def encode_axis(kind, v):
    return lib().bpo_encode_axis(kind, int(v))


def decode_axis(kind, o):
    return lib().bpo_decode_axis(kind, int(o))


def make_index(kind, depth, origin):
    o = np.ascontiguousarray(origin, dtype=np.uint32)
    return lib().bpo_make_index(kind, int(depth), _ptr(o))


def level_mask(kind, depth):
    return lib().bpo_level_mask(kind, int(depth))


def overlaps(kind, a, b):
    return bool(lib().bpo_overlaps(kind, int(a), int(b)))


def to_local(dim, system_bounds, bounds):
    sysb = np.ascontiguousarray(system_bounds, dtype=np.float32)
    b = np.ascontiguousarray(bounds, dtype=np.float32)
    out = np.zeros(2 * dim, dtype=np.uint32)
    lib().bpo_to_local(dim, _ptr(sysb), _ptr(b), _ptr(out))
    return out


def to_global(dim, system_bounds, local):
    sysb = np.ascontiguousarray(system_bounds, dtype=np.float32)
    l = np.ascontiguousarray(local, dtype=np.uint32)
    out = np.zeros(2 * dim, dtype=np.float32)
    lib().bpo_to_global(dim, _ptr(sysb), _ptr(l), _ptr(out))
    return out
