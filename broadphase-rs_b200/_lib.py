"""ctypes binding of libbroadphase_b200.so (include/bp.h).  Fails loudly if the library is missing:
there is no CPU fallback and no other backend."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libbroadphase_b200.so")

BP_OK = 0
STATUS = {0: "BP_OK", 1: "BP_ERR_INVALID_ARG", 2: "BP_ERR_CUDA", 3: "BP_ERR_OOM", 4: "BP_ERR_TOO_LARGE",
          5: "BP_ERR_INTERNAL", 6: "BP_ERR_MISMATCH"}
BP_K_COUNT = 13
KERNEL_CLASSES = ["encode", "sort_hist", "sort_pass", "merge", "scan_runs", "scan_emit", "pair_hist", "pair_pass",
                  "pair_unique", "misc", "query", "partition", "sort_finish"]


class BpError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__("%s: %s" % (STATUS.get(status, status), message))


class LayerConfig(ctypes.Structure):
    _fields_ = [("index_kind", ctypes.c_int32), ("id_bytes", ctypes.c_int32), ("min_depth", ctypes.c_uint32),
                ("device", ctypes.c_int32), ("index_capacity", ctypes.c_size_t),
                ("collision_capacity", ctypes.c_size_t), ("test_capacity", ctypes.c_size_t)]


class Filter(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("table_on_device", ctypes.c_int32), ("arg", ctypes.c_uint64),
                ("table", ctypes.c_void_p), ("n_table", ctypes.c_size_t)]


class Stats(ctypes.Structure):
    _fields_ = [("n_records", ctypes.c_uint64), ("n_invalid", ctypes.c_uint64), ("n_work_items", ctypes.c_uint64),
                ("n_raw_pairs", ctypes.c_uint64), ("n_pairs", ctypes.c_uint64), ("sort_passes", ctypes.c_uint32),
                ("pair_sort_passes", ctypes.c_uint32), ("merged", ctypes.c_uint32), ("rescans", ctypes.c_uint32),
                ("launches_total", ctypes.c_uint64), ("launches", ctypes.c_uint64 * BP_K_COUNT),
                ("kernel_ms", ctypes.c_double * BP_K_COUNT), ("algo_bytes", ctypes.c_double * BP_K_COUNT)]


BP_DIST_PHASES = 9
DIST_PHASE_NAMES = ["encode", "splitters", "counts", "exchange", "sort", "scan", "pair_counts", "pair_exchange", "unique"]
DIST_OPT_REUSE_SPLITTERS, DIST_OPT_FUSE_COUNTS, DIST_OPT_GLOBAL_DEDUP_DECISION, DIST_OPT_TRACE = 0, 1, 2, 3


class DistConfig(ctypes.Structure):
    _fields_ = [("index_kind", ctypes.c_int32), ("min_depth", ctypes.c_uint32), ("device", ctypes.c_int32),
                ("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("record_capacity", ctypes.c_size_t),
                ("pair_capacity", ctypes.c_size_t)]


class DistInfo(ctypes.Structure):
    _fields_ = [("records_local", ctypes.c_uint64), ("records_owned", ctypes.c_uint64), ("n_halo", ctypes.c_uint64),
                ("raw_pairs", ctypes.c_uint64), ("pairs", ctypes.c_uint64), ("records_needed", ctypes.c_uint64),
                ("pairs_needed", ctypes.c_uint64), ("fused", ctypes.c_int32), ("rescanned", ctypes.c_int32),
                ("rebalance_records", ctypes.c_int32), ("rebalance_pairs", ctypes.c_int32),
                ("phase_ms", ctypes.c_double * BP_DIST_PHASES)]


# every symbol include/bp.h declares: name -> (restype, argtypes)
_vp, _sz, _i, _u32, _u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
_P = ctypes.POINTER
SYMBOLS = {
    "bp_layer_create": (_i, [_P(LayerConfig), _P(_vp)]),
    "bp_layer_destroy": (_i, [_vp]),
    "bp_layer_set_stream": (_i, [_vp, _vp]),
    "bp_layer_clear": (_i, [_vp]),
    "bp_layer_extend_host": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "bp_layer_extend_device": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "bp_layer_merge": (_i, [_vp, _vp]),
    "bp_layer_sort": (_i, [_vp]),
    "bp_layer_scan": (_i, [_vp, _P(Filter), _P(_vp), _P(_sz)]),
    "bp_layer_scan_device": (_i, [_vp, _P(Filter), _P(_vp), _P(_sz)]),
    "bp_layer_test_box_batch": (_i, [_vp, _vp, _vp, _sz, ctypes.c_int32, _i, _P(_vp), _P(_vp), _P(_sz)]),
    "bp_layer_test_ray_batch": (_i, [_vp, _vp, _vp, _sz, ctypes.c_int32, _i, _P(_vp), _P(_vp), _P(_sz)]),
    "bp_layer_pick_ray_batch": (_i, [_vp, _vp, _vp, _sz, ctypes.c_float, ctypes.c_int32, ctypes.c_int32, _vp, _sz, _i, _P(_vp)]),
    "bp_layer_records": (_i, [_vp, _P(_vp), _P(_vp), _P(_sz), _P(_i)]),
    "bp_layer_records_device": (_i, [_vp, _P(_vp), _P(_vp), _P(_sz), _P(_i)]),
    "bp_layer_set_records": (_i, [_vp, _vp, _vp, _sz, _i, _i]),
    "bp_layer_set_records_flagged": (_i, [_vp, _vp, _vp, _sz, _i, _i, _i]),
    "bp_layer_sort_from_device": (_i, [_vp, _vp, _vp, _sz, _i, _u64, _u64, _u64, _u64, _i]),
    "bp_layer_id_order": (_i, [_vp, _P(_u64), _P(_u64), _P(_i)]),
    "bp_dist_count_records_rows": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _i, _vp, _i]),
    "bp_dist_extend_count_rows": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _vp, _i]),
    "bp_dist_count_pairs_rows": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _i, _vp, _i]),
    "bp_dist_scatter_records_flagged": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, _vp, _vp, _i]),
    "bp_layer_unique_pairs_inplace_device": (_i, [_vp, _vp, _sz, _u64, _P(_vp), _P(_sz)]),
    "bp_dist_partition_records": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, _vp]),
    "bp_dist_partition_pairs": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _vp]),
    "bp_dist_lookup_ranges": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _vp]),
    "bp_dist_count_records": (_i, [_vp, _vp, _sz, _vp, _i, _vp, _vp]),
    "bp_dist_scatter_records": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, _vp, _vp]),
    "bp_dist_count_pairs": (_i, [_vp, _vp, _sz, _vp, _i, _vp]),
    "bp_dist_scatter_pairs": (_i, [_vp, _vp, _sz, _vp, _i, _vp]),
    "bp_dist_create": (_i, [_P(DistConfig), _P(_vp)]),
    "bp_dist_destroy": (_i, [_vp]),
    "bp_dist_handle_bytes": (_sz, []),
    "bp_dist_export": (_i, [_vp, _vp]),
    "bp_dist_connect": (_i, [_vp, _vp]),
    "bp_dist_set_stream": (_i, [_vp, _vp]),
    "bp_dist_set_option": (_i, [_vp, _i, _i]),
    "bp_dist_set_static": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "bp_dist_frame": (_i, [_vp, _vp, _vp, _vp, _sz, _P(Filter), _P(_vp), _P(_sz)]),
    "bp_dist_last_info": (_i, [_vp, _P(DistInfo)]),
    "bp_dist_layer": (_vp, [_vp, _i]),
    "bp_dist_last_error": (ctypes.c_char_p, [_vp]),
    "bp_layer_set_halo": (_i, [_vp, _sz]),
    "bp_layer_set_scan_dedup": (_i, [_vp, _i]),
    "bp_layer_set_pair_later_fixed": (_i, [_vp, _u64]),
    "bp_layer_scan_raw_device": (_i, [_vp, _P(Filter), _P(_vp), _P(_sz)]),
    "bp_layer_unique_pairs_device": (_i, [_vp, _vp, _sz, _u64, _P(_vp), _P(_sz)]),
    "bp_layer_len": (_i, [_vp, _P(_sz)]),
    "bp_layer_is_sorted": (_i, [_vp, _P(_i)]),
    "bp_layer_min_depth": (_i, [_vp, _P(_u32)]),
    "bp_layer_masks": (_i, [_vp, _P(_u64), _P(_u64), _P(_u64), _P(_u64)]),
    "bp_layer_set_profiling": (_i, [_vp, _i]),
    "bp_layer_reset_stats": (_i, [_vp]),
    "bp_layer_stats": (_i, [_vp, _P(Stats)]),
    "bp_layer_last_error": (ctypes.c_char_p, [_vp]),
    "bp_status_string": (ctypes.c_char_p, [_i]),
    "bp_version": (_i, []),
    "bp_device_count": (_i, [_P(_i)]),
    "bp_plan_radix_passes": (_i, [_u64, _P(_u32), _P(_u32), _i]),
    "bp_plan_sort_finish": (_i, [_u64, _u64, _P(_u64), _P(_u32)]),
    "bp_dist_plan_splitters": (_i, [_P(_u64), _sz, _i, _P(_u64)]),
    "bp_dist_plan_shard_bits": (_i, [_P(_u64), _i, _i, _u64, _P(_u64), _P(_u64)]),
}

_lib = None


def lib():
    """Loads the CUDA library.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "libbroadphase_b200.so is missing at %s: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'`.  There is no CPU fallback." % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status, handle=None):
    if status != BP_OK:
        msg = ""
        if handle:
            m = lib().bp_layer_last_error(handle)
            msg = m.decode() if m else ""
        raise BpError(status, msg)
