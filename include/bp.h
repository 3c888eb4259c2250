/*
 * bp.h -- C ABI of libbroadphase_b200.so, the B200-native (sm_100a CUDA) implementation of the
 * broadphase-rs hot path: Layer::{clear, extend, merge, sort/par_sort, scan/par_scan,
 * scan_filtered/par_scan_filtered, iter} for Index32_2D / Index64_2D / Index64_3D.
 *
 * The reference (zvxryb/broadphase-rs, crate zvxryb-broadphase 0.1.2) has no FFI of its own; its
 * boundary is the generic Rust type Layer<Index, ID> (src/layer.rs:42-47, re-exported at
 * src/lib.rs:80-82).  Each entry point below names the reference method it replaces (file:line
 * relative to the reference root).  A Rust `Layer<Index, ID>` shim binds exactly these symbols
 * (see INTEGRATION.md); the C++ mirror is broadphase-rs_b200/cpp/broadphase/layer.hpp and the
 * Python mirror (used by the tests) is broadphase-rs_b200/layer.py.
 *
 * Conventions
 *  - Every function returns a bp_status (0 = BP_OK).  The reference never returns errors: it
 *    drops out-of-bounds objects silently (src/layer.rs:108-111) and aborts on allocation failure;
 *    here allocation/CUDA failures are status codes and bp_layer_last_error() has the text.
 *  - A layer is single-caller (`&mut self` in the reference): no internal locking.
 *  - Result pointers (records, pairs) are owned by the layer and stay valid until the next
 *    mutating call on it, like the `&'a Vec<(ID, ID)>` the reference returns.
 *  - There is no CPU fallback: without a CUDA device bp_layer_create fails with BP_ERR_CUDA.
 *  - Limits (BP_ERR_TOO_LARGE): fewer than 2^30 records per layer; a scan may visit fewer than 2^30 (ancestor, descendant)
 *    record pairs -- about 46 000 objects sharing ONE cell, or a depth-0 cell, reach that -- and emits fewer than 2^30
 *    raw pairs; an object may own at most 2^20 cells and 4095 per axis (only a min_depth far above its natural depth does
 *    that; the reference warn!s and continues, src/geom.rs:299-301): the extend call that contains such an object is
 *    rejected as a whole -- the error comes back from that call or from the next call on the layer, whose tree is what
 *    it was before the rejected call.
 */
#ifndef BP_H
#define BP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_VERSION 100 /* 0.1.0 */

typedef enum bp_status {
    BP_OK = 0,
    BP_ERR_INVALID_ARG = 1,
    BP_ERR_CUDA = 2,      /* a CUDA runtime call failed (including "no device") */
    BP_ERR_OOM = 3,       /* device or pinned-host allocation failed */
    BP_ERR_TOO_LARGE = 4, /* more than 2^30 records / raw pairs */
    BP_ERR_INTERNAL = 5,  /* a kernel reported an inconsistency (e.g. look-back time-out) */
    BP_ERR_MISMATCH = 6   /* layers of different Index / ID types or devices */
} bp_status;

/* Index types -- src/index.rs:293-295 */
typedef enum bp_index_kind {
    BP_INDEX32_2D = 0, /* u32 key: 4 depth bits,  2 x 14 origin bits; bounds are 4 floats */
    BP_INDEX64_2D = 1, /* u64 key: 5 depth bits,  2 x 29 origin bits; bounds are 4 floats */
    BP_INDEX64_3D = 2  /* u64 key: 5 depth bits,  3 x 19 origin bits; bounds are 6 floats */
} bp_index_kind;

/* Device functors standing in for scan_filtered's `F: FnMut(ID, ID) -> bool` (src/layer.rs:456-460).
 * The filter is called as filter(later_id, earlier_id) before duplicate removal, like the reference.
 * Rust closures cannot cross the ABI; stateful filters are unsupported. */
typedef enum bp_filter_kind {
    BP_FILTER_NONE = 0,      /* |_, _| true  (scan / par_scan) */
    BP_FILTER_ID_PARITY = 1, /* ((a ^ b) & 1) == 1 */
    BP_FILTER_XOR_MASK = 2,  /* ((a ^ b) & arg) != 0 */
    BP_FILTER_CATEGORY = 3,  /* (cat[a] & msk[b]) != 0 && (cat[b] & msk[a]) != 0; table = n_table x {u32 cat, u32 msk}
                                indexed by ID; IDs >= n_table act as all-ones */
    BP_FILTER_SPHERES = 4    /* fused narrow phase: the pair test of the reference's example,
                                examples/main.rs:461-479, `!(|c[b] - c[a]| > r[a] + r[b])`, as the scan filter, so that only
                                touching spheres leave the device.  table = n_table x {f32 x, y, z, r} indexed by ID (z = 0
                                for the 2-D index types; the same 4 bytes per entry, passed through the u32 pointer);
                                IDs >= n_table pass */
} bp_filter_kind;

typedef struct bp_filter {
    int32_t kind;          /* bp_filter_kind */
    int32_t table_on_device; /* 1: `table` is a device pointer; 0: host pointer (copied per call) */
    uint64_t arg;
    const uint32_t *table;
    size_t n_table;
} bp_filter;

/* LayerBuilder -- src/layer.rs:620-696 */
typedef struct bp_layer_config {
    int32_t index_kind;        /* bp_index_kind */
    int32_t id_bytes;          /* 4 (u32 IDs) or 8 (u64 IDs) -- ObjectID, src/traits.rs:6-16 */
    uint32_t min_depth;        /* with_min_depth         src/layer.rs:646-649 */
    int32_t device;            /* CUDA device ordinal; -1 = current device */
    size_t index_capacity;     /* with_index_capacity    src/layer.rs:652-655 (records) */
    size_t collision_capacity; /* with_collision_capacity src/layer.rs:658-661 (pairs) */
    size_t test_capacity;      /* with_test_capacity     src/layer.rs:664-667 (accepted; query buffers grow on demand) */
} bp_layer_config;

typedef struct bp_layer bp_layer;

/* kernel classes for bp_stats */
enum {
    BP_K_ENCODE = 0,    /* fused quantise + encode (extend) */
    BP_K_SORT_HIST = 1, /* record sort: digit histograms */
    BP_K_SORT_PASS = 2, /* record sort: one onesweep pass */
    BP_K_MERGE = 3,     /* merge-path merge of two sorted runs */
    BP_K_SCAN_RUNS = 4, /* scan: descendant-run lengths + compaction */
    BP_K_SCAN_EMIT = 5, /* scan: load-balanced pair emission */
    BP_K_PAIR_HIST = 6, /* pair sort: digit histograms */
    BP_K_PAIR_PASS = 7, /* pair sort: one onesweep pass */
    BP_K_PAIR_UNIQUE = 8, /* dedup + final pair layout */
    BP_K_MISC = 9,      /* small helpers (histogram scans, masks, gathers) */
    BP_K_QUERY = 10,    /* batched box / ray queries: hierarchy descent (count pass + write pass) */
    BP_K_PARTITION = 11, /* multi-GPU: splitter counts + the partition pass that stores into the peers' receive buffers */
    BP_K_SORT_FINISH = 12, /* record sort: the pass that orders the low key bits group by group after radix passes over the top bits */
    BP_K_COUNT = 13
};

typedef struct bp_stats {
    uint64_t n_records;       /* tree length */
    uint64_t n_invalid;       /* objects rejected by contains() since the last scan (src/layer.rs:108-111) */
    uint64_t n_work_items;    /* (ancestor, descendant) record pairs visited by the last scan */
    uint64_t n_raw_pairs;     /* pairs emitted before sort + dedup in the last scan */
    uint64_t n_pairs;         /* unique pairs returned by the last scan */
    uint32_t sort_passes;     /* radix passes of the last record sort (0 if merge-only / clean) */
    uint32_t pair_sort_passes;
    uint32_t merged;          /* 1 if the last sort finished with a merge-path merge */
    uint32_t rescans;         /* 1 if the last scan needed the same-ID (inactive) re-emission */
    uint64_t launches_total;  /* kernels launched by this layer since creation */
    /* per kernel class, accumulated since bp_layer_reset_stats(); times only when profiling is on */
    uint64_t launches[BP_K_COUNT];
    double kernel_ms[BP_K_COUNT];
    double algo_bytes[BP_K_COUNT]; /* algorithmic bytes summed over those launches */
} bp_stats;

/* ---- lifetime ------------------------------------------------------------------------------- */

/* LayerBuilder::build -- src/layer.rs:670-695.  Starts empty with the sorted flag set (:681). */
int bp_layer_create(const bp_layer_config *config, bp_layer **out);
int bp_layer_destroy(bp_layer *layer);
/* Run this layer's work on a caller-owned cudaStream_t (default: a stream the layer creates). */
int bp_layer_set_stream(bp_layer *layer, void *cuda_stream);

/* ---- Layer methods ---------------------------------------------------------------------------- */

/* Layer::clear -- src/layer.rs:84-88: empties the tree, sets the sorted flag. */
int bp_layer_clear(bp_layer *layer);

/* Layer::extend -- src/layer.rs:94-121.  system_bounds: 2*D floats (min.., max..) on the host;
 * bounds: n x 2*D floats (min.., max.. like Bounds{min, max}, src/geom.rs:84-87); ids: n x id_bytes.
 * Appends, for every object inside system_bounds, one record per cell in the reference's order
 * (object order; z outermost, then y, x innermost).  _host takes host pointers (copied to the
 * device inside the call), _device takes device pointers (bounds 16-byte aligned). */
int bp_layer_extend_host(bp_layer *layer, const float *system_bounds, const float *bounds, const void *ids, size_t n);
int bp_layer_extend_device(bp_layer *layer, const float *system_bounds, const float *d_bounds, const void *d_ids, size_t n);

/* Layer::merge -- src/layer.rs:127-138: min_depth = min(self, other); appends other's tree
 * verbatim; clears the sorted flag.  When both trees were sorted the next sort is a merge-path
 * merge instead of a full re-sort. */
int bp_layer_merge(bp_layer *layer, const bp_layer *other);

/* Layer::sort and Layer::par_sort -- src/layer.rs:146-165: no-op when the flag is set, else sorts
 * the tree by (Index, ID) and sets the flag. */
int bp_layer_sort(bp_layer *layer);

/* Layer::scan / scan_filtered / par_scan / par_scan_filtered -- src/layer.rs:449-520: sorts if
 * needed, then returns the sorted, duplicate-free vector of (later_id, earlier_id) pairs, each pair
 * stored as two consecutive IDs of id_bytes.  filter == NULL means BP_FILTER_NONE.
 * bp_layer_scan copies the pairs to pinned host memory; _device leaves them on the device. */
int bp_layer_scan(bp_layer *layer, const bp_filter *filter, const void **out_pairs, size_t *out_count);
int bp_layer_scan_device(bp_layer *layer, const bp_filter *filter, const void **out_d_pairs, size_t *out_count);

/* Layer::test_box -- src/layer.rs:278-299 (BoxTestGeometry, src/geom.rs:353-460) and Layer::test_ray --
 * src/layer.rs:313-351 (RayTestGeometry, src/geom.rs:462-615), both through Layer::test / test_impl
 * (src/layer.rs:167-277), for a BATCH of test geometries in one call.  Sorts if needed, like the reference.
 *   boxes: n_queries x 2*D floats (test_bounds min.., max..)
 *   rays:  n_queries x (2*D + 2) floats (origin.., direction.., range_min, range_max; the ranges may be
 *          -inf / +inf and are clamped to the system bounds like RayTestGeometry::with_system_bounds)
 *   max_depth: Option<u32> of the reference; < 0 = None
 *   on_device: 0 = `boxes` / `rays` are host pointers and the results are copied to pinned host memory,
 *              1 = device pointers in, device pointers out
 * Result: out_pairs = out_count x {query number, object ID} (two consecutive IDs of id_bytes each), sorted
 * by (query, ID) without duplicates; out_offsets[q] .. out_offsets[q + 1] delimit query q (n_queries + 1
 * entries) -- slice q is exactly the `&Vec<ID>` the reference returns for that geometry. */
int bp_layer_test_box_batch(bp_layer *layer, const float *system_bounds, const float *boxes, size_t n_queries, int32_t max_depth,
                            int on_device, const void **out_pairs, const uint32_t **out_offsets, size_t *out_count);
int bp_layer_test_ray_batch(bp_layer *layer, const float *system_bounds, const float *rays, size_t n_queries, int32_t max_depth,
                            int on_device, const void **out_pairs, const uint32_t **out_offsets, size_t *out_count);

/* Layer::pick_ray -- src/layer.rs:424-446 (through Layer::pick / test_impl, :167-242, 364-408), for a batch of rays:
 * the nearest object each ray hits within max_dist.  The reference asks a user closure `get_dist(origin, direction,
 * nearest, id)` for the hit distance of a candidate; a closure cannot cross the ABI, so an enumerated device functor
 * over a table of shapes indexed by ID stands in (like bp_filter for scan_filtered):
 *   BP_PICK_SPHERE  shapes = n_shapes x (D + 1) floats (centre.., radius): the ray / sphere distance of the reference's
 *                   example (examples/main.rs:427-449); 0 when the origin is inside
 *   BP_PICK_AABB    shapes = n_shapes x 2*D floats (min.., max..): slab test; 0 when the origin is inside
 * IDs >= n_shapes are never hit.  The walk is the reference's: children in RayTestGeometry::test_order
 * (src/geom.rs:579-610), cells whose part of the ray starts at or behind the nearest hit so far are skipped, and
 * between equal distances the ID met first wins.  rays: n_queries x 2*D floats (origin.., direction..);
 * max_depth < 0 = None; on_device as for the test_* calls (rays, shapes and results).  One bp_pick_result per ray. */
typedef enum bp_pick_kind { BP_PICK_SPHERE = 0, BP_PICK_AABB = 1 } bp_pick_kind;
typedef struct bp_pick_result {
    float dist;     /* Some((dist, id, point)): distance of the nearest hit ... */
    uint32_t hit;   /* 0 = None */
    uint64_t id;    /* ... its ID ... */
    float point[3]; /* ... and origin + direction * dist (point[2] = 0 for the 2-D index types) */
    uint32_t pad;
} bp_pick_result;
int bp_layer_pick_ray_batch(bp_layer *layer, const float *system_bounds, const float *rays, size_t n_queries, float max_dist,
                            int32_t max_depth, int32_t shape_kind, const float *shapes, size_t n_shapes, int on_device,
                            const bp_pick_result **out_results);

/* Layer::iter -- src/layer.rs:79-81, and the state PartialEq compares (src/layer.rs:582-585):
 * keys (4 or 8 bytes each) and ids as separate arrays + the sorted flag. */
int bp_layer_records(bp_layer *layer, const void **out_keys, const void **out_ids, size_t *out_n, int *out_sorted);
int bp_layer_records_device(bp_layer *layer, const void **out_d_keys, const void **out_d_ids, size_t *out_n, int *out_sorted);
/* Replaces the tree (the serde Deserialize path of Layer, src/layer.rs:41; also how a
 * multi-GPU exchange hands a shard its records).  Host or device pointers per `on_device`. */
int bp_layer_set_records(bp_layer *layer, const void *keys, const void *ids, size_t n, int sorted, int on_device);
/* Dedup at the source across an exchange (multi-GPU).  encode_kernel writes 3 cell flags for every record of a freshly
 * extended tree; bp_dist_scatter_records_flagged ships them inside the top 3 bits of the IDs (if the IDs leave those
 * bits free).  bp_layer_set_records_flagged(.., flagged = 1) / bp_layer_sort_from_device(.., flagged = 1) load such
 * records: the scan of that layer emits every ID pair from its canonical shared cell only, and every accessor strips
 * the flags before IDs are shown. */
int bp_layer_set_records_flagged(bp_layer *layer, const void *keys, const void *ids, size_t n, int sorted, int on_device, int flagged);

/* bp_layer_set_records_flagged(.., sorted = 0, on_device = 1) followed by bp_layer_sort, as one step and without the two
 * passes over the records that combination spends before the sort starts: the first radix pass reads d_keys / d_ids where
 * they are (e.g. a multi-GPU receive buffer; not modified, no staging copy), and the digit plan comes from the caller's
 * masks instead of a mask pass -- key_or / id_or may have extra bits set and key_and / id_and extra bits clear (a plan
 * with idle digits, never a wrong order); ids_ascending must only be non-zero if the IDs do not decrease in record order
 * (then equal keys keep their order and no ID digits are sorted).  The masks describe the IDs without cell flags. */
int bp_layer_sort_from_device(bp_layer *layer, const void *d_keys, const void *d_ids, size_t n, int flagged, uint64_t key_or,
                              uint64_t key_and, uint64_t id_or, uint64_t id_and, int ids_ascending);
/* What a sender knows about the order of the IDs of a tree built by extend calls since the last clear: the IDs of the
 * first and the last object passed in (valid or not) and whether no ID was smaller than its predecessor.  With
 * *out_ascending set, [first, last] bounds the tree's IDs; several such trees concatenated in an order in which these
 * ranges do not overlap are ascending again (what bp_layer_sort_from_device wants to hear).  An empty tree reports
 * first = 2^64 - 1, last = 0, ascending = 1. */
int bp_layer_id_order(bp_layer *layer, uint64_t *out_first, uint64_t *out_last, int *out_ascending);

int bp_layer_len(bp_layer *layer, size_t *out_n);
int bp_layer_is_sorted(bp_layer *layer, int *out_sorted);
int bp_layer_min_depth(const bp_layer *layer, uint32_t *out_min_depth);
/* OR and AND over all keys / IDs of the tree (what the sort planner uses; a superset after clear-less reuse). */
int bp_layer_masks(bp_layer *layer, uint64_t *key_or, uint64_t *key_and, uint64_t *id_or, uint64_t *id_and);

/* ---- multi-GPU building blocks ------------------------------------------------------------------
 * The reference is single-process (no counterpart in the crate); these are the device operations the
 * Morton-prefix range sharding of SURVEY.md section 8e is assembled from (broadphase-rs_b200/dist.py):
 * sample sort of the records, ancestor halos, shard-local scan, range partition of the raw pairs for
 * the global dedup.  All pointers prefixed d_ are device pointers on the layer's device; the layer is
 * used as the execution context (stream + scratch).  32-bit IDs only for the pair functions. */

/* Stable partition of n records by `n_splitters` (<= 15) ascending key splitters: bucket b receives
 * the records with exactly b splitters <= key.  out_counts[0..n_splitters] are the bucket sizes. */
int bp_dist_partition_records(bp_layer *ctx, const void *d_keys, const void *d_ids, size_t n, const uint64_t *splitters,
                              int n_splitters, void *d_out_keys, void *d_out_ids, uint64_t *out_counts);
/* The same for packed raw pairs ((later << 32) | earlier), partitioned on `later`. */
int bp_dist_partition_pairs(bp_layer *ctx, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                            void *d_out_pairs, uint64_t *out_counts);
/* The fused form used over NVLink: bucket sizes first (out_counts), then a partition pass that writes
 * every bucket straight to its own destination array -- dst_keys[b] / dst_ids[b] are device ADDRESSES,
 * typically inside the receive buffers of peer GPUs (symmetric memory), so the pass is the all-to-all.
 * Halos: a record whose cell reaches past later splitters is an ancestor of records those shards will
 * own; out_halo_counts[b] counts the copies bucket b receives and bp_dist_scatter_records writes them
 * (unordered) to halo_dst_keys[b] / halo_dst_ids[b] (NULL: no halo copies). */
int bp_dist_count_records(bp_layer *ctx, const void *d_keys, size_t n, const uint64_t *splitters, int n_splitters,
                          uint64_t *out_counts, uint64_t *out_halo_counts);
int bp_dist_scatter_records(bp_layer *ctx, const void *d_keys, const void *d_ids, size_t n, const uint64_t *splitters,
                            int n_splitters, const uint64_t *dst_keys, const uint64_t *dst_ids, const uint64_t *halo_dst_keys,
                            const uint64_t *halo_dst_ids);
/* bp_dist_scatter_records with the layer's cell flags (see bp_layer_set_records_flagged) OR-ed into the top 3 bits of every
 * ID as the pass loads it -- no separate folding pass, the layer's own records stay unflagged.  d_keys / d_ids must be the
 * layer's own record arrays (bp_layer_records_device) of a freshly extended tree whose IDs leave those bits free. */
int bp_dist_scatter_records_flagged(bp_layer *ctx, const void *d_keys, const void *d_ids, size_t n, const uint64_t *splitters,
                                    int n_splitters, const uint64_t *dst_keys, const uint64_t *dst_ids,
                                    const uint64_t *halo_dst_keys, const uint64_t *halo_dst_ids, int fold_cell_flags);
int bp_dist_count_pairs(bp_layer *ctx, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                        uint64_t *out_counts);
int bp_dist_scatter_pairs(bp_layer *ctx, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                          const uint64_t *dst_pairs);
/* The counts of bp_dist_count_records / _pairs left ON THE DEVICE as one row of 64-bit words -- [bucket sizes
 * 0..n_splitters] [halo copies 0..n_splitters] (records only) [tags] -- typically this rank's row of a count matrix in
 * symmetric memory: the matrix is then complete after one barrier instead of a host round trip plus an NCCL
 * all-gather.  Asynchronous on the layer's stream.  Up to 8 tag words close the row (the sort masks and ID order of the
 * sender travel with its counts, so that the receivers can plan their sort without looking at the records:
 * bp_layer_sort_from_device), and the row is stored to up to 16 device ADDRESSES at once -- this rank's row inside
 * every rank's copy of the matrix, the peers' through NVLink: the kernel that finishes the counts also distributes
 * them, no copy-engine transfers in between. */
int bp_dist_count_records_rows(bp_layer *ctx, const void *d_keys, size_t n, const uint64_t *splitters, int n_splitters,
                               const uint64_t *tags, int n_tags, const uint64_t *d_out_rows, int n_out_rows);
/* bp_layer_clear + bp_layer_extend_device + bp_dist_count_records_rows in one asynchronous step, for frames whose
 * splitters are known before the objects are encoded (cached from the previous frame): encode_kernel counts every
 * record it generates per destination shard (and per halo copy), so no counting pass over the keys runs, and the row --
 * [counts | halo counts | 7 tag words: id_or (| 1 << 63 if allow_fold and the IDs leave their top 3 bits free), key_or,
 * key_and, id_and, first ID, last ID, IDs ascending (bp_layer_masks / bp_layer_id_order)] -- is assembled on the device
 * and stored to every address in d_out_rows.  Nothing returns to the host: the caller's first synchronisation is the one
 * that fetches the finished count matrix.  The layer must have min_depth 0. */
int bp_dist_extend_count_rows(bp_layer *layer, const float *system_bounds, const float *d_bounds, const void *d_ids, size_t n,
                              const uint64_t *splitters, int n_splitters, int allow_fold, const uint64_t *d_out_rows,
                              int n_out_rows);
int bp_dist_count_pairs_rows(bp_layer *ctx, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                             const uint64_t *tags, int n_tags, const uint64_t *d_out_rows, int n_out_rows);
/* Equal range [lo, hi) of every query key in a sorted device key array (halo look-ups). */
int bp_dist_lookup_ranges(bp_layer *ctx, const void *d_sorted_keys, size_t n, const uint64_t *queries, int n_queries,
                          uint64_t *out_lo, uint64_t *out_hi);
/* Records [0, n_halo) of the tree are halo: they are ancestors only, never the later record of an
 * emitted pair.  Reset to 0 by bp_layer_clear / bp_layer_set_records. */
int bp_layer_set_halo(bp_layer *layer, size_t n_halo);
/* Dedup at the source (the scan emits an ID pair from the canonical one of the cells two objects share, DESIGN.md
 * section 4) is only valid while NO record of the whole scene is inactive (an ID owning nested bounds,
 * src/layer.rs:562-564).  A single layer decides that by itself; shards of a distributed scene must decide it
 * together: a shard that saw such a record (bp_stats.rescans != 0 after the scan) tells the others, and every shard
 * whose scan ran with the dedup scans again with it switched off (enabled = 0).  Default: enabled. */
int bp_layer_set_scan_dedup(bp_layer *layer, int enabled);
/* scan up to the raw pairs (filtered, not yet sorted / deduplicated), packed (later << 32) | earlier. */
int bp_layer_scan_raw_device(bp_layer *layer, const bp_filter *filter, const void **out_d_raw, size_t *out_count);
/* Sorts + deduplicates n packed raw pairs (possibly received from other shards) into the final
 * (later, earlier) layout.  id_mask: OR of all bits in which two IDs may differ (0: use the layer's own). */
int bp_layer_unique_pairs_device(bp_layer *layer, const void *d_raw, size_t n, uint64_t id_mask, const void **out_d_pairs,
                                 size_t *out_count);
/* What the caller knows about the pairs of the NEXT bp_layer_unique_pairs_*_device call: the bits of fixed_bits (bits 0-31)
 * have the same value in the later ID of every pair -- a shard's slice of a range partition on the later ID lies between
 * two splitters -- so the pair sort spends no radix pass on them.  Forgotten after that call; 0 = nothing known. */
int bp_layer_set_pair_later_fixed(bp_layer *layer, uint64_t fixed_bits);
/* The same without the staging copy: d_raw itself is one side of the sort's ping-pong (its contents are destroyed). */
int bp_layer_unique_pairs_inplace_device(bp_layer *layer, void *d_raw, size_t n, uint64_t id_mask, const void **out_d_pairs,
                                         size_t *out_count);

/* ---- the sharded frame as one call ---------------------------------------------------------------
 * A bp_dist context runs the whole multi-GPU frame -- clear -> extend -> par_sort -> [merge] -> par_scan(_filtered) of
 * one scene spread over `world` GPUs of one NVLink domain, one process (or thread) per GPU -- in C++ on top of the
 * building blocks above (csrc/bp_dist.cu).  No reference counterpart: the crate is single-process;
 * the calls mirror Layer::clear / extend / merge / scan_filtered (src/layer.rs:84-138, 449-520) for the distributed scene.
 *
 * Peer memory: every rank owns ONE device arena (barrier flags, sample and count matrices, receive buffers of
 * record_capacity records and pair_capacity raw pairs).  bp_dist_export writes an opaque blob of bp_dist_handle_bytes()
 * bytes (a CUDA IPC handle); the caller all-gathers the blobs with whatever transport it has (MPI, torch.distributed, a
 * file) and passes all `world` of them, in rank order, to bp_dist_connect.  world == 1 needs neither call.
 * Every rank must call bp_dist_set_static / bp_dist_frame the same number of times in the same order (they contain
 * device-side barriers).  BP_ERR_TOO_LARGE: a receive buffer is too small -- the same on every rank, because the count
 * matrices are global; bp_dist_last_info tells the sizes needed, recreate the contexts with larger capacities. */
typedef struct bp_dist bp_dist;
typedef struct bp_dist_config {
    int32_t index_kind;     /* BP_INDEX64_2D or BP_INDEX64_3D; IDs are u32 */
    uint32_t min_depth;     /* LayerBuilder::with_min_depth, src/layer.rs:646-649 */
    int32_t device;         /* CUDA device ordinal; -1 = current device */
    int32_t rank, world;    /* world <= 16 */
    size_t record_capacity; /* records one rank can RECEIVE in a frame (owned + halo copies) */
    size_t pair_capacity;   /* raw pairs one rank can receive in a frame */
} bp_dist_config;
enum { BP_DIST_PHASES = 9 }; /* encode, splitters, counts, exchange, sort, scan, pair_counts, pair_exchange, unique */
typedef struct bp_dist_info { /* the last frame on this rank */
    uint64_t records_local;  /* records encoded from this rank's objects */
    uint64_t records_owned;  /* records of this rank's Morton range after the exchange */
    uint64_t n_halo;         /* ancestor (halo) records in front of them, static halo included */
    uint64_t raw_pairs, pairs;
    uint64_t records_needed, pairs_needed; /* fullest receive buffer of the frame (what the capacities must hold) */
    int32_t fused;           /* counts taken by the encode kernel (cached splitters) */
    int32_t rescanned;       /* the scan ran again without dedup at the source (another shard saw an inactive record) */
    int32_t rebalance_records, rebalance_pairs; /* cached splitters dropped after this frame */
    double phase_ms[BP_DIST_PHASES]; /* with BP_DIST_OPT_TRACE: device-synchronised wall time of every phase */
} bp_dist_info;
enum { BP_DIST_OPT_REUSE_SPLITTERS = 0, BP_DIST_OPT_FUSE_COUNTS = 1, BP_DIST_OPT_GLOBAL_DEDUP_DECISION = 2, BP_DIST_OPT_TRACE = 3 };

int bp_dist_create(const bp_dist_config *config, bp_dist **out);
int bp_dist_destroy(bp_dist *ctx);
size_t bp_dist_handle_bytes(void);
int bp_dist_export(bp_dist *ctx, void *out_blob);
int bp_dist_connect(bp_dist *ctx, const void *all_blobs);
int bp_dist_set_stream(bp_dist *ctx, void *cuda_stream);
int bp_dist_set_option(bp_dist *ctx, int option, int value);
/* Shards a static scene once: sorted, kept resident, merged into every frame (the reference's "static scene layer"
 * use of Layer::merge, README + src/layer.rs:127-138).  Its splitters stay fixed from then on. */
int bp_dist_set_static(bp_dist *ctx, const float *system_bounds, const float *d_bounds, const void *d_ids, size_t n);
/* One frame on this rank's n objects (device pointers).  *out_d_pairs: this rank's slice of the globally sorted,
 * duplicate-free (later, earlier) pair list (u32 IDs), on the device, valid until the next call; the slices of ranks
 * 0 .. world-1 concatenated are exactly the reference's scan() vector of the whole scene. */
int bp_dist_frame(bp_dist *ctx, const float *system_bounds, const float *d_bounds, const void *d_ids, size_t n,
                  const bp_filter *filter, const void **out_d_pairs, size_t *out_count);
int bp_dist_last_info(const bp_dist *ctx, bp_dist_info *out);
/* The context's layers, for bp_layer_stats / bp_layer_set_profiling: 0 = encode, 1 = shard, 2 = static. */
bp_layer *bp_dist_layer(bp_dist *ctx, int which);
const char *bp_dist_last_error(const bp_dist *ctx);

/* ---- instrumentation -------------------------------------------------------------------------- */
int bp_layer_set_profiling(bp_layer *layer, int enabled); /* CUDA-event timing of every kernel launch */
int bp_layer_reset_stats(bp_layer *layer);
int bp_layer_stats(bp_layer *layer, bp_stats *out);
const char *bp_layer_last_error(const bp_layer *layer);
const char *bp_status_string(int status);
int bp_version(void);
int bp_device_count(int *out_count);

/* Host-side radix planning, exposed for tests: given the mask of key bits that differ between
 * any two records, writes up to 16 digit descriptors and returns how many.  A digit is one or two
 * bit-fields: the low half-words of out_shift[i] / out_bits[i] describe the first field, the high
 * half-words the second (bits == 0: absent), stacked above the first. */
int bp_plan_radix_passes(uint64_t varying_mask, uint32_t *out_shift, uint32_t *out_bits, int max_passes);

/* The "top bits + finish" plan of a record sort, exposed for tests: n_records records whose keys differ in the bits
 * of varying_mask.  Returns 1 and writes the bits the radix passes sort on (*out_top_mask) and the shift that defines a
 * finish group (records with equal key >> *out_group_shift) when the plan replaces at least two radix passes, else 0:
 * R records only carry ~log2(R) bits of position, so after a stable sort on the top log2(R) - 3 varying bits (rounded up
 * to whole 8-bit passes) every group of records agreeing on them is a handful of neighbours, ordered by one more pass
 * (record_finish_kernel) however many low bits remain.  Replaces the tail of `tree.par_sort_unstable()`,
 * src/layer.rs:149, :162 -- a total order, so the result is the same sequence. */
int bp_plan_sort_finish(uint64_t varying_mask, uint64_t n_records, uint64_t *out_top_mask, uint32_t *out_group_shift);

/* Host-side planning of the sharded frame, exposed for tests (no device involved).  bp_dist_plan_splitters: parts - 1
 * ascending splitters from a gathered key sample (the array is sorted in place) -- the sample quantiles, each moved to the
 * roundest value (most trailing zero bits) whose sample rank stays within 1/32 of a shard's size.  A key v belongs to shard
 * d iff splitters[d-1] <= v < splitters[d], so every key of a shard carries the bits its two ends share:
 * bp_dist_plan_shard_bits writes their positions (*out_fixed) and values (*out_value) for shard `shard`, `top` being the
 * largest key that can occur (the last shard's upper end) -- the bits the shard's record sort, and the pair sort of its
 * slice of the later IDs, never look at. */
int bp_dist_plan_splitters(uint64_t *sample, size_t n, int parts, uint64_t *out_splitters);
int bp_dist_plan_shard_bits(const uint64_t *splitters, int parts, int shard, uint64_t top, uint64_t *out_fixed, uint64_t *out_value);

#ifdef __cplusplus
}
#endif
#endif /* BP_H */
