/*
 * bp_oracle.h -- C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the broadphase-rs hot path
 * (extend / sort / par_sort / merge / scan / par_scan / scan_filtered).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product (libbroadphase_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY STATUS: PINNED by reference-produced data.  The codec and quantiser are pinned by the reference's
 * in-source known-answer tests (src/index.rs:343-374, src/geom.rs:696-706).  extend / sort / scan end to end
 * are pinned by the reference's own golden files: tests/data/ holds Git-LFS pointers (SHA-256 + size) of the
 * seven gen_boxes input scenes and of the three validation files tests/test_layer.rs:25-124 compares against;
 * the inputs are regenerated bit for bit (rand_core 0.5 seed_from_u64 + ChaCha20 + rand 0.7 gen_range restated in
 * broadphase-rs_b200/rust_rand.py, all seven hashes equal), and this oracle's extend, sort and scan of the
 * n = 10 000 scene serialise to files with exactly the three validation hashes (tests/test_reference_fixtures.py).
 * Still unpinned: the queries (test_box / test_ray / pick_ray) -- the reference holds no vector for them.
 */
#ifndef BP_ORACLE_H
#define BP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* index kinds -- src/index.rs:293-295 */
enum { BPO_INDEX32_2D = 0, BPO_INDEX64_2D = 1, BPO_INDEX64_3D = 2 };

/* filters for scan_filtered (src/layer.rs:456-477); (a, b) = (later id, earlier id) */
enum {
    BPO_FILTER_NONE = 0,      /* |_, _| true                                  */
    BPO_FILTER_ID_PARITY = 1, /* ((a ^ b) & 1) == 1                           */
    BPO_FILTER_XOR_MASK = 2,  /* ((a ^ b) & arg) != 0                         */
    BPO_FILTER_CATEGORY = 3,  /* (cat[a] & msk[b]) != 0 && (cat[b] & msk[a]) != 0,
                                 table = n_table x {u32 cat, u32 msk}; ids >= n_table act as all-ones */
    BPO_FILTER_SPHERES = 4    /* !(|c[b] - c[a]| > r[a] + r[b]), table = n_table x {f32 x, y, z, r} (16 bytes per ID, passed
                                 through the u32 pointer); ids >= n_table pass */
};

typedef struct bpo_layer bpo_layer;

bpo_layer *bpo_layer_new(int kind, int id_bytes, uint32_t min_depth);
void bpo_layer_free(bpo_layer *);
void bpo_layer_clear(bpo_layer *);
/* sys_bounds: 2*D floats (min.., max..); bounds: n x 2*D floats; ids: n x id_bytes */
void bpo_layer_extend(bpo_layer *, const float *sys_bounds, const float *bounds, const void *ids, size_t n);
void bpo_layer_merge(bpo_layer *, const bpo_layer *other);
void bpo_layer_sort(bpo_layer *);
void bpo_layer_par_sort(bpo_layer *);
size_t bpo_layer_scan(bpo_layer *, int filter, uint64_t arg, const uint32_t *table, size_t n_table);
size_t bpo_layer_par_scan(bpo_layer *, int filter, uint64_t arg, const uint32_t *table, size_t n_table);

size_t bpo_layer_len(const bpo_layer *);
int bpo_layer_sorted(const bpo_layer *);
uint32_t bpo_layer_min_depth(const bpo_layer *);
size_t bpo_layer_num_collisions(const bpo_layer *);
size_t bpo_layer_num_raw_collisions(const bpo_layer *); /* pairs before sort+dedup in the last scan */
/* copies widened to u64 */
void bpo_layer_records(const bpo_layer *, uint64_t *keys, uint64_t *ids);
void bpo_layer_collisions(const bpo_layer *, uint64_t *a, uint64_t *b);
/* loads records directly (keys/ids widened to u64) -- used to test sort/scan on arbitrary trees */
void bpo_layer_set_records(bpo_layer *, const uint64_t *keys, const uint64_t *ids, size_t n, int sorted);

/* Layer::test_box / test_ray (src/layer.rs:244-351) for one geometry; max_depth < 0 = None.  box: 2*D floats
 * (min.., max..); ray: 2*D + 2 floats (origin.., direction.., range_min, range_max).  Returns the number of
 * IDs; bpo_layer_test_results copies them (sorted, unique, widened to u64).  PARITY UNPINNED: the reference has
 * no test for its queries, and the cell centres come from cgmath's midpoint (restated, see bp_oracle.cpp). */
size_t bpo_layer_test_box(bpo_layer *, const float *sys_bounds, const float *box, int max_depth);
size_t bpo_layer_test_ray(bpo_layer *, const float *sys_bounds, const float *ray, int max_depth);
void bpo_layer_test_results(const bpo_layer *, uint64_t *ids);
/* Layer::pick_ray (src/layer.rs:424-446) for one ray (2*D floats: origin.., direction..) with an enumerated shape functor
 * in place of the get_dist closure: shapes = n_shapes x (D + 1) floats (centre.., radius) for BPO_PICK_SPHERE, n_shapes x
 * 2*D floats (min.., max..) for BPO_PICK_AABB, indexed by ID.  Returns 1 on a hit: out = {dist, point[D]}, *out_id. */
enum { BPO_PICK_SPHERE = 0, BPO_PICK_AABB = 1 };
int bpo_layer_pick_ray(bpo_layer *, const float *sys_bounds, const float *ray, float max_dist, int max_depth, int shape_kind,
                       const float *shapes, size_t n_shapes, float *out, uint64_t *out_id);

/* codec + quantiser, exposed for the known-answer tests */
uint64_t bpo_encode_axis(int kind, uint32_t v);
uint32_t bpo_decode_axis(int kind, uint64_t origin_bits);
uint64_t bpo_make_index(int kind, uint32_t depth, const uint32_t *origin);
uint64_t bpo_level_mask(int kind, uint32_t depth);
int bpo_overlaps(int kind, uint64_t a, uint64_t b);
void bpo_to_local(int dim, const float *sys_bounds, const float *bounds, uint32_t *out_local);
void bpo_to_global(int dim, const float *sys_bounds, const uint32_t *local, float *out_bounds);

int bpo_max_threads(void);
void bpo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
