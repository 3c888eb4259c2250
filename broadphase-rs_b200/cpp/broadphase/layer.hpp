// broadphase/layer.hpp -- C++ host-side mirror of the reference crate's public interface for the hot
// path, over the C ABI of include/bp.h.  (The reference is Rust; no Rust toolchain exists in this
// image, so the host side above the C ABI is written in C++ with the crate's names, argument
// meaning and implicit behaviour.  The equivalent Rust shim is in INTEGRATION.md.)
//
//   reference (Rust)                                       here (C++)
//   ------------------------------------------------------------------------------------------------
//   broadphase::Index32_2D / Index64_2D / Index64_3D       broadphase::Index32_2D / Index64_2D / Index64_3D
//   broadphase::Bounds<Point3<f32>>{min, max}              broadphase::Bounds<3>{min, max}
//   LayerBuilder::new().with_min_depth(4).build()          LayerBuilder().with_min_depth(4).build<Index, ID>()
//   layer.clear()                          src/layer.rs:84  layer.clear()
//   layer.extend(system_bounds, iter)      src/layer.rs:94  layer.extend(system_bounds, first, last) / (bounds*, ids*, n)
//   layer.merge(&other)                    src/layer.rs:127 layer.merge(other)
//   layer.sort() / par_sort()              src/layer.rs:146 layer.sort() / par_sort()
//   layer.scan() / par_scan()              src/layer.rs:449 layer.scan() / par_scan()   -> const std::vector-like view
//   layer.scan_filtered(f) / par_..        src/layer.rs:456 layer.scan_filtered(Filter) / par_scan_filtered(Filter)
//   layer.iter()                           src/layer.rs:79  layer.iter()
//   layer.clone() / layer == other     src/layer.rs:576-617 layer.clone() / layer == other (min_depth, records, sorted flag)
//   layer.test_box(system_bounds, b, d)    src/layer.rs:293 layer.test_box(system_bounds, b, max_depth) -> std::vector<ID>
//   layer.test_ray(system_bounds, o, v, ..) src/layer.rs:326 layer.test_ray(system_bounds, origin, direction, rmin, rmax, max_depth)
//   (many geometries in one call)                           layer.test_box_batch(...) / test_ray_batch(...) -> QueryResults<ID>
//   layer.pick_ray(system_bounds, o, v, max_dist, d, f) src/layer.rs:424 layer.pick_ray(system_bounds, origin, direction, max_dist, Shapes, max_depth)
//                                                           (get_dist closure -> enumerated shape functor over a table indexed by ID)
//
// Errors: the reference never returns errors from these methods; allocation failure aborts.  Here a
// failed C-ABI call throws broadphase::Error (status + message).
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/bp.h"

namespace broadphase {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(std::string(bp_status_string(s)) + ": " + m), status(s) {}
};

// index tags -- src/index.rs:293-295
struct Index32_2D { static constexpr int KIND = BP_INDEX32_2D, DIM = 2; typedef uint32_t key_type; };
struct Index64_2D { static constexpr int KIND = BP_INDEX64_2D, DIM = 2; typedef uint64_t key_type; };
struct Index64_3D { static constexpr int KIND = BP_INDEX64_3D, DIM = 3; typedef uint64_t key_type; };

// Bounds -- src/geom.rs:84-87: min and max are inclusive; memory layout = 2*DIM floats (min.., max..)
template <int DIM> struct Bounds {
    float min[DIM];
    float max[DIM];
};

// device functors for scan_filtered -- include/bp.h bp_filter_kind
struct Filter {
    bp_filter f{BP_FILTER_NONE, 0, 0, nullptr, 0};
    static Filter none() { return Filter(); }
    static Filter id_parity() { Filter r; r.f.kind = BP_FILTER_ID_PARITY; return r; }
    static Filter xor_mask(uint64_t m) { Filter r; r.f.kind = BP_FILTER_XOR_MASK; r.f.arg = m; return r; }
    static Filter category(const uint32_t *table, size_t n, bool on_device = false) {
        Filter r; r.f.kind = BP_FILTER_CATEGORY; r.f.table = table; r.f.n_table = n; r.f.table_on_device = on_device; return r;
    }
    // fused narrow phase: table = n x {x, y, z, r} floats indexed by ID; only pairs whose spheres touch pass
    static Filter spheres(const float *table, size_t n, bool on_device = false) {
        Filter r; r.f.kind = BP_FILTER_SPHERES; r.f.table = reinterpret_cast<const uint32_t *>(table); r.f.n_table = n;
        r.f.table_on_device = on_device; return r;
    }
};

// a borrowed view of the layer's result buffer (the reference returns &Vec<(ID, ID)>)
template <class ID> struct PairView {
    const std::pair<ID, ID> *data_ = nullptr;
    size_t size_ = 0;
    const std::pair<ID, ID> *begin() const { return data_; }
    const std::pair<ID, ID> *end() const { return data_ + size_; }
    size_t size() const { return size_; }
    bool empty() const { return size_ == 0; }
    const std::pair<ID, ID> &operator[](size_t i) const { return data_[i]; }
};

// results of a batched query: ids[offsets[q] .. offsets[q + 1]) is the sorted, duplicate-free ID list of geometry q
template <class ID> struct QueryResults {
    std::vector<uint32_t> offsets;
    std::vector<ID> ids;
};

template <class Index, class ID> class Layer {
    static_assert(sizeof(ID) == 4 || sizeof(ID) == 8, "ObjectID must be a 32- or 64-bit integer (src/traits.rs:6-16)");
    static_assert(sizeof(std::pair<ID, ID>) == 2 * sizeof(ID), "pair layout");
    bp_layer *h_ = nullptr;
    bp_layer_config cfg_{}; // what the layer was built with (clone() builds its copy the same way)
    void ck(int st) const {
        if (st != BP_OK) throw Error(st, h_ ? bp_layer_last_error(h_) : "");
    }
    friend class LayerBuilder;
    explicit Layer(const bp_layer_config &cfg) : cfg_(cfg) {
        int st = bp_layer_create(&cfg, &h_);
        if (st != BP_OK) throw Error(st, "bp_layer_create");
    }

public:
    typedef typename Index::key_type key_type;
    Layer(Layer &&o) noexcept : h_(o.h_), cfg_(o.cfg_) { o.h_ = nullptr; }
    Layer &operator=(Layer &&o) noexcept { std::swap(h_, o.h_); std::swap(cfg_, o.cfg_); return *this; }
    Layer(const Layer &) = delete;
    Layer &operator=(const Layer &) = delete;
    ~Layer() { if (h_) bp_layer_destroy(h_); }

    bp_layer *handle() const { return h_; }

    void clear() { ck(bp_layer_clear(h_)); }                                             // src/layer.rs:84-88

    // extend from parallel host arrays
    void extend(const Bounds<Index::DIM> &system_bounds, const Bounds<Index::DIM> *bounds, const ID *ids, size_t n) {
        ck(bp_layer_extend_host(h_, system_bounds.min, bounds ? bounds[0].min : nullptr, ids, n)); // src/layer.rs:94-121
    }
    // extend from an iterator of (Bounds, ID), like the reference's `objects: Iter`
    template <class It> void extend(const Bounds<Index::DIM> &system_bounds, It first, It last) {
        std::vector<Bounds<Index::DIM>> b;
        std::vector<ID> ids;
        for (; first != last; ++first) {
            b.push_back(first->first);
            ids.push_back(first->second);
        }
        extend(system_bounds, b.data(), ids.data(), b.size());
    }
    // extend from device-resident arrays (no host round trip)
    void extend_device(const Bounds<Index::DIM> &system_bounds, const float *d_bounds, const ID *d_ids, size_t n) {
        ck(bp_layer_extend_device(h_, system_bounds.min, d_bounds, d_ids, n));
    }

    void merge(const Layer &other) { ck(bp_layer_merge(h_, other.h_)); }                 // src/layer.rs:127-138
    void sort() { ck(bp_layer_sort(h_)); }                                               // src/layer.rs:157-165
    void par_sort() { ck(bp_layer_sort(h_)); }                                           // src/layer.rs:146-152

    PairView<ID> scan_filtered(const Filter &filter) {                                   // src/layer.rs:456-477
        const void *p = nullptr;
        size_t n = 0;
        ck(bp_layer_scan(h_, &filter.f, &p, &n));
        return PairView<ID>{static_cast<const std::pair<ID, ID> *>(p), n};
    }
    PairView<ID> scan() { return scan_filtered(Filter::none()); }                        // src/layer.rs:449-453
    PairView<ID> par_scan() { return scan_filtered(Filter::none()); }                    // src/layer.rs:482-487
    PairView<ID> par_scan_filtered(const Filter &f) { return scan_filtered(f); }         // src/layer.rs:489-520

    // Layer::test_box / test_ray -- src/layer.rs:293-351, batched (max_depth < 0 = None).  boxes: n x Bounds;
    // rays: n x (origin[DIM], direction[DIM], range_min, range_max) floats.
    QueryResults<ID> test_box_batch(const Bounds<Index::DIM> &system_bounds, const Bounds<Index::DIM> *boxes, size_t n, int max_depth = -1) {
        return collect(0, system_bounds, boxes ? boxes[0].min : nullptr, n, max_depth);
    }
    QueryResults<ID> test_ray_batch(const Bounds<Index::DIM> &system_bounds, const float *rays, size_t n, int max_depth = -1) {
        return collect(1, system_bounds, rays, n, max_depth);
    }
    std::vector<ID> test_box(const Bounds<Index::DIM> &system_bounds, const Bounds<Index::DIM> &test_bounds, int max_depth = -1) {
        return test_box_batch(system_bounds, &test_bounds, 1, max_depth).ids;
    }
    std::vector<ID> test_ray(const Bounds<Index::DIM> &system_bounds, const float (&origin)[Index::DIM], const float (&direction)[Index::DIM],
                             float range_min, float range_max, int max_depth = -1) {
        float ray[2 * Index::DIM + 2];
        for (int i = 0; i < Index::DIM; ++i) {
            ray[i] = origin[i];
            ray[Index::DIM + i] = direction[i];
        }
        ray[2 * Index::DIM] = range_min;
        ray[2 * Index::DIM + 1] = range_max;
        return test_ray_batch(system_bounds, ray, 1, max_depth).ids;
    }

    // Layer::pick_ray -- src/layer.rs:424-446, batched: rays = n x (origin[DIM], direction[DIM]); shapes = table indexed by ID
    // (BP_PICK_SPHERE: centre[DIM], radius; BP_PICK_AABB: min[DIM], max[DIM]).  One bp_pick_result per ray (hit == 0: None).
    std::vector<bp_pick_result> pick_ray_batch(const Bounds<Index::DIM> &system_bounds, const float *rays, size_t n, float max_dist,
                                               bp_pick_kind kind, const float *shapes, size_t n_shapes, int max_depth = -1) {
        const bp_pick_result *res = nullptr;
        ck(bp_layer_pick_ray_batch(h_, system_bounds.min, rays, n, max_dist, max_depth, kind, shapes, n_shapes, 0, &res));
        return std::vector<bp_pick_result>(res, res + n);
    }

    // Layer::iter -- src/layer.rs:79-81
    std::vector<std::pair<key_type, ID>> iter() {
        const void *k = nullptr, *i = nullptr;
        size_t n = 0;
        int sorted = 0;
        ck(bp_layer_records(h_, &k, &i, &n, &sorted));
        std::vector<std::pair<key_type, ID>> out(n);
        for (size_t r = 0; r < n; ++r) out[r] = {static_cast<const key_type *>(k)[r], static_cast<const ID *>(i)[r]};
        return out;
    }
    // Clone -- src/layer.rs:597-617: min_depth and the tree with its sorted flag; the result buffers of the copy start
    // empty.  (Through the host mirror of bp_layer_records: cloning is not on the hot path.)
    Layer clone() {
        bp_layer_config c = cfg_;
        c.min_depth = min_depth(); // (merge may have lowered it, src/layer.rs:131-134)
        Layer out(c);
        const void *k = nullptr, *i = nullptr;
        size_t n = 0;
        int sorted = 0;
        ck(bp_layer_records(h_, &k, &i, &n, &sorted));
        out.ck(bp_layer_set_records(out.h_, k, i, n, sorted, 0));
        return out;
    }
    // PartialEq -- src/layer.rs:576-587: min_depth and `tree`, which is the (Index, ID) sequence AND its sorted flag
    bool equals(Layer &other) {
        const void *ka = nullptr, *ia = nullptr, *kb = nullptr, *ib = nullptr;
        size_t na = 0, nb = 0;
        int sa = 0, sb = 0;
        ck(bp_layer_records(h_, &ka, &ia, &na, &sa));
        other.ck(bp_layer_records(other.h_, &kb, &ib, &nb, &sb));
        if (min_depth() != other.min_depth() || na != nb || (sa != 0) != (sb != 0)) return false;
        for (size_t r = 0; r < na; ++r)
            if (static_cast<const key_type *>(ka)[r] != static_cast<const key_type *>(kb)[r] ||
                static_cast<const ID *>(ia)[r] != static_cast<const ID *>(ib)[r])
                return false;
        return true;
    }
    friend bool operator==(Layer &a, Layer &b) { return a.equals(b); }
    friend bool operator!=(Layer &a, Layer &b) { return !a.equals(b); }
    size_t len() { size_t n = 0; ck(bp_layer_len(h_, &n)); return n; }
    bool is_sorted() { int s = 0; ck(bp_layer_is_sorted(h_, &s)); return s != 0; }
    uint32_t min_depth() const { uint32_t d = 0; bp_layer_min_depth(h_, &d); return d; }
    bp_stats stats() { bp_stats s; ck(bp_layer_stats(h_, &s)); return s; }

private:
    QueryResults<ID> collect(int ray, const Bounds<Index::DIM> &system_bounds, const float *params, size_t n, int max_depth) {
        const void *pairs = nullptr;
        const uint32_t *offsets = nullptr;
        size_t count = 0;
        ck(ray ? bp_layer_test_ray_batch(h_, system_bounds.min, params, n, max_depth, 0, &pairs, &offsets, &count)
               : bp_layer_test_box_batch(h_, system_bounds.min, params, n, max_depth, 0, &pairs, &offsets, &count));
        QueryResults<ID> r;
        r.offsets.assign(offsets, offsets + n + 1);
        r.ids.resize(count);
        for (size_t i = 0; i < count; ++i) r.ids[i] = static_cast<const ID *>(pairs)[2 * i + 1]; // {query, id}
        return r;
    }
};

// LayerBuilder -- src/layer.rs:620-696
class LayerBuilder {
    bp_layer_config cfg_{};

public:
    LayerBuilder() { cfg_.device = -1; }
    LayerBuilder &with_min_depth(uint32_t depth) { cfg_.min_depth = depth; return *this; }
    LayerBuilder &with_index_capacity(size_t c) { cfg_.index_capacity = c; return *this; }
    LayerBuilder &with_collision_capacity(size_t c) { cfg_.collision_capacity = c; return *this; }
    LayerBuilder &with_test_capacity(size_t c) { cfg_.test_capacity = c; return *this; }
    LayerBuilder &with_device(int device) { cfg_.device = device; return *this; }
    template <class Index, class ID> Layer<Index, ID> build() const {
        bp_layer_config c = cfg_;
        c.index_kind = Index::KIND;
        c.id_bytes = (int32_t)sizeof(ID);
        return Layer<Index, ID>(c);
    }
};

// One scene spread over the GPUs of an NVLink domain, one DistFrame per rank (one process per GPU).  No counterpart in the
// reference crate (single process): frame() is clear -> extend -> par_sort -> [merge(static)] -> par_scan_filtered
// (src/layer.rs:84-165, 449-520) of the whole distributed scene, run by ONE C-ABI call (bp_dist_frame, csrc/bp_dist.cu).
// The host brings its own transport for the one thing that has to travel at start-up: an all-gather of every rank's blob.
template <class Index> class DistFrame {
    bp_dist *h_ = nullptr;
    void ck(int st) const {
        if (st != BP_OK) throw Error(st, bp_dist_last_error(h_));
    }

public:
    DistFrame(int rank, int world, int device, size_t record_capacity, size_t pair_capacity, uint32_t min_depth = 0) {
        bp_dist_config c{};
        c.index_kind = Index::KIND;
        c.min_depth = min_depth;
        c.device = device;
        c.rank = rank;
        c.world = world;
        c.record_capacity = record_capacity;
        c.pair_capacity = pair_capacity;
        const int st = bp_dist_create(&c, &h_);
        if (st != BP_OK) throw Error(st, "bp_dist_create");
    }
    DistFrame(const DistFrame &) = delete;
    DistFrame &operator=(const DistFrame &) = delete;
    ~DistFrame() { bp_dist_destroy(h_); }

    // this rank's blob; all_gather the blobs of ranks 0 .. world-1 (MPI_Allgather, a socket, a file) and connect()
    std::vector<unsigned char> export_blob() {
        std::vector<unsigned char> b(bp_dist_handle_bytes());
        ck(bp_dist_export(h_, b.data()));
        return b;
    }
    void connect(const std::vector<unsigned char> &all_blobs) { ck(bp_dist_connect(h_, all_blobs.data())); }
    void set_static(const Bounds<Index::DIM> &system_bounds, const float *d_bounds, const uint32_t *d_ids, size_t n) {
        ck(bp_dist_set_static(h_, system_bounds.min, d_bounds, d_ids, n));
    }
    // -> this rank's slice of the globally sorted, duplicate-free (later, earlier) pair list: device pointer + length,
    // valid until the next call (like the reference's borrowed &Vec<(ID, ID)>)
    std::pair<const uint32_t *, size_t> frame(const Bounds<Index::DIM> &system_bounds, const float *d_bounds, const uint32_t *d_ids, size_t n,
                                              const bp_filter *filter = nullptr) {
        const void *pairs = nullptr;
        size_t count = 0;
        ck(bp_dist_frame(h_, system_bounds.min, d_bounds, d_ids, n, filter, &pairs, &count));
        return {static_cast<const uint32_t *>(pairs), count};
    }
    bp_dist_info last_info() const {
        bp_dist_info i;
        bp_dist_last_info(h_, &i);
        return i;
    }
};

} // namespace broadphase
