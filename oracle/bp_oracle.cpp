// bp_oracle.cpp -- CPU oracle: a from-scratch C++ restatement of the broadphase-rs hot path.
//
// TEST INFRASTRUCTURE ONLY (see bp_oracle.h).  Never linked into or called by the product.
//
// PARITY STATUS: pinned.  Codec/quantiser by the reference's in-source KATs; extend/sort/scan end to end by
// the SHA-256 of the reference's three validation files (tests/data/validation/*, LFS pointers), which this
// restatement reproduces byte for byte from the regenerated n = 10 000 input scene
// (tests/test_reference_fixtures.py); additionally cross-checked against the independent numpy restatement
// in oracle/pyref.py.  The query functions (test_box / test_ray / pick_ray) remain unpinned.
//
// Every function cites the reference file:line (relative to /root/reference) that it follows.
// Build: g++ -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off (never -ffast-math), see Makefile.

#include "bp_oracle.h"

#include <omp.h>

#include <algorithm>
#include <cstring>
#include <parallel/algorithm>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Index codec -- src/index.rs:65-295
// ---------------------------------------------------------------------------------------------

// Spread the low 21 bits of v so that bit i lands on bit 3*i.  Equivalent (bit for bit) to the
// three-step octal-mask spread of src/index.rs:193-207; checked exhaustively against the
// bit-by-bit definition in tests/test_oracle.py.
inline uint64_t spread3(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// Spread the low 32 bits of v so that bit i lands on bit 2*i (src/index.rs:155-172).
inline uint64_t spread2(uint64_t v) {
    v &= 0xffffffffull;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

inline uint32_t compact_bits(uint64_t v, int stride, int nbits) {
    uint32_t out = 0;
    for (int i = 0; i < nbits; ++i) out |= (uint32_t)((v >> (stride * i)) & 1u) << i;
    return out;
}

// index_impl!{index: $name, $dim, $bits, $depth_bits, $axis_bits} -- src/index.rs:65-123, 293-295
template <class K, int DIM_, int DEPTH_BITS_, int AXIS_BITS_> struct IndexT {
    typedef K key_t;
    static const int DIM = DIM_;
    static const int DEPTH_BITS = DEPTH_BITS_;
    static const int AXIS_BITS = AXIS_BITS_;
    static const int ORIGIN_BITS = DIM_ * AXIS_BITS_; // src/index.rs:75
    static const int ORIGIN_SHIFT = DEPTH_BITS_;      // src/index.rs:76 (DEPTH_SHIFT = 0)
    static K depth_mask() { return (K)(((K)1 << DEPTH_BITS) - 1); }                       // :73
    static K origin_mask() { return (K)((((K)1 << ORIGIN_BITS) - 1) << ORIGIN_SHIFT); }   // :77

    // encode_axis -- src/index.rs:155-172 (2D), :193-207 (3D): the top AXIS_BITS bits of the u32
    // coordinate, bit i -> origin bit DIM*i.
    static K encode_axis(uint32_t v) {
        uint64_t top = (uint64_t)(v >> (32 - AXIS_BITS));
        return (K)(DIM == 2 ? spread2(top) : spread3(top));
    }
    // decode_axis -- src/index.rs:134-151 (2D), :176-190 (3D)
    static uint32_t decode_axis(K origin_bits) {
        return compact_bits((uint64_t)origin_bits, DIM, AXIS_BITS) << (32 - AXIS_BITS);
    }
    // clamp_depth -- src/index.rs:93-95
    static uint32_t clamp_depth(uint32_t d) { return std::min<uint32_t>(d, AXIS_BITS); }
    // depth -- src/index.rs:99-102
    static uint32_t depth(K k) { return (uint32_t)(k & depth_mask()); }
    // Index::default().set_depth(depth).set_origin(p) -- src/index.rs:106-112, 230-250
    static K make(uint32_t d, const uint32_t *p) {
        K origin = 0;
        for (int a = 0; a < DIM; ++a) origin |= (K)(encode_axis(p[a]) << a);
        K k = (K)(depth_mask() & (K)clamp_depth(d));
        k |= (K)(origin_mask() & (K)(origin << ORIGIN_SHIFT));
        return k;
    }
    // level_mask -- src/index.rs:82-86
    static K level_mask(uint32_t d) {
        if (d == 0) return 0;
        return (K)((((K)1 << (DIM * d)) - 1) << (ORIGIN_BITS + ORIGIN_SHIFT - DIM * d));
    }
    // same_cell_at_depth -- src/index.rs:120-122
    static bool same_cell_at_depth(K a, K b, uint32_t d) { return ((a ^ b) & level_mask(d)) == 0; }
    // overlaps -- src/index.rs:116-118
    static bool overlaps(K a, K b) { return same_cell_at_depth(a, b, std::min(depth(a), depth(b))); }
};

typedef IndexT<uint32_t, 2, 4, 14> Index32_2D; // src/index.rs:293
typedef IndexT<uint64_t, 2, 5, 29> Index64_2D; // src/index.rs:294
typedef IndexT<uint64_t, 3, 5, 19> Index64_3D; // src/index.rs:295

// ---------------------------------------------------------------------------------------------
// Quantiser + index generator -- src/geom.rs:40-61, 104-128, 148-163, 189-304
// ---------------------------------------------------------------------------------------------

// Rust `f32 as u32`: truncate toward zero, saturate, NaN -> 0.
inline uint32_t f32_as_u32(float x) {
    if (!(x == x)) return 0u;
    if (x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)x;
}

// SystemBounds::to_local, one scalar -- src/geom.rs:148-156.  Separate IEEE roundings for the
// subtract, divide, multiply and add (the file is compiled with -ffp-contract=off).
inline uint32_t to_local_scalar(float g, float sys_min, float sys_size) {
    const float MIN_VALUE = 0.0f;
    const float MAX_VALUE = 4294967040.0f; // 0xffff_ff00u32 as f32
    const float RANGE = MAX_VALUE - MIN_VALUE;
    const float t = g - sys_min;
    const float q = t / sys_size;
    const float m = q * RANGE;
    const float r = m + MIN_VALUE;
    return f32_as_u32(r);
}

// SystemBounds::to_global, one scalar -- src/geom.rs:165-174
inline float to_global_scalar(uint32_t l, float sys_min, float sys_size) {
    const float RANGE = 4294967040.0f;
    const float a = (float)l - 0.0f;
    const float q = a / RANGE;
    const float m = q * sys_size;
    const float r = sys_min + m;
    return r;
}

// Bounds::contains -- src/geom.rs:121-128 (NaN passes: both comparisons are false)
template <int DIM> inline bool contains(const float *sys, const float *b) {
    for (int i = 0; i < DIM; ++i)
        if (sys[i] > b[i] || sys[DIM + i] < b[DIM + i]) return false;
    return true;
}

inline uint32_t leading_zeros32(uint32_t v) { return v == 0 ? 32u : (uint32_t)__builtin_clz(v); }

// scale_at_depth / truncate_to_depth -- src/geom.rs:48-61 (depth >= 1 here)
inline uint32_t scale_at_depth(uint32_t depth) { return 1u << (32 - depth); }
inline uint32_t truncate_to_depth(uint32_t x, uint32_t depth) {
    return depth == 0 ? x : (x & ~(scale_at_depth(depth) - 1u));
}

// IndexGenerator::indices + indices_at_depth -- src/geom.rs:189-238 (2D), :247-304 (3D).
// Appends (index, id) for every cell, z outermost, then y, x innermost.
template <class Ix, class ID>
inline void gen_indices(const uint32_t *lmin, const uint32_t *lmax, uint32_t min_depth, ID id,
                        std::vector<std::pair<typename Ix::key_t, ID>> &tree) {
    const int DIM = Ix::DIM;
    // max_axis(sizei) -- src/geom.rs:40-46, 104-110; u32 wrapping arithmetic as in a release build
    uint32_t max_axis = 0;
    for (int i = 0; i < DIM; ++i) max_axis = std::max(max_axis, (uint32_t)(lmax[i] - lmin[i] + 1u));
    uint32_t depth = leading_zeros32(max_axis - 1u);
    if (depth < min_depth) depth = min_depth;
    depth = Ix::clamp_depth(depth);

    if (depth == 0) { // src/geom.rs:203-205, :261-263 -- Index::default()
        tree.emplace_back((typename Ix::key_t)0, id);
        return;
    }
    uint32_t tmin[3] = {0, 0, 0}, tmax[3] = {0, 0, 0};
    for (int i = 0; i < DIM; ++i) {
        tmin[i] = truncate_to_depth(lmin[i], depth);
        tmax[i] = truncate_to_depth(lmax[i], depth);
    }
    const uint32_t step = scale_at_depth(depth);
    uint32_t p[3];
    p[2] = tmin[2];
    for (;;) { // z loop (single trip for 2D)
        p[1] = tmin[1];
        for (;;) {
            p[0] = tmin[0];
            for (;;) {
                tree.emplace_back(Ix::make(depth, p), id);
                if (p[0] >= tmax[0]) break;
                p[0] += step;
            }
            if (p[1] >= tmax[1]) break;
            p[1] += step;
        }
        if (DIM < 3 || p[2] >= tmax[2]) break;
        p[2] += step;
    }
}

// ---------------------------------------------------------------------------------------------
// Filters (device functors on the GPU side; plain functors here)
// ---------------------------------------------------------------------------------------------
struct Filter {
    int kind;
    uint64_t arg;
    const uint32_t *table;
    size_t n_table;
    inline bool operator()(uint64_t a, uint64_t b) const {
        switch (kind) {
        case BPO_FILTER_NONE: return true;
        case BPO_FILTER_ID_PARITY: return ((a ^ b) & 1u) == 1u;
        case BPO_FILTER_XOR_MASK: return ((a ^ b) & arg) != 0;
        case BPO_FILTER_CATEGORY: {
            uint32_t ca = 0xffffffffu, ma = 0xffffffffu, cb = 0xffffffffu, mb = 0xffffffffu;
            if (a < n_table) { ca = table[2 * a]; ma = table[2 * a + 1]; }
            if (b < n_table) { cb = table[2 * b]; mb = table[2 * b + 1]; }
            return (ca & mb) != 0 && (cb & ma) != 0;
        }
        case BPO_FILTER_SPHERES: { // the narrow phase of examples/main.rs:461-479; table = n_table x {f32 x, y, z, r}
            if (a >= n_table || b >= n_table) return true;
            const float *sa = (const float *)table + 4 * a, *sb = (const float *)table + 4 * b;
            const float dx = sb[0] - sa[0], dy = sb[1] - sa[1], dz = sb[2] - sa[2];
            const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
            const float xy = xx + yy;
            const float m = xy + zz;
            const float dist = __builtin_sqrtf(m), dist_min = sa[3] + sb[3];
            return !(dist > dist_min);
        }
        default: return true;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Test geometries -- src/geom.rs:322-615.  QUERIES ARE "PARITY UNPINNED": the reference holds no test
// or fixture for test_box / test_ray, and Bounds::center (src/geom.rs:130-132) calls cgmath ^0.17's
// EuclideanSpace::midpoint (un-vendored dependency, version range from Cargo.toml:21), restated here from
// its published source as `self + (other - self) / 2`.
// ---------------------------------------------------------------------------------------------
inline float midpoint_f32(float a, float b) {
    const float d = b - a;
    const float h = d / 2.0f;
    return a + h;
}
inline bool is_finite_f32(float x) { return x - x == 0.0f; } // false for NaN and +-inf
inline float f32_max(float a, float b) { return a != a ? b : (b != b ? a : (a > b ? a : b)); } // Rust f32::max: NaN loses
inline float f32_min(float a, float b) { return a != a ? b : (b != b ? a : (a < b ? a : b)); }

// BoxTestGeometry -- src/geom.rs:353-460
template <int DIM> struct BoxGeom {
    float cmin[DIM], cmax[DIM]; // cell_bounds
    float tmin[DIM], tmax[DIM]; // test_bounds
    // with_system_bounds -- src/geom.rs:367-378; params = test_bounds (min.., max..)
    static BoxGeom make(const float *sys, const float *params) {
        BoxGeom g;
        for (int i = 0; i < DIM; ++i) {
            g.cmin[i] = sys[i];
            g.cmax[i] = sys[DIM + i];
            g.tmin[i] = params[i];
            g.tmax[i] = params[DIM + i];
        }
        return g;
    }
    // subdivide -- src/geom.rs:386-407 (2D), :425-448 (3D): child `cell`, bit `axis` set = upper half
    void subdivide(BoxGeom *out) const {
        float center[DIM];
        for (int i = 0; i < DIM; ++i) center[i] = midpoint_f32(cmin[i], cmax[i]);
        for (int cell = 0; cell < (1 << DIM); ++cell) {
            out[cell] = *this;
            for (int axis = 0; axis < DIM; ++axis) {
                if (cell & (1 << axis)) out[cell].cmin[axis] = center[axis];
                else out[cell].cmax[axis] = center[axis];
            }
        }
    }
    // test_order -- src/geom.rs:409-411: identity
    void test_order(int *order) const { for (int i = 0; i < (1 << DIM); ++i) order[i] = i; }
    // should_test -- src/geom.rs:413-416: cell_bounds.overlaps(test_bounds) (src/geom.rs:104-111)
    bool should_test(float) const {
        for (int i = 0; i < DIM; ++i)
            if (cmin[i] > tmax[i] || cmax[i] < tmin[i]) return false;
        return true;
    }
};

// RayTestGeometry -- src/geom.rs:462-615
template <int DIM> struct RayGeom {
    float cmin[DIM], cmax[DIM], origin[DIM], direction[DIM], range_min, range_max;
    // with_system_bounds -- src/geom.rs:512-535; params = origin.., direction.., range_min, range_max
    static RayGeom make(const float *sys, const float *params) {
        RayGeom g;
        g.range_min = params[2 * DIM];
        g.range_max = params[2 * DIM + 1];
        for (int i = 0; i < DIM; ++i) {
            g.cmin[i] = sys[i];
            g.cmax[i] = sys[DIM + i];
            g.origin[i] = params[i];
            g.direction[i] = params[DIM + i];
        }
        for (int axis = 0; axis < DIM; ++axis) {
            const float n0 = g.cmin[axis] - g.origin[axis];
            const float distance_0 = n0 / g.direction[axis];
            const float n1 = g.cmax[axis] - g.origin[axis];
            const float distance_1 = n1 / g.direction[axis];
            const bool is_forward = g.direction[axis] > 0.0f;
            const float d0 = is_forward ? distance_0 : distance_1;
            const float d1 = is_forward ? distance_1 : distance_0;
            if (is_finite_f32(d0)) g.range_min = f32_max(g.range_min, d0);
            if (is_finite_f32(d1)) g.range_max = f32_min(g.range_max, d1);
        }
        return g;
    }
    // subdivide -- src/geom.rs:537-577 (2D and 3D are the same loop over their axes)
    void subdivide(RayGeom *out) const {
        float center[DIM], distance[DIM];
        for (int i = 0; i < DIM; ++i) {
            center[i] = midpoint_f32(cmin[i], cmax[i]);
            const float n = center[i] - origin[i];
            distance[i] = n / direction[i];
        }
        for (int cell = 0; cell < (1 << DIM); ++cell) {
            out[cell] = *this;
            for (int axis = 0; axis < DIM; ++axis) {
                const bool side = (cell & (1 << axis)) != 0;
                if (is_finite_f32(distance[axis])) {
                    const bool is_towards = (direction[axis] > 0.0f) != side;
                    if (is_towards) out[cell].range_max = f32_min(out[cell].range_max, distance[axis]);
                    else out[cell].range_min = f32_max(out[cell].range_min, distance[axis]);
                } else if ((origin[axis] > center[axis]) != side) {
                    out[cell].range_min = __builtin_inff();
                    out[cell].range_max = -__builtin_inff();
                }
            }
            for (int axis = 0; axis < DIM; ++axis) {
                if (cell & (1 << axis)) out[cell].cmin[axis] = center[axis];
                else out[cell].cmax[axis] = center[axis];
            }
        }
    }
    // test_order -- src/geom.rs:579-610: cells nearest along the ray first (only matters for `pick`)
    void test_order(int *order) const {
        float ab[DIM];
        for (int i = 0; i < DIM; ++i) ab[i] = direction[i] < 0 ? -direction[i] : direction[i];
        int axes[3] = {0, 1, 2};
        if (DIM == 2) {
            if (!(ab[0] <= ab[1])) { axes[0] = 1; axes[1] = 0; }
        } else {
            const float x = ab[0], y = ab[1], z = ab[DIM - 1];
            if (x <= y && x <= z) { axes[0] = 0; if (y <= z) { axes[1] = 1; axes[2] = 2; } else { axes[1] = 2; axes[2] = 1; } }
            else if (y <= z) { axes[0] = 1; if (x <= z) { axes[1] = 0; axes[2] = 2; } else { axes[1] = 2; axes[2] = 0; } }
            else { axes[0] = 2; if (x <= y) { axes[1] = 0; axes[2] = 1; } else { axes[1] = 1; axes[2] = 0; } }
        }
        for (int src = 0; src < (1 << DIM); ++src) {
            int dst = 0;
            for (int k = 0; k < DIM; ++k) {
                const bool ik = ((src >> k) & 1) == (direction[axes[k]] >= 0.0f ? 1 : 0);
                dst |= (ik ? 1 : 0) << axes[k];
            }
            order[src] = dst;
        }
    }
    // should_test -- src/geom.rs:612-614
    bool should_test(float nearest) const { return range_min < range_max && range_min < nearest; }
};

// Distance functors standing in for pick_ray's `get_dist` closure (src/layer.rs:431-436); the same expressions, one
// IEEE rounding per operation, as csrc/bp_query.cuh PickSphere / PickAabb.  BPO_PICK_SPHERE is the closure of the
// reference's example (examples/main.rs:427-449); shape = centre.., radius.  BPO_PICK_AABB: slab test, shape = min.., max...
template <int DIM> inline int shape_width(int kind) { return kind == BPO_PICK_SPHERE ? DIM + 1 : 2 * DIM; }
template <int DIM> inline float shape_distance(int kind, const float *shape, const float *org, const float *dir) {
    const float INF = __builtin_inff();
    if (kind == BPO_PICK_SPHERE) {
        float proj = 0.f, mag2 = 0.f;
        for (int i = 0; i < DIM; ++i) {
            const float b = shape[i] - org[i];
            const float p = dir[i] * b, m = b * b;
            proj = i == 0 ? p : proj + p;
            mag2 = i == 0 ? m : mag2 + m;
        }
        const float r = shape[DIM];
        const float pp = proj * proj, rr = r * r;
        const float t = pp - mag2;
        const float ext = __builtin_sqrtf(t + rr);
        const float lo = proj - ext, hi = proj + ext;
        if (hi < 0.f) return INF;
        if (lo < 0.f) return 0.f;
        return lo;
    }
    float t0 = 0.f, t1 = INF;
    for (int i = 0; i < DIM; ++i) {
        if (dir[i] != 0.f) {
            const float na = shape[i] - org[i], nb = shape[DIM + i] - org[i];
            const float a = na / dir[i], b = nb / dir[i];
            t0 = f32_max(t0, f32_min(a, b));
            t1 = f32_min(t1, f32_max(a, b));
        } else if (org[i] < shape[i] || org[i] > shape[DIM + i]) {
            return INF;
        }
    }
    return t0 <= t1 ? t0 : INF;
}

// ---------------------------------------------------------------------------------------------
// Layer -- src/layer.rs:40-165, 448-573
// ---------------------------------------------------------------------------------------------
struct LayerBase {
    virtual ~LayerBase() {}
    virtual void clear() = 0;
    virtual void extend(const float *sys, const float *bounds, const void *ids, size_t n) = 0;
    virtual void merge(const LayerBase *other) = 0;
    virtual void sort(bool par) = 0;
    virtual size_t scan(const Filter &f, bool par) = 0;
    virtual size_t len() const = 0;
    virtual size_t num_collisions() const = 0;
    virtual void records(uint64_t *keys, uint64_t *ids) const = 0;
    virtual void collisions_out(uint64_t *a, uint64_t *b) const = 0;
    virtual void set_records(const uint64_t *keys, const uint64_t *ids, size_t n, bool sorted) = 0;
    virtual size_t test(int ray, const float *sys, const float *params, int max_depth) = 0;
    virtual void test_results_out(uint64_t *ids) const = 0;
    virtual int pick_ray(const float *sys, const float *ray, float max_dist, int max_depth, int shape_kind, const float *shapes,
                         size_t n_shapes, float *out, uint64_t *out_id) = 0;
    int kind = 0, id_bytes = 0;
    uint32_t min_depth = 0;
    bool sorted = true; // LayerBuilder::build starts with sorted = true -- src/layer.rs:681
    size_t raw_collisions = 0;
};

template <class Ix, class ID> struct LayerT : LayerBase {
    typedef typename Ix::key_t K;
    typedef std::pair<K, ID> Rec;   // (Index, ID): derived lexicographic Ord -- src/index.rs:67
    typedef std::pair<ID, ID> Pair; // (ID, ID)
    std::vector<Rec> tree;
    std::vector<Pair> collisions;
    std::vector<ID> invalid;
    std::vector<std::vector<Pair>> collisions_tls; // CachedThreadLocal -- src/layer.rs:67

    // Layer::clear -- src/layer.rs:84-88
    void clear() override {
        tree.clear();
        sorted = true;
    }

    // Layer::extend -- src/layer.rs:94-121 (sequential, like the reference)
    void extend(const float *sys, const float *bounds, const void *ids_, size_t n) override {
        const int DIM = Ix::DIM;
        const ID *ids = (const ID *)ids_;
        tree.reserve(tree.size() + n); // size_hint upper bound -- src/layer.rs:103-105
        float size[3];
        for (int i = 0; i < DIM; ++i) { // sizef -- src/geom.rs:97-102
            const float s = sys[DIM + i] - sys[i];
            size[i] = s;
        }
        for (size_t o = 0; o < n; ++o) {
            const float *b = bounds + o * 2 * DIM;
            if (!contains<DIM>(sys, b)) { // src/layer.rs:108-111
                invalid.push_back(ids[o]);
                continue;
            }
            uint32_t lmin[3] = {0, 0, 0}, lmax[3] = {0, 0, 0};
            for (int i = 0; i < DIM; ++i) { // to_local -- src/geom.rs:148-163
                lmin[i] = to_local_scalar(b[i], sys[i], size[i]);
                lmax[i] = to_local_scalar(b[DIM + i], sys[i], size[i]);
            }
            gen_indices<Ix, ID>(lmin, lmax, min_depth, ids[o], tree);
            sorted = false; // src/layer.rs:119
        }
    }

    // Layer::merge -- src/layer.rs:127-138
    void merge(const LayerBase *other_) override {
        const LayerT *other = static_cast<const LayerT *>(other_);
        if (other->min_depth < min_depth) min_depth = other->min_depth;
        tree.insert(tree.end(), other->tree.begin(), other->tree.end());
        sorted = false;
    }

    // Layer::sort / par_sort -- src/layer.rs:146-165
    void sort(bool par) override {
        if (sorted) return;
        if (par)
            __gnu_parallel::sort(tree.begin(), tree.end());
        else
            std::sort(tree.begin(), tree.end());
        sorted = true;
    }

    // Layer::scan_impl -- src/layer.rs:550-573 (the literal stack sweep)
    static void scan_impl(const Rec *tree, size_t n, std::vector<Pair> &out, const Filter &filter) {
        std::vector<Rec> stack;
        stack.reserve(256);
        for (size_t t = 0; t < n; ++t) {
            const K index = tree[t].first;
            const ID id = tree[t].second;
            while (!stack.empty()) {
                if (Ix::overlaps(index, stack.back().first)) break;
                stack.pop_back();
            }
            bool same = false;
            for (const Rec &s : stack)
                if (s.second == id) { same = true; break; }
            if (same) continue;
            for (const Rec &s : stack)
                if (id != s.second && filter((uint64_t)id, (uint64_t)s.second)) out.emplace_back(id, s.second);
            stack.push_back(tree[t]);
        }
    }

    // Layer::par_scan_impl -- src/layer.rs:523-548 (rayon::join -> OpenMP tasks)
    void par_scan_impl(size_t threads, const Rec *t, size_t n, const Filter &filter) {
        const size_t SPLIT_THRESHOLD = 64;
        if (threads <= 1 || n <= SPLIT_THRESHOLD) {
            scan_impl(t, n, collisions_tls[omp_get_thread_num()], filter);
            return;
        }
        size_t i = n / 2;
        while (i < n) {
            if (!Ix::same_cell_at_depth(t[i - 1].first, t[i].first, min_depth)) break;
            ++i;
        }
#pragma omp task default(shared) firstprivate(threads, t, i)
        par_scan_impl(threads >> 1, t, i, filter);
#pragma omp task default(shared) firstprivate(threads, t, i, n)
        par_scan_impl(threads >> 1, t + i, n - i, filter);
#pragma omp taskwait
    }

    // Layer::scan_filtered -- src/layer.rs:456-477; par_scan_filtered -- :489-520
    size_t scan(const Filter &filter, bool par) override {
        sort(par);
        collisions.clear();
        invalid.clear();
        if (!par) {
            scan_impl(tree.data(), tree.size(), collisions, filter);
            raw_collisions = collisions.size();
            std::sort(collisions.begin(), collisions.end());
        } else {
            const int nt = omp_get_max_threads();
            collisions_tls.resize(nt);
            for (auto &v : collisions_tls) v.clear();
#pragma omp parallel
#pragma omp single
            par_scan_impl((size_t)nt, tree.data(), tree.size(), filter);
            for (auto &v : collisions_tls) collisions.insert(collisions.end(), v.begin(), v.end());
            raw_collisions = collisions.size();
            __gnu_parallel::sort(collisions.begin(), collisions.end());
        }
        collisions.erase(std::unique(collisions.begin(), collisions.end()), collisions.end()); // dedup()
        return collisions.size();
    }

    // Layer::test_impl -- src/layer.rs:167-242, restated literally (recursion, binary searches, fold order)
    std::vector<ID> test_results;
    template <class Geom, class Callback>
    static float test_impl(const Rec *tree, size_t n, K cell, const Geom &geom, float nearest, int max_depth, Callback &callback) {
        const int NC = 1 << Ix::DIM;
        if (n == 0 || !geom.should_test(nearest)) return nearest;                         // :180-182
        const uint32_t depth = Ix::depth(cell);
        if (max_depth >= 0 && depth >= (uint32_t)max_depth) {                              // :189-197
            for (size_t i = 0; i < n; ++i) nearest = f32_min(callback(geom, nearest, tree[i].second), nearest); // the fold of :193-196
            return nearest;
        }
        if (depth < (uint32_t)Ix::AXIS_BITS) {                                             // cell.subdivide() -- src/index.rs:251-290
            K sub_cells[8];
            const int shift = Ix::ORIGIN_BITS + Ix::ORIGIN_SHIFT - Ix::DIM * ((int)depth + 1);
            for (int c = 0; c < NC; ++c) {
                K k = (K)(cell | ((K)c << shift));
                k = (K)((k & ~Ix::depth_mask()) | (K)Ix::clamp_depth(depth + 1));         // set_depth -- src/index.rs:106-112
                sub_cells[c] = k;
            }
            // :199-212 -- split the slice at every child key; the head before the first child = records at this cell
            const Rec *heads[9];
            size_t lens[9];
            const Rec *rest = tree;
            size_t nrest = n;
            for (int c = 0; c < NC; ++c) {
                size_t lo = 0, hi = nrest; // partition point of `index < cell`
                while (lo < hi) {
                    const size_t mid = lo + (hi - lo) / 2;
                    if (rest[mid].first < sub_cells[c]) lo = mid + 1; else hi = mid;
                }
                heads[c] = rest;
                lens[c] = lo;
                rest += lo;
                nrest -= lo;
            }
            heads[NC] = rest;
            lens[NC] = nrest;
            for (size_t i = 0; i < lens[0]; ++i) nearest = f32_min(callback(geom, nearest, heads[0][i].second), nearest); // :213-217
            Geom sub_tests[8];
            geom.subdivide(sub_tests);
            int order[8];
            geom.test_order(order);
            for (int k = 0; k < NC; ++k) {                                                 // :222-230
                const int i = order[k];
                nearest = test_impl(heads[i + 1], lens[i + 1], sub_cells[i], sub_tests[i], nearest, max_depth, callback);
            }
            return nearest;
        }
        for (size_t i = 0; i < n; ++i) nearest = f32_min(callback(geom, nearest, tree[i].second), nearest); // :236-240
        return nearest;
    }

    // Layer::test -- src/layer.rs:254-280; test_box -- :293-311; test_ray -- :326-351
    size_t test(int ray, const float *sys, const float *params, int max_depth) override {
        sort(false);
        test_results.clear();
        std::vector<ID> &results = test_results;
        if (ray) {
            const RayGeom<Ix::DIM> g = RayGeom<Ix::DIM>::make(sys, params);
            auto cb = [&results](const RayGeom<Ix::DIM> &, float nearest, ID id) { results.push_back(id); return nearest; }; // :267-270
            test_impl(tree.data(), tree.size(), (K)0, g, __builtin_inff(), max_depth, cb);
        } else {
            const BoxGeom<Ix::DIM> g = BoxGeom<Ix::DIM>::make(sys, params);
            auto cb = [&results](const BoxGeom<Ix::DIM> &, float nearest, ID id) { results.push_back(id); return nearest; };
            test_impl(tree.data(), tree.size(), (K)0, g, __builtin_inff(), max_depth, cb);
        }
        std::sort(test_results.begin(), test_results.end());
        test_results.erase(std::unique(test_results.begin(), test_results.end()), test_results.end());
        return test_results.size();
    }
    // Layer::pick -- src/layer.rs:364-408 and pick_ray -- :424-446, with the user's get_dist closure replaced by one of
    // the enumerated shape functors (a table of shapes indexed by ID; IDs past the table never hit).
    // out = {dist, point[DIM]}; returns 1 and *out_id if something was hit.
    int pick_ray(const float *sys, const float *ray, float max_dist, int max_depth, int shape_kind, const float *shapes, size_t n_shapes,
                 float *out, uint64_t *out_id) override {
        const int DIM = Ix::DIM;
        sort(false);                                                                        // :375
        float params[2 * 3 + 2];
        for (int i = 0; i < 2 * DIM; ++i) params[i] = ray[i];
        params[2 * DIM] = 0.0f;                                                             // with_system_bounds(.., 0f32, max_dist) -- :437-442
        params[2 * DIM + 1] = max_dist;
        const RayGeom<DIM> g = RayGeom<DIM>::make(sys, params);
        std::unordered_set<uint64_t> processed;                                             // :377, :384
        bool have = false;
        ID result = 0;
        auto cb = [&](const RayGeom<DIM> &, float nearest, ID id) -> float {                // :383-399
            if (!processed.insert((uint64_t)id).second) return __builtin_inff();
            const float dist = (uint64_t)id < n_shapes ? shape_distance<DIM>(shape_kind, shapes + (size_t)id * shape_width<DIM>(shape_kind), ray, ray + DIM)
                                                       : __builtin_inff();
            if (is_finite_f32(dist)) {
                if (dist < nearest) {
                    result = id;
                    have = true;
                }
                return dist;
            }
            return __builtin_inff();
        };
        const float dist = test_impl(tree.data(), tree.size(), (K)0, g, max_dist, max_depth, cb);
        if (!have) return 0;
        out[0] = dist;
        for (int i = 0; i < DIM; ++i) {                                                     // origin + direction * dist -- :443-446
            const float m = ray[DIM + i] * dist;
            out[1 + i] = ray[i] + m;
        }
        *out_id = (uint64_t)result;
        return 1;
    }

    void test_results_out(uint64_t *ids) const override {
        for (size_t i = 0; i < test_results.size(); ++i) ids[i] = (uint64_t)test_results[i];
    }

    size_t len() const override { return tree.size(); }
    size_t num_collisions() const override { return collisions.size(); }
    void records(uint64_t *keys, uint64_t *ids) const override {
        for (size_t i = 0; i < tree.size(); ++i) {
            keys[i] = (uint64_t)tree[i].first;
            ids[i] = (uint64_t)tree[i].second;
        }
    }
    void collisions_out(uint64_t *a, uint64_t *b) const override {
        for (size_t i = 0; i < collisions.size(); ++i) {
            a[i] = (uint64_t)collisions[i].first;
            b[i] = (uint64_t)collisions[i].second;
        }
    }
    void set_records(const uint64_t *keys, const uint64_t *ids, size_t n, bool sorted_) override {
        tree.resize(n);
        for (size_t i = 0; i < n; ++i) tree[i] = Rec((K)keys[i], (ID)ids[i]);
        sorted = sorted_;
    }
};

LayerBase *make_layer(int kind, int id_bytes) {
    if (id_bytes == 4) {
        if (kind == BPO_INDEX32_2D) return new LayerT<Index32_2D, uint32_t>();
        if (kind == BPO_INDEX64_2D) return new LayerT<Index64_2D, uint32_t>();
        if (kind == BPO_INDEX64_3D) return new LayerT<Index64_3D, uint32_t>();
    } else if (id_bytes == 8) {
        if (kind == BPO_INDEX32_2D) return new LayerT<Index32_2D, uint64_t>();
        if (kind == BPO_INDEX64_2D) return new LayerT<Index64_2D, uint64_t>();
        if (kind == BPO_INDEX64_3D) return new LayerT<Index64_3D, uint64_t>();
    }
    return nullptr;
}

} // namespace

struct bpo_layer {
    LayerBase *impl;
};

extern "C" {

bpo_layer *bpo_layer_new(int kind, int id_bytes, uint32_t min_depth) {
    LayerBase *impl = make_layer(kind, id_bytes);
    if (!impl) return nullptr;
    impl->kind = kind;
    impl->id_bytes = id_bytes;
    impl->min_depth = min_depth;
    bpo_layer *l = new bpo_layer;
    l->impl = impl;
    return l;
}
void bpo_layer_free(bpo_layer *l) {
    if (!l) return;
    delete l->impl;
    delete l;
}
void bpo_layer_clear(bpo_layer *l) { l->impl->clear(); }
void bpo_layer_extend(bpo_layer *l, const float *sys, const float *bounds, const void *ids, size_t n) {
    l->impl->extend(sys, bounds, ids, n);
}
void bpo_layer_merge(bpo_layer *l, const bpo_layer *other) { l->impl->merge(other->impl); }
void bpo_layer_sort(bpo_layer *l) { l->impl->sort(false); }
void bpo_layer_par_sort(bpo_layer *l) { l->impl->sort(true); }
size_t bpo_layer_scan(bpo_layer *l, int filter, uint64_t arg, const uint32_t *table, size_t n_table) {
    Filter f = {filter, arg, table, n_table};
    return l->impl->scan(f, false);
}
size_t bpo_layer_par_scan(bpo_layer *l, int filter, uint64_t arg, const uint32_t *table, size_t n_table) {
    Filter f = {filter, arg, table, n_table};
    return l->impl->scan(f, true);
}
size_t bpo_layer_len(const bpo_layer *l) { return l->impl->len(); }
int bpo_layer_sorted(const bpo_layer *l) { return l->impl->sorted ? 1 : 0; }
uint32_t bpo_layer_min_depth(const bpo_layer *l) { return l->impl->min_depth; }
size_t bpo_layer_num_collisions(const bpo_layer *l) { return l->impl->num_collisions(); }
size_t bpo_layer_num_raw_collisions(const bpo_layer *l) { return l->impl->raw_collisions; }
void bpo_layer_records(const bpo_layer *l, uint64_t *keys, uint64_t *ids) { l->impl->records(keys, ids); }
void bpo_layer_collisions(const bpo_layer *l, uint64_t *a, uint64_t *b) { l->impl->collisions_out(a, b); }
void bpo_layer_set_records(bpo_layer *l, const uint64_t *keys, const uint64_t *ids, size_t n, int sorted) {
    l->impl->set_records(keys, ids, n, sorted != 0);
}

uint64_t bpo_encode_axis(int kind, uint32_t v) {
    switch (kind) {
    case BPO_INDEX32_2D: return Index32_2D::encode_axis(v);
    case BPO_INDEX64_2D: return Index64_2D::encode_axis(v);
    default: return Index64_3D::encode_axis(v);
    }
}
uint32_t bpo_decode_axis(int kind, uint64_t o) {
    switch (kind) {
    case BPO_INDEX32_2D: return Index32_2D::decode_axis((uint32_t)o);
    case BPO_INDEX64_2D: return Index64_2D::decode_axis(o);
    default: return Index64_3D::decode_axis(o);
    }
}
uint64_t bpo_make_index(int kind, uint32_t depth, const uint32_t *origin) {
    switch (kind) {
    case BPO_INDEX32_2D: return Index32_2D::make(depth, origin);
    case BPO_INDEX64_2D: return Index64_2D::make(depth, origin);
    default: return Index64_3D::make(depth, origin);
    }
}
uint64_t bpo_level_mask(int kind, uint32_t depth) {
    switch (kind) {
    case BPO_INDEX32_2D: return Index32_2D::level_mask(depth);
    case BPO_INDEX64_2D: return Index64_2D::level_mask(depth);
    default: return Index64_3D::level_mask(depth);
    }
}
int bpo_overlaps(int kind, uint64_t a, uint64_t b) {
    switch (kind) {
    case BPO_INDEX32_2D: return Index32_2D::overlaps((uint32_t)a, (uint32_t)b);
    case BPO_INDEX64_2D: return Index64_2D::overlaps(a, b);
    default: return Index64_3D::overlaps(a, b);
    }
}
void bpo_to_local(int dim, const float *sys, const float *b, uint32_t *out) {
    for (int i = 0; i < dim; ++i) {
        const float s = sys[dim + i] - sys[i];
        out[i] = to_local_scalar(b[i], sys[i], s);
        out[dim + i] = to_local_scalar(b[dim + i], sys[i], s);
    }
}
void bpo_to_global(int dim, const float *sys, const uint32_t *local, float *out) {
    for (int i = 0; i < dim; ++i) {
        const float s = sys[dim + i] - sys[i];
        out[i] = to_global_scalar(local[i], sys[i], s);
        out[dim + i] = to_global_scalar(local[dim + i], sys[i], s);
    }
}

size_t bpo_layer_test_box(bpo_layer *l, const float *sys, const float *box, int max_depth) { return l->impl->test(0, sys, box, max_depth); }
size_t bpo_layer_test_ray(bpo_layer *l, const float *sys, const float *ray, int max_depth) { return l->impl->test(1, sys, ray, max_depth); }
void bpo_layer_test_results(const bpo_layer *l, uint64_t *ids) { l->impl->test_results_out(ids); }
int bpo_layer_pick_ray(bpo_layer *l, const float *sys, const float *ray, float max_dist, int max_depth, int shape_kind, const float *shapes,
                       size_t n_shapes, float *out, uint64_t *out_id) {
    return l->impl->pick_ray(sys, ray, max_dist, max_depth, shape_kind, shapes, n_shapes, out, out_id);
}

int bpo_max_threads(void) { return omp_get_max_threads(); }
void bpo_set_threads(int n) { omp_set_num_threads(n); }

} // extern "C"
