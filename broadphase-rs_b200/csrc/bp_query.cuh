// bp_query.cuh -- batched box / ray queries against the sorted tree.
//
// Replaces Layer::test / test_box / test_ray (src/layer.rs:167-351) for a whole batch of test
// geometries at once.  The reference walks the cell hierarchy recursively (test_impl,
// src/layer.rs:167-242): at a cell it splits its slice of the sorted tree at the 2^D child keys by
// binary search (:201-211), reports the records that sit at the cell itself (:214-217), subdivides
// the test geometry alongside (TestGeometry::subdivide, src/geom.rs:386-407 / 537-577) and descends
// into every child whose slice is not empty and whose geometry still passes should_test
// (src/geom.rs:413-416, 612-614); at max_depth (or the deepest level) it reports the whole slice
// (:189-197, :236-240).  `test` then sorts and deduplicates the reported IDs (:276-277).
//
// Here one warp owns one query and runs the same descent with an explicit stack in shared memory:
// the 2^D child boundaries are 2^D binary searches on 2^D lanes, the same lanes subdivide the
// geometry for "their" child, and reporting a slice is a coalesced copy of its IDs.  A stack entry is
// only (cell key, slice): the geometry of a popped cell is rebuilt from the root by replaying the
// key's Morton digits, with exactly the reference's f32 operations in the reference's order (midpoint
// = min + (max - min) / 2, cgmath 0.17 EuclideanSpace::midpoint; ray distances = (center - origin) /
// direction), so the same cells pass and fail.  The traversal runs twice: a counting pass, an
// exclusive scan over the queries, and a writing pass that needs no atomics and leaves the raw
// (query, ID) pairs grouped by query; pair_finish_kernel then orders and deduplicates every group.
// The reported order (test_order, src/geom.rs:409-411, 579-610) does not matter for `test` because of
// the final sort.  pick_ray (further down) walks in that order, prunes with the nearest hit so far and asks an
// enumerated shape functor -- in place of the user's closure -- for the hit distances.
#pragma once

#include "bp_common.cuh"

namespace bp {

// ---- test geometries -------------------------------------------------------------------------------
// Both keep the f32 bounds of the current cell (cell_bounds) and are refined per level by child(c):
// bit `axis` of c set = upper half along that axis (SpatialIndex::subdivide order, src/index.rs:251-290).
template <int DIM> struct BoxTestGeom { // BoxTestGeometry -- src/geom.rs:353-460
    static constexpr int PARAMS = 2 * DIM; // test_bounds: min.., max..
    float cmin[DIM], cmax[DIM];
    float tmin[DIM], tmax[DIM];
    __device__ __forceinline__ void init(const float *sysb, const float *q) { // with_system_bounds -- src/geom.rs:367-378
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            cmin[i] = sysb[i];
            cmax[i] = sysb[DIM + i];
            tmin[i] = q[i];
            tmax[i] = q[DIM + i];
        }
    }
    __device__ __forceinline__ void child(uint32_t c) { // subdivide -- src/geom.rs:386-407, 425-448
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const float center = __fadd_rn(cmin[i], __fdiv_rn(__fsub_rn(cmax[i], cmin[i]), 2.0f));
            if ((c >> i) & 1u)
                cmin[i] = center;
            else
                cmax[i] = center;
        }
    }
    __device__ __forceinline__ bool should_test() const { // cell_bounds.overlaps(test_bounds) -- src/geom.rs:104-111, 413-416
#pragma unroll
        for (int i = 0; i < DIM; ++i)
            if (cmin[i] > tmax[i] || cmax[i] < tmin[i]) return false;
        return true;
    }
};

template <int DIM> struct RayTestGeom { // RayTestGeometry -- src/geom.rs:462-615
    static constexpr int PARAMS = 2 * DIM + 2; // origin.., direction.., range_min, range_max
    float cmin[DIM], cmax[DIM];
    float org[DIM], dir[DIM];
    float rmin, rmax;
    static __device__ __forceinline__ bool finite(float x) { return fabsf(x) < __int_as_float(0x7f800000); } // false for NaN
    __device__ __forceinline__ void init(const float *sysb, const float *q) { // with_system_bounds -- src/geom.rs:512-535
        rmin = q[2 * DIM];
        rmax = q[2 * DIM + 1];
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            cmin[i] = sysb[i];
            cmax[i] = sysb[DIM + i];
            org[i] = q[i];
            dir[i] = q[DIM + i];
            const float dist0 = __fdiv_rn(__fsub_rn(cmin[i], org[i]), dir[i]);
            const float dist1 = __fdiv_rn(__fsub_rn(cmax[i], org[i]), dir[i]);
            const bool forward = dir[i] > 0.0f;
            const float d0 = forward ? dist0 : dist1, d1 = forward ? dist1 : dist0;
            if (finite(d0)) rmin = fmaxf(rmin, d0); // f32::max / f32::min: the non-NaN operand wins, like fmaxf / fminf
            if (finite(d1)) rmax = fminf(rmax, d1);
        }
    }
    __device__ __forceinline__ void child(uint32_t c) { // subdivide -- src/geom.rs:537-577
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const float center = __fadd_rn(cmin[i], __fdiv_rn(__fsub_rn(cmax[i], cmin[i]), 2.0f));
            const float dist = __fdiv_rn(__fsub_rn(center, org[i]), dir[i]);
            const bool side = ((c >> i) & 1u) != 0;
            if (finite(dist)) {
                const bool towards = (dir[i] > 0.0f) != side;
                if (towards)
                    rmax = fminf(rmax, dist);
                else
                    rmin = fmaxf(rmin, dist);
            } else if ((org[i] > center) != side) {
                rmin = __int_as_float(0x7f800000);
                rmax = __int_as_float(0xff800000);
            }
            if (side)
                cmin[i] = center;
            else
                cmax[i] = center;
        }
    }
    // should_test(nearest) -- src/geom.rs:612-614; `test` passes +inf, `pick` the nearest hit so far
    __device__ __forceinline__ bool should_test(float nearest) const { return rmin < rmax && rmin < nearest; }
    __device__ __forceinline__ bool should_test() const { return should_test(__int_as_float(0x7f800000)); }
    // test_order -- src/geom.rs:579-610: the order in which the children are visited (it decides which cells a
    // pick can prune and, between equal distances, which ID wins); order[k] = k-th child to visit
    __device__ __forceinline__ void test_order(uint32_t *order) const {
        float ab[DIM];
#pragma unroll
        for (int i = 0; i < DIM; ++i) ab[i] = fabsf(dir[i]);
        int ax[3] = {0, 1, 2};
        if (DIM == 2) {
            if (!(ab[0] <= ab[1])) {
                ax[0] = 1;
                ax[1] = 0;
            }
        } else {
            const float x = ab[0], y = ab[1], z = ab[DIM - 1];
            if (x <= y && x <= z) {
                ax[0] = 0;
                ax[1] = y <= z ? 1 : 2;
                ax[2] = y <= z ? 2 : 1;
            } else if (y <= z) {
                ax[0] = 1;
                ax[1] = x <= z ? 0 : 2;
                ax[2] = x <= z ? 2 : 0;
            } else {
                ax[0] = 2;
                ax[1] = x <= y ? 0 : 1;
                ax[2] = x <= y ? 1 : 0;
            }
        }
        for (uint32_t src = 0; src < (1u << DIM); ++src) {
            uint32_t dst = 0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const bool ik = (((src >> k) & 1u) != 0) == (dir[ax[k]] >= 0.0f);
                dst |= (ik ? 1u : 0u) << ax[k];
            }
            order[src] = dst;
        }
    }
};

// ---- distance functors for pick_ray -----------------------------------------------------------------------
// Layer::pick_ray (src/layer.rs:424-446) asks a user closure for the distance at which the ray hits object `id`
// (+inf: no hit).  A closure cannot cross the ABI; these device functors over a per-ID shape table stand in, like
// the enumerated scan filters.  Pure functions of (ray, shape): the reference's `processed` set (an ID is asked
// once per pick, src/layer.rs:384-399) then has no observable effect.  Every operation is one IEEE rounding, in
// the order written (the oracle computes the same expression).
template <int DIM> struct PickSphere { // the closure of the reference's example, examples/main.rs:427-449
    static constexpr int WIDTH = DIM + 1; // centre.., radius
    __device__ __forceinline__ static float dist(const float *shape, const float *org, const float *dir) {
        float proj = 0.f, mag2 = 0.f;
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const float b = __fsub_rn(shape[i], org[i]); // ball_dir = centre - ray_origin
            const float p = __fmul_rn(dir[i], b), m = __fmul_rn(b, b);
            proj = i == 0 ? p : __fadd_rn(proj, p);       // ray_direction.dot(ball_dir)
            mag2 = i == 0 ? m : __fadd_rn(mag2, m);       // ball_dir.magnitude2()
        }
        const float r = shape[DIM];
        const float ext = __fsqrt_rn(__fadd_rn(__fsub_rn(__fmul_rn(proj, proj), mag2), __fmul_rn(r, r)));
        const float lo = __fsub_rn(proj, ext), hi = __fadd_rn(proj, ext);
        if (hi < 0.f) return __int_as_float(0x7f800000);
        if (lo < 0.f) return 0.f;
        return lo; // NaN when the ray misses (negative radicand): not finite -> no hit
    }
};
template <int DIM> struct PickAabb { // slab test against the object's own bounds
    static constexpr int WIDTH = 2 * DIM; // min.., max..
    __device__ __forceinline__ static float dist(const float *shape, const float *org, const float *dir) {
        float t0 = 0.f, t1 = __int_as_float(0x7f800000);
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            if (dir[i] != 0.f) {
                const float a = __fdiv_rn(__fsub_rn(shape[i], org[i]), dir[i]);
                const float b = __fdiv_rn(__fsub_rn(shape[DIM + i], org[i]), dir[i]);
                t0 = fmaxf(t0, fminf(a, b));
                t1 = fminf(t1, fmaxf(a, b));
            } else if (org[i] < shape[i] || org[i] > shape[DIM + i]) {
                return __int_as_float(0x7f800000);
            }
        }
        return t0 <= t1 ? t0 : __int_as_float(0x7f800000);
    }
};


// ---- traversal ------------------------------------------------------------------------------------------
constexpr int QUERY_WARPS = 4;
constexpr int QUERY_THREADS = QUERY_WARPS * 32;

template <class T, class IdT> struct QueryArgs {
    const typename T::key_t *keys; // sorted tree
    const IdT *ids;
    IdT id_mask;                   // removes the cell flags a sorted tree may carry in its IDs' top bits
    uint32_t n;                    // records
    const float *params;           // device, n_queries x Geom::PARAMS
    uint32_t n_queries;
    float sysb[6];                 // system bounds: min.., max..
    int max_depth;                 // < 0: none
    uint32_t *counts;              // COUNT pass: reported records per query; WRITE pass: exclusive prefix of the counts
    uint64_t *out_packed;          // u32 IDs: (query << 32) | id
    uint64_t *out_a;               // u64 IDs: query
    uint64_t *out_b;               //          id
    unsigned long long *total;     // COUNT pass: sum of all counts
    int *err;                      // set if the stack overflowed (cannot happen for a valid tree)
};

template <class T> struct QueryStack {
    static constexpr int CHILDREN = 1 << T::DIM;
    static constexpr int ENTRIES = (CHILDREN - 1) * (T::AXIS_BITS + 1) + CHILDREN + 1; // DFS bound
};

template <class T, class IdT, class Geom, bool COUNT>
__global__ void __launch_bounds__(QUERY_THREADS) query_kernel(const QueryArgs<T, IdT> a) {
    typedef typename T::key_t K;
    constexpr int DIM = T::DIM, CHILDREN = 1 << DIM;
    constexpr int TOTAL = DIM * T::AXIS_BITS + T::DEPTH_BITS;
    constexpr int ENTRIES = QueryStack<T>::ENTRIES;
    constexpr bool WIDE = sizeof(IdT) == 8;
    constexpr K DEPTH_MASK = (K)((1u << T::DEPTH_BITS) - 1u);
    __shared__ K skey[QUERY_WARPS][ENTRIES];
    __shared__ uint2 srange[QUERY_WARPS][ENTRIES];

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    K *stk = skey[warp];
    uint2 *str = srange[warp];
    unsigned long long block_total = 0;

    for (uint32_t q = blockIdx.x * QUERY_WARPS + warp; q < a.n_queries; q += gridDim.x * QUERY_WARPS) {
        const float *qp = a.params + (size_t)q * Geom::PARAMS;
        uint32_t reported = 0;                          // COUNT: records reported so far
        uint64_t wpos = COUNT ? 0 : (uint64_t)a.counts[q]; // WRITE: next output slot of this query
        int sp = 0;
        {
            Geom root;
            root.init(a.sysb, qp);
            if (a.n != 0 && root.should_test()) { // test_impl's entry check -- src/layer.rs:180-182
                if (lane == 0) {
                    stk[0] = (K)0; // Index::default(): depth 0, the whole system
                    str[0] = make_uint2(0u, a.n);
                }
                sp = 1;
            }
        }
        __syncwarp();
        while (sp > 0) {
            --sp;
            const K cell = stk[sp];
            const uint2 range = str[sp];
            __syncwarp();
            const uint32_t depth = (uint32_t)(cell & DEPTH_MASK);
            // the cell's geometry: replay the subdivisions from the root (uniform across the warp)
            Geom g;
            g.init(a.sysb, qp);
            for (uint32_t l = 1; l <= depth; ++l) g.child((uint32_t)(cell >> (TOTAL - DIM * (int)l)) & (CHILDREN - 1));

            uint32_t rep_lo = range.x, rep_hi = range.y; // slice to report at this cell
            const bool leaf = (a.max_depth >= 0 && depth >= (uint32_t)a.max_depth) || depth >= (uint32_t)T::AXIS_BITS;
            if (!leaf) {
                // SpatialIndex::subdivide -- src/index.rs:251-290: child c = cell | c << shift, depth + 1
                const int shift = TOTAL - DIM * ((int)depth + 1);
                const uint32_t c = lane & (CHILDREN - 1);
                const K child_key = (K)(((cell & ~DEPTH_MASK) | ((K)c << shift)) | (K)(depth + 1));
                // first record >= child_key inside the slice -- src/layer.rs:203-206
                uint32_t lo = range.x, hi = range.y;
                if (lane < (unsigned)CHILDREN) {
                    while (lo < hi) {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if (a.keys[mid] < child_key)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                }
                const uint32_t begin = lo;
                uint32_t end = __shfl_down_sync(BP_FULL_MASK, begin, 1);
                if (lane == (unsigned)CHILDREN - 1) end = range.y;
                rep_hi = __shfl_sync(BP_FULL_MASK, begin, 0); // records before the first child sit at this cell itself
                bool push = false;
                if (lane < (unsigned)CHILDREN && begin < end) {
                    Geom gc = g;
                    gc.child(c);
                    push = gc.should_test();
                }
                const uint32_t pm = __ballot_sync(BP_FULL_MASK, push);
                if (sp + __popc(pm) > ENTRIES) {
                    if (lane == 0) *a.err = 1;
                } else if (push) {
                    const int slot = sp + __popc(pm & lt);
                    stk[slot] = child_key;
                    str[slot] = make_uint2(begin, end);
                }
                if (sp + __popc(pm) <= ENTRIES) sp += __popc(pm);
                __syncwarp();
            }
            const uint32_t nrep = rep_hi - rep_lo;
            if (COUNT) {
                reported += nrep;
            } else {
                for (uint32_t i = lane; i < nrep; i += 32) {
                    const IdT id = a.ids[rep_lo + i] & a.id_mask;
                    if (WIDE) {
                        a.out_a[wpos + i] = (uint64_t)q;
                        a.out_b[wpos + i] = (uint64_t)id;
                    } else {
                        a.out_packed[wpos + i] = ((uint64_t)q << 32) | (uint64_t)id;
                    }
                }
                wpos += nrep;
            }
        }
        if (COUNT && lane == 0) {
            a.counts[q] = reported;
            block_total += reported;
        }
    }
    if (COUNT && lane == 0 && block_total) atomicAdd(a.total, block_total);
}

// ---- pick ---------------------------------------------------------------------------------------------------
// Layer::pick / pick_ray -- src/layer.rs:364-446: the same descent, but the children are visited in test_order,
// every visited record is asked for its distance, the smallest finite one so far (`nearest`) prunes the cells
// whose part of the ray starts behind it, and the first ID to reach the final minimum wins.  One warp per ray
// replays that sequential walk exactly: children are pushed in reverse visiting order, should_test(nearest) is
// evaluated when a cell is popped, and a slice is evaluated 32 records at a time with (minimum, lowest position)
// selection -- what the reference's left-to-right fold with its strict `<` computes.
struct PickResult { // one per ray; 32 bytes
    float dist;          // nearest hit distance (valid if hit)
    uint32_t hit;        // 0: None
    unsigned long long id;
    float point[3];      // origin + direction * dist
    uint32_t pad;
};

template <class T, class IdT> struct PickArgs {
    const typename T::key_t *keys;
    const IdT *ids;
    IdT id_mask;
    uint32_t n;
    const float *rays;    // device, n_queries x 2*DIM: origin.., direction..
    uint32_t n_queries;
    float sysb[6];
    float max_dist;
    int max_depth;
    const float *shapes;  // device, n_shapes x Shape::WIDTH, indexed by ID
    unsigned long long n_shapes;
    PickResult *out;
    int *err;
};

template <class T, class IdT, class Shape>
__global__ void __launch_bounds__(QUERY_THREADS) pick_kernel(const PickArgs<T, IdT> a) {
    typedef typename T::key_t K;
    typedef RayTestGeom<T::DIM> Geom;
    constexpr int DIM = T::DIM, CHILDREN = 1 << DIM;
    constexpr int TOTAL = DIM * T::AXIS_BITS + T::DEPTH_BITS;
    constexpr int ENTRIES = QueryStack<T>::ENTRIES;
    constexpr K DEPTH_MASK = (K)((1u << T::DEPTH_BITS) - 1u);
    const float INF = __int_as_float(0x7f800000);
    __shared__ K skey[QUERY_WARPS][ENTRIES];
    __shared__ uint2 srange[QUERY_WARPS][ENTRIES];

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    K *stk = skey[warp];
    uint2 *str = srange[warp];

    for (uint32_t q = blockIdx.x * QUERY_WARPS + warp; q < a.n_queries; q += gridDim.x * QUERY_WARPS) {
        float qp[2 * DIM + 2]; // RayTestGeometry::with_system_bounds(.., 0, max_dist) -- src/layer.rs:437-442
#pragma unroll
        for (int i = 0; i < 2 * DIM; ++i) qp[i] = a.rays[(size_t)q * 2 * DIM + i];
        qp[2 * DIM] = 0.f;
        qp[2 * DIM + 1] = a.max_dist;
        float nearest = a.max_dist;
        bool hit = false;
        unsigned long long best = 0;
        uint32_t order[CHILDREN];
        int sp = 0;
        {
            Geom root;
            root.init(a.sysb, qp);
            root.test_order(order);
            if (a.n != 0) {
                if (lane == 0) {
                    stk[0] = (K)0;
                    str[0] = make_uint2(0u, a.n);
                }
                sp = 1;
            }
        }
        __syncwarp();
        while (sp > 0) {
            --sp;
            const K cell = stk[sp];
            const uint2 range = str[sp];
            __syncwarp();
            const uint32_t depth = (uint32_t)(cell & DEPTH_MASK);
            Geom g;
            g.init(a.sysb, qp);
            for (uint32_t l = 1; l <= depth; ++l) g.child((uint32_t)(cell >> (TOTAL - DIM * (int)l)) & (CHILDREN - 1));
            if (!g.should_test(nearest)) continue; // test_impl's entry check with the nearest hit so far -- src/layer.rs:180-182

            uint32_t rep_lo = range.x, rep_hi = range.y;
            const bool leaf = (a.max_depth >= 0 && depth >= (uint32_t)a.max_depth) || depth >= (uint32_t)T::AXIS_BITS;
            uint32_t begin = 0, end = 0;
            K child_key = 0;
            if (!leaf) {
                const int shift = TOTAL - DIM * ((int)depth + 1);
                const uint32_t c = lane & (CHILDREN - 1);
                child_key = (K)(((cell & ~DEPTH_MASK) | ((K)c << shift)) | (K)(depth + 1));
                uint32_t lo = range.x, hi = range.y;
                if (lane < (unsigned)CHILDREN) {
                    while (lo < hi) {
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if (a.keys[mid] < child_key)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                }
                begin = lo;
                end = __shfl_down_sync(BP_FULL_MASK, begin, 1);
                if (lane == (unsigned)CHILDREN - 1) end = range.y;
                rep_hi = __shfl_sync(BP_FULL_MASK, begin, 0);
            }
            // the records at this cell (or, at a leaf, its whole slice), in order -- src/layer.rs:189-197, 213-217
            for (uint32_t base = rep_lo; base < rep_hi; base += 32) {
                float d = INF;
                unsigned long long id = 0;
                if (base + lane < rep_hi) {
                    id = (unsigned long long)(a.ids[base + lane] & a.id_mask);
                    if (id < a.n_shapes) {
                        const float v = Shape::dist(a.shapes + (size_t)id * Shape::WIDTH, qp, qp + DIM);
                        if (fabsf(v) < INF) d = v; // `if dist.is_finite() { .. dist } else { INFINITY }` -- src/layer.rs:388-396
                    }
                }
                float m = d;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(BP_FULL_MASK, m, o));
                if (m < nearest) { // strict: an equal distance later in the walk does not replace the earlier ID
                    const uint32_t who = __ballot_sync(BP_FULL_MASK, d == m);
                    best = __shfl_sync(BP_FULL_MASK, id, __ffs(who) - 1);
                    nearest = m;
                    hit = true;
                }
            }
            if (!leaf) {
                // children in REVERSE visiting order, so that the first one to visit is popped first; empty slices are
                // skipped (test_impl returns at once for them), should_test is evaluated at the pop
                int pushed = 0;
                for (int k = CHILDREN - 1; k >= 0; --k) {
                    const uint32_t c = order[k];
                    const uint32_t cb = __shfl_sync(BP_FULL_MASK, begin, c), ce = __shfl_sync(BP_FULL_MASK, end, c);
                    const K ck = (K)__shfl_sync(BP_FULL_MASK, (unsigned long long)child_key, c);
                    if (cb < ce) {
                        if (sp + pushed < ENTRIES) {
                            if (lane == 0) {
                                stk[sp + pushed] = ck;
                                str[sp + pushed] = make_uint2(cb, ce);
                            }
                            ++pushed;
                        } else if (lane == 0) {
                            *a.err = 1;
                        }
                    }
                }
                sp += pushed;
                __syncwarp();
            }
        }
        if (lane == 0) {
            PickResult r;
            r.dist = nearest;
            r.hit = hit ? 1u : 0u;
            r.id = best;
            r.pad = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) r.point[i] = i < DIM ? __fadd_rn(qp[i], __fmul_rn(qp[DIM + i], nearest)) : 0.f; // origin + direction * dist
            a.out[q] = r;
        }
    }
}

// CSR offsets of the final (query, id) pairs: offsets[q] = first pair of query q, offsets[n_queries] = n_pairs.
template <class IdT>
__global__ void __launch_bounds__(256) query_offsets_kernel(const IdT *__restrict__ pairs, uint32_t n_pairs, uint32_t n_queries,
                                                             uint32_t *__restrict__ offsets) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_pairs) return;
    const uint64_t q_here = i < n_pairs ? (uint64_t)pairs[2 * (size_t)i] : (uint64_t)n_queries;
    const uint64_t q_first = i > 0 ? (uint64_t)pairs[2 * (size_t)(i - 1)] + 1 : 0;
    for (uint64_t q = q_first; q <= q_here; ++q) offsets[q] = i;
}

} // namespace bp
