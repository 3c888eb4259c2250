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
// the final sort; `pick` (first hit along the ray with a user closure) is not offered.
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
    // should_test(nearest = +inf) -- src/geom.rs:612-614
    __device__ __forceinline__ bool should_test() const { return rmin < rmax && rmin < __int_as_float(0x7f800000); }
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
