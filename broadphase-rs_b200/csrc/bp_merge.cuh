// bp_merge.cuh -- K5: merge-path merge of two sorted record runs.
//
// The reference's Layer::merge (src/layer.rs:127-138) appends the other tree and clears the sorted
// flag, so the next sort (src/layer.rs:146-165) re-sorts the whole concatenation.  When both parts
// are already sorted (the "static scene layer merged into this frame's dynamic layer" use-case of
// the reference's README), the same total order is produced by one linear merge: (Index, ID)
// lexicographic, src/index.rs:67.
//
//   merge_partition_kernel  one thread per output tile: bisects the tile's diagonal of the merge
//                           matrix for its split point (a_i + b_i = tile * MERGE_TILE).
//   merge_tiles_kernel      one CTA per output tile: stages its A and B segments in shared memory
//                           (coalesced), every thread bisects its own diagonal inside the tile and
//                           merges MERGE_IPT records sequentially, results are staged and written
//                           coalesced.
#pragma once

#include "bp_common.cuh"

namespace bp {

constexpr int MERGE_THREADS = 256;
constexpr int MERGE_IPT = 8;
constexpr int MERGE_TILE = MERGE_THREADS * MERGE_IPT;

template <class K, class V> struct MergeArgs {
    const K *ka;
    const V *va;
    uint32_t na;
    const K *kb;
    const V *vb;
    uint32_t nb;
    K *kout;
    V *vout;
    uint32_t *partition; // [tiles + 1]
    V id_mask;           // IDs compare under this mask (cell flags may ride in their top 3 bits: dedup at the source)
};

// (ka, va) sorts strictly before (kb, vb)
template <class K, class V> __device__ __forceinline__ bool rec_less(K ka, V va, K kb, V vb) {
    return ka < kb || (ka == kb && va < vb);
}

template <class K, class V>
__global__ void __launch_bounds__(256) merge_partition_kernel(const MergeArgs<K, V> a, uint32_t tiles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles) return;
    const uint64_t total = (uint64_t)a.na + a.nb;
    const uint64_t diag = min((uint64_t)t * MERGE_TILE, total);
    // smallest ai such that A[ai] does not precede B[diag - 1 - ai]  (A wins ties: stable)
    uint32_t lo = (uint32_t)(diag > a.nb ? diag - a.nb : 0), hi = (uint32_t)min(diag, (uint64_t)a.na);
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint32_t bi = (uint32_t)(diag - 1 - mid);
        if (!rec_less(a.kb[bi], (V)(a.vb[bi] & a.id_mask), a.ka[mid], (V)(a.va[mid] & a.id_mask)))
            lo = mid + 1;
        else
            hi = mid;
    }
    a.partition[t] = lo;
}

template <class K, class V>
__global__ void __launch_bounds__(MERGE_THREADS) merge_tiles_kernel(const MergeArgs<K, V> a) {
    __shared__ K sk[MERGE_TILE];
    __shared__ V sv[MERGE_TILE];
    const unsigned tid = threadIdx.x;
    const uint32_t t = blockIdx.x;
    const uint64_t total = (uint64_t)a.na + a.nb;
    const uint64_t d0 = (uint64_t)t * MERGE_TILE;
    const uint32_t tile_n = (uint32_t)min((uint64_t)MERGE_TILE, total - d0);
    const uint32_t a0 = a.partition[t], a1 = a.partition[t + 1];
    const uint32_t b0 = (uint32_t)(d0 - a0);
    const uint32_t la = a1 - a0, lb = tile_n - la; // segment lengths; A occupies sk[0, la), B sk[la, tile_n)

    for (uint32_t i = tid; i < tile_n; i += MERGE_THREADS) {
        if (i < la) {
            sk[i] = a.ka[a0 + i];
            sv[i] = a.va[a0 + i];
        } else {
            sk[i] = a.kb[b0 + (i - la)];
            sv[i] = a.vb[b0 + (i - la)];
        }
    }
    __syncthreads();

    // this thread's diagonal inside the tile
    const uint32_t diag = min(tid * MERGE_IPT, tile_n);
    uint32_t lo = diag > lb ? diag - lb : 0, hi = min(diag, la);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t bi = la + (diag - 1 - mid);
        if (!rec_less(sk[bi], (V)(sv[bi] & a.id_mask), sk[mid], (V)(sv[mid] & a.id_mask)))
            lo = mid + 1;
        else
            hi = mid;
    }
    uint32_t ai = lo, bi = la + (diag - lo);
    K rk[MERGE_IPT];
    V rv[MERGE_IPT];
#pragma unroll
    for (int q = 0; q < MERGE_IPT; ++q) {
        const bool has_a = ai < la, has_b = bi < tile_n;
        bool take_a = has_a;
        if (has_a && has_b) take_a = !rec_less(sk[bi], (V)(sv[bi] & a.id_mask), sk[ai], (V)(sv[ai] & a.id_mask));
        if (has_a || has_b) {
            const uint32_t s = take_a ? ai : bi;
            rk[q] = sk[s];
            rv[q] = sv[s];
            if (take_a)
                ++ai;
            else
                ++bi;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MERGE_IPT; ++q) {
        const uint32_t o = tid * MERGE_IPT + q;
        if (o < tile_n) {
            sk[o] = rk[q];
            sv[o] = rv[q];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < tile_n; i += MERGE_THREADS) {
        a.kout[d0 + i] = sk[i];
        a.vout[d0 + i] = sv[i];
    }
}

} // namespace bp
