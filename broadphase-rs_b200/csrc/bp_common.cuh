// bp_common.cuh -- shared device/host helpers for the sm_100a kernels.
//
// Index layouts follow the reference's index_impl! instantiations (src/index.rs:293-295); nothing
// here is translated from the reference -- the codec is the plain "axis bit i -> origin bit
// DIM*i + axis" definition implemented with 64-bit magic-mask spreads.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bp.h"

#define BP_FULL_MASK 0xffffffffu

namespace bp {

// ---------------------------------------------------------------------------------------------
// Index traits -- src/index.rs:65-123, 293-295
// ---------------------------------------------------------------------------------------------
template <int KIND> struct IndexTraits;

template <> struct IndexTraits<BP_INDEX32_2D> {
    typedef uint32_t key_t;
    static constexpr int DIM = 2, DEPTH_BITS = 4, AXIS_BITS = 14;
};
template <> struct IndexTraits<BP_INDEX64_2D> {
    typedef uint64_t key_t;
    static constexpr int DIM = 2, DEPTH_BITS = 5, AXIS_BITS = 29;
};
template <> struct IndexTraits<BP_INDEX64_3D> {
    typedef uint64_t key_t;
    static constexpr int DIM = 3, DEPTH_BITS = 5, AXIS_BITS = 19;
};

// bit i -> bit 2*i (low 32 bits of v)
__host__ __device__ __forceinline__ uint64_t spread2(uint64_t v) {
    v &= 0xffffffffull;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}
// bit i -> bit 3*i (low 21 bits of v)
__host__ __device__ __forceinline__ uint64_t spread3(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// encode_axis (src/index.rs:155-172, 193-207): top AXIS_BITS of the u32 coordinate, spread by DIM.
template <class T> __host__ __device__ __forceinline__ uint64_t encode_axis(uint32_t v) {
    const uint64_t top = (uint64_t)(v >> (32 - T::AXIS_BITS));
    return T::DIM == 2 ? spread2(top) : spread3(top);
}

// Index::default().set_depth(depth).set_origin(origin) -- src/index.rs:106-112, 230-250.
// `origin_bits` = OR over axes of encode_axis(p[axis]) << axis.  depth is already clamped.
template <class T> __host__ __device__ __forceinline__ typename T::key_t make_key(uint32_t depth, uint64_t origin_bits) {
    typedef typename T::key_t K;
    const K origin_mask = (K)(((((uint64_t)1 << (T::DIM * T::AXIS_BITS)) - 1)) << T::DEPTH_BITS);
    return (K)(((K)(origin_bits << T::DEPTH_BITS) & origin_mask) | (K)depth);
}

template <class T> __host__ __device__ __forceinline__ uint32_t key_depth(typename T::key_t k) {
    return (uint32_t)(k & (typename T::key_t)((1u << T::DEPTH_BITS) - 1));
}

// level_mask -- src/index.rs:82-86
template <class T> __host__ __device__ __forceinline__ typename T::key_t level_mask(uint32_t depth) {
    typedef typename T::key_t K;
    if (depth == 0) return (K)0;
    const int total = T::DIM * T::AXIS_BITS + T::DEPTH_BITS;
    return (K)(((((uint64_t)1 << (T::DIM * depth)) - 1)) << (total - T::DIM * depth));
}

// Largest key a descendant-or-equal of `k` can have: every record j > i with key_j <= this value
// lies in cell(i) (the contiguity lemma of DESIGN.md; reference semantics src/layer.rs:550-573).
template <class T> __host__ __device__ __forceinline__ typename T::key_t run_upper_key(typename T::key_t k) {
    typedef typename T::key_t K;
    constexpr int total = T::DIM * T::AXIS_BITS + T::DEPTH_BITS; // bits a key uses (the rest are always 0)
    const K used = total >= (int)(8 * sizeof(K)) ? (K) ~(K)0 : (K)((((uint64_t)1) << (total & 63)) - 1);
    return (K)(k | (~level_mask<T>(key_depth<T>(k)) & used));
}

// ---------------------------------------------------------------------------------------------
// warp / block helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BP_FULL_MASK, v, o);
    return v;
}
template <class T> __device__ __forceinline__ T warp_or(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(BP_FULL_MASK, v, o);
    return v;
}
template <class T> __device__ __forceinline__ T warp_and(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v &= __shfl_xor_sync(BP_FULL_MASK, v, o);
    return v;
}
// inclusive warp scan
template <class T> __device__ __forceinline__ T warp_inclusive_sum(T v) {
    const unsigned lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(BP_FULL_MASK, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// Block-wide exclusive sum over one value per thread.  `warp_totals` is shared scratch of
// (THREADS/32 + 1) entries.  Returns the exclusive prefix; *block_total gets the sum.
// Contains two __syncthreads(); all threads of the block must call it.
template <int THREADS, class T> __device__ __forceinline__ T block_exclusive_sum(T v, T *warp_totals, T *block_total) {
    constexpr int WARPS = THREADS / 32;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    const T incl = warp_inclusive_sum(v);
    if (lane == 31) warp_totals[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T t = (lane < (unsigned)WARPS) ? warp_totals[lane] : (T)0;
        const T ti = warp_inclusive_sum(t);
        if (lane < (unsigned)WARPS) warp_totals[lane] = ti - t; // exclusive warp offsets
        if (lane == 31) warp_totals[WARPS] = ti;                // total
    }
    __syncthreads();
    *block_total = warp_totals[WARPS];
    return warp_totals[warp] + incl - v;
}

// ---------------------------------------------------------------------------------------------
// Decoupled look-back (single-pass chained scan), one 64-bit status word per tile:
//   bits 63..62 = flag (0 empty, 1 aggregate, 2 inclusive prefix), bits 61..0 = value.
// Tiles take their index from an atomic ticket, so every predecessor of a running tile has
// already started: the spin below cannot deadlock.  It is still bounded (BP_SPIN_LIMIT polls) and
// raises *err instead of hanging the GPU if that invariant is ever broken.
// ---------------------------------------------------------------------------------------------
// (2^26 polls, the later ones ~64 ns apart: seconds, so that a predecessor delayed by time-slicing (MPS, a debugger, preemption)
// is waited for rather than reported -- ADVICE round 1; the kernels stay memory-safe on a time-out: offsets only shrink.)
#define BP_SPIN_LIMIT (1u << 26)
constexpr uint64_t LB_FLAG_AGG = 1ull << 62;
constexpr uint64_t LB_FLAG_INC = 2ull << 62;
constexpr uint64_t LB_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) { return *(const volatile uint64_t *)p; }
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) { *(volatile uint64_t *)p = v; }
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return *(const volatile uint32_t *)p; }
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }

// Called by all 32 lanes of ONE warp.  Publishes `aggregate` for `tile`, walks back over the
// predecessors 32 at a time and returns the exclusive prefix (sum of all earlier tiles) to every
// lane; then publishes the inclusive prefix.
__device__ __forceinline__ uint64_t lookback_exclusive(uint64_t *status, uint32_t tile, uint64_t aggregate, int *err) {
    const unsigned lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(&status[0], LB_FLAG_INC | aggregate);
        return 0;
    }
    if (lane == 0) st_volatile_u64(&status[tile], LB_FLAG_AGG | aggregate);
    uint64_t exclusive = 0;
    int64_t base = (int64_t)tile - 1;
    for (;;) {
        const int64_t idx = base - (int64_t)lane;
        uint64_t s = LB_FLAG_INC; // virtual tiles before tile 0: inclusive prefix 0
        if (idx >= 0) {
            s = ld_volatile_u64(&status[idx]);
            uint32_t spins = 0;
            while ((s >> 62) == 0) {
                if (++spins > BP_SPIN_LIMIT) {
                    *err = 1;
                    s = LB_FLAG_INC;
                    break;
                }
                __nanosleep(spins > 4096 ? 64 : 20);
                s = ld_volatile_u64(&status[idx]);
            }
        }
        const unsigned inc_mask = __ballot_sync(BP_FULL_MASK, (s >> 62) == 2);
        const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;
        uint64_t v = ((int)lane <= first_inc) ? (s & LB_VALUE_MASK) : 0;
        exclusive += warp_sum(v);
        if (inc_mask) break;
        base -= 32;
    }
    if (lane == 0) st_volatile_u64(&status[tile], LB_FLAG_INC | ((exclusive + aggregate) & LB_VALUE_MASK));
    return exclusive;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copies (1-D cp.async.bulk, global -> shared, completion on an mbarrier): a whole tile arrives
// without passing through registers, one thread issues it.  Addresses and size must be multiples of 16.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "BP_MBAR_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra BP_MBAR_DONE;\n\t"
                 "bra BP_MBAR_WAIT;\n\t"
                 "BP_MBAR_DONE:\n\t"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// Shard of a key: the number of splitters <= v, over 15 ascending splitters in shared memory (unused ones = ~0: no key
// reaches them).  Four dependent probes instead of fifteen comparisons.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t splitter_rank15(const uint64_t *sspl, uint64_t v) {
    uint32_t lo = (sspl[7] <= v) ? 8u : 0u;
    lo += (sspl[lo + 3] <= v) ? 4u : 0u;
    lo += (sspl[lo + 1] <= v) ? 2u : 0u;
    lo += (sspl[lo] <= v) ? 1u : 0u;
    return lo;
}

// ---------------------------------------------------------------------------------------------
// streaming loads: data that is read exactly once should not displace the look-back state in L1
// ---------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }

} // namespace bp
