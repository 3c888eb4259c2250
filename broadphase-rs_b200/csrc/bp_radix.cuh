// bp_radix.cuh -- LSD radix sort ("onesweep": one histogram read, then one read + one write per
// digit) of key arrays with an optional payload array, for sm_100a.
//
// Replaces the reference's `tree.par_sort_unstable()` (src/layer.rs:149, :162) and
// `collisions.par_sort_unstable()` (src/layer.rs:473, :516).  Both sort a total order, so any
// correct sort reproduces the reference's sequence exactly.
//
// Per pass, each CTA takes a tile (atomic ticket), ranks its keys by digit with warp-level
// match-any multisplit into per-warp shared-memory counters, publishes its per-digit counts, resolves
// its global digit offsets by a decoupled look-back over the preceding tiles, stages the tile in shared
// memory in digit order and writes each digit run to its final place with coalesced stores.
#pragma once

#include <type_traits>

#include "bp_common.cuh"

namespace bp {

struct NoVal {};

constexpr int RADIX_MAX_PASSES = 16;
constexpr int RADIX = 256;      // bins of an 8-bit digit (the default digit width)
constexpr int RADIX_MAX_BITS = 10; // widest digit a pass kernel is instantiated for (record sort: 9 or 10 bits when that saves a pass)

// One digit = up to two bit-fields of the key: (key >> shift) & ((1 << bits) - 1), with the second field
// (shift2, bits2; bits2 may be 0) stacked above the first.  Two fields let the planner skip a gap of
// constant bits (e.g. between the depth tag and the significant Morton bits) inside one pass.
struct RadixPlan {
    int npasses;
    unsigned char shift[RADIX_MAX_PASSES];
    unsigned char bits[RADIX_MAX_PASSES];
    unsigned char shift2[RADIX_MAX_PASSES];
    unsigned char bits2[RADIX_MAX_PASSES];
};

template <class K> __host__ __device__ __forceinline__ uint32_t plan_digit(const RadixPlan &pl, int p, K k) {
    uint32_t d = (uint32_t)(k >> pl.shift[p]) & ((1u << pl.bits[p]) - 1u);
    if (pl.bits2[p]) d |= ((uint32_t)(k >> pl.shift2[p]) & ((1u << pl.bits2[p]) - 1u)) << pl.bits[p];
    return d;
}

// ---- histograms for every planned digit in one read of the keys ---------------------------------
// NB = bins per pass (a power of two >= 2^bits of every planned digit); dynamic shared memory: np * NB counters.
template <class K>
__global__ void __launch_bounds__(512) radix_hist_kernel(const K *__restrict__ keys, uint32_t n_host,
                                                          const uint32_t *__restrict__ n_dev, RadixPlan plan,
                                                          uint32_t *__restrict__ ghist, int NB = RADIX) {
    extern __shared__ uint32_t sh[];
    const int np = plan.npasses;
    for (int i = threadIdx.x; i < np * NB; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t n = n_dev ? *n_dev : n_host;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    constexpr int UNROLL = 8;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // keys in registers, passes in the outer loop: the digit descriptor of a pass is fetched once per
    // UNROLL keys (warp-uniform constant-bank loads)
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        K k[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) k[u] = ld_stream(keys + i + u * stride);
        for (int p = 0; p < np; ++p) {
            const uint32_t s0 = plan.shift[p], m0 = (1u << plan.bits[p]) - 1u, b0 = plan.bits[p];
            const uint32_t s1 = plan.shift2[p], m1 = (1u << plan.bits2[p]) - 1u;
            uint32_t *h = sh + p * NB;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                atomicAdd(&h[((uint32_t)(k[u] >> s0) & m0) | (((uint32_t)(k[u] >> s1) & m1) << b0)], 1u);
        }
    }
    for (; i < n; i += stride) {
        const K k = ld_stream(keys + i);
        for (int p = 0; p < np; ++p) atomicAdd(&sh[p * NB + plan_digit<K>(plan, p, k)], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < np * NB; j += blockDim.x) {
        const uint32_t c = sh[j];
        if (c) atomicAdd(&ghist[j], c);
    }
}

// ---- exclusive scan of each pass's NB-bin histogram (one CTA of NB threads per pass) -----------
template <int NB = RADIX> __global__ void __launch_bounds__(NB) radix_scan_hist_kernel(uint32_t *__restrict__ ghist) {
    __shared__ uint32_t wt[NB / 32 + 1];
    uint32_t *h = ghist + (size_t)blockIdx.x * NB;
    const uint32_t c = h[threadIdx.x];
    uint32_t total;
    const uint32_t ex = block_exclusive_sum<NB, uint32_t>(c, wt, &total);
    h[threadIdx.x] = ex;
}

// ---- digit functors -----------------------------------------------------------------------------------
// The pass kernel is generic over how a key maps to a bin.  Pads (all-ones keys) must map to max().
template <class K> struct ShiftMaskDigit { // LSD radix digit: one or two bit-fields (see RadixPlan)
    static constexpr bool SCATTER = false;
    uint32_t shift, mask, shift2, mask2, bits;
    __device__ __forceinline__ uint32_t operator()(K k) const {
        return ((uint32_t)(k >> shift) & mask) | (((uint32_t)(k >> shift2) & mask2) << bits);
    }
    __device__ __forceinline__ uint32_t max() const { return mask | (mask2 << bits); }
};
template <class K> struct OneFieldDigit { // the common case: one bit-field, (key >> shift) & mask
    static constexpr bool SCATTER = false;
    uint32_t shift, mask;
    __device__ __forceinline__ uint32_t operator()(K k) const { return (uint32_t)(k >> shift) & mask; }
    __device__ __forceinline__ uint32_t max() const { return mask; }
};
constexpr int MAX_SPLITTERS = 15; // up to 16 shards
template <class K> struct SplitterDigit { // range partition: number of splitters <= (key >> shift)
    static constexpr bool SCATTER = false;
    uint64_t spl[MAX_SPLITTERS];
    uint32_t n, shift;
    __device__ __forceinline__ uint32_t operator()(K k) const {
        const uint64_t v = (uint64_t)k >> shift;
        uint32_t d = 0;
        for (uint32_t i = 0; i < n; ++i) d += (spl[i] <= v) ? 1u : 0u; // n is warp-uniform and small
        return d;
    }
    __device__ __forceinline__ uint32_t max() const { return n; }
};
// The same partition, but every bucket has its own destination array -- e.g. the receive buffer of
// another GPU mapped over NVLink (symmetric memory): the partition pass IS the all-to-all.
template <class K> struct SplitterScatterDigit : SplitterDigit<K> {
    static constexpr bool SCATTER = true;
    uint64_t kdst[MAX_SPLITTERS + 1]; // device addresses of the key destination of every bucket
    uint64_t vdst[MAX_SPLITTERS + 1]; // ... and of the payload
};

// ---- one onesweep pass --------------------------------------------------------------------------------
template <class K, class V, class Op = ShiftMaskDigit<K>> struct RadixPassArgs {
    const K *kin;
    K *kout;
    const V *vin;
    V *vout;
    uint32_t n_host;
    const uint32_t *n_dev;      // optional: element count in device memory (grid sized by n_host >= *n_dev)
    const uint32_t *ghist_excl; // [NB] exclusive digit offsets of this pass (NB = 2^RB bins)
    uint32_t *status;           // [tiles][NB], zeroed; bits 31..30 flag, 29..0 count
    uint32_t *tile_counter;     // zeroed
    const uint8_t *vflags;      // optional: 3 flag bits per input element, OR-ed into the top bits of the payload as it is loaded
    Op op;
    int *err;
};

// Lanes of the warp holding the same 8-bit digit.  Eight ballots instead of one MATCH.ANY: on
// sm_100a MATCH runs on the ADU pipe at a rate that drops with the number of distinct values in
// the warp -- the ncu capture profiles/r1_sortpass_before.txt shows it at 56% of peak, the top
// unit, on high-entropy digits -- while VOTE does not depend on the data.
template <int RB = 8> __device__ __forceinline__ unsigned match_digit(uint32_t d) {
    unsigned m = BP_FULL_MASK;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        // m &= (bit b of d) ? ballot : ~ballot, as  m & ~(ballot ^ e)  with e = all-ones iff the bit is set
        asm("{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t, v, e;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
            "selp.b32 e, 0xffffffff, 0, p;\n\t"
            "lop3.b32 %0, %0, v, e, 0x90;\n\t"
            "}"
            : "+r"(m)
            : "r"(d), "r"(1u << b));
    }
    return m;
}

constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_INC = 2u << 30, RS_VALUE_MASK = (1u << 30) - 1;

// DPT consecutive status words (the digits one look-back thread owns) with one volatile vector load / store.  Every word
// carries its own flag, so tearing between the words of a vector is harmless.
template <int DPT> struct StatusVec { uint32_t w[DPT]; };
template <int DPT> __device__ __forceinline__ StatusVec<DPT> ld_status(const uint32_t *p) {
    StatusVec<DPT> v;
    if constexpr (DPT == 1) {
        v.w[0] = ld_volatile_u32(p);
    } else if constexpr (DPT == 2) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.w[0]), "=r"(v.w[1]) : "l"(p) : "memory");
    } else {
        static_assert(DPT == 4, "1, 2 or 4 digits per look-back thread");
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
    }
    return v;
}
template <int DPT> __device__ __forceinline__ void st_status(uint32_t *p, const StatusVec<DPT> &v) {
    if constexpr (DPT == 1) {
        st_volatile_u32(p, v.w[0]);
    } else if constexpr (DPT == 2) {
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]) : "memory");
    } else {
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]) : "memory");
    }
}

// One round of the per-digit look-back: B predecessor tiles t, t-1, .. of the DPT digits starting at `d0`, loaded together.
// NB = status words per tile.
template <int B, int DPT>
__device__ __forceinline__ void lookback_round(const uint32_t *status, int NB, unsigned d0, int64_t &t, uint32_t (&excl)[DPT],
                                               unsigned &done, int *err) {
    StatusVec<DPT> sv[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        if (t - b >= 0) {
            sv[b] = ld_status<DPT>(status + (size_t)(t - b) * NB + d0);
        } else {
#pragma unroll
            for (int j = 0; j < DPT; ++j) sv[b].w[j] = RS_FLAG_INC;
        }
    }
    if constexpr (DPT == 1) { // the 8-bit pass: one digit per thread, leave the batch at the first inclusive prefix
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (done) break;
            uint32_t sb = sv[b].w[0];
            if ((sb >> 30) == 0) { // not published yet: poll this one
                const uint32_t *ps = status + (size_t)(t - b) * NB + d0;
                uint32_t spins = 0;
                do {
                    if (++spins > BP_SPIN_LIMIT) {
                        *err = 1;
                        sb = RS_FLAG_INC;
                        break;
                    }
                    if (spins > 4096) __nanosleep(64);
                    sb = ld_volatile_u32(ps);
                } while ((sb >> 30) == 0);
            }
            excl[0] += sb & RS_VALUE_MASK;
            if ((sb >> 30) == 2) done = 1u;
        }
    } else {
#pragma unroll
        for (int b = 0; b < B; ++b) {
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                uint32_t sb = sv[b].w[j];
                const bool live = !(done & (1u << j));
                if (live && (sb >> 30) == 0) { // not published yet: poll this one
                    const uint32_t *ps = status + (size_t)(t - b) * NB + d0 + j;
                    uint32_t spins = 0;
                    do {
                        if (++spins > BP_SPIN_LIMIT) {
                            *err = 1;
                            sb = RS_FLAG_INC;
                            break;
                        }
                        sb = ld_volatile_u32(ps);
                    } while ((sb >> 30) == 0);
                }
                excl[j] += live ? (sb & RS_VALUE_MASK) : 0u; // branch-free apart from the (rare) poll
                done |= (live && (sb >> 30) == 2) ? (1u << j) : 0u;
            }
        }
    }
    t -= B;
}

// RB = digit width in bits: NB = 2^RB bins.  The OWN = min(NB, 256) threads that own the digits (publish, look-back, digit
// offsets) own DPT = NB / OWN consecutive digits each; above 8 bits the per-warp digit counters are packed 16-bit halves (a
// warp ranks 32 * ITEMS < 2^16 keys), so the shared-memory footprint -- and with it 3 CTAs per SM -- stays that of the 8-bit
// pass.  Below 8 bits (a sort whose passes need not be full: 27 varying bits = 4 x 7) a pass votes once less per key and its
// digit runs are twice as long.
template <class K, class V, int THREADS, int ITEMS, int RB = 8> struct RadixPassCfg {
    static constexpr bool HAS_V = !std::is_same<V, NoVal>::value;
    static constexpr int NB = 1 << RB;
    static constexpr int OWN = NB < 256 ? NB : 256;
    static constexpr int DPT = NB / OWN;
    static constexpr bool PACK = RB > 8;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int WARPS = THREADS / 32;
    static constexpr size_t ELEM = HAS_V ? (sizeof(K) > sizeof(V) ? sizeof(K) : sizeof(V)) : sizeof(K);
    static constexpr size_t STAGE_BYTES = (size_t)TILE * ELEM;
    static constexpr size_t WHIST_WORDS = (size_t)WARPS * NB / (PACK ? 2 : 1);
    static constexpr size_t SMEM_BYTES = STAGE_BYTES + (WHIST_WORDS + 2 * NB + 16) * sizeof(uint32_t);
    static_assert(RB >= 5 && RB <= RADIX_MAX_BITS, "digit width");
    static_assert(THREADS >= 256 && THREADS % 32 == 0, "up to 256 threads own the digits");
    static_assert(TILE < 65536, "ranks are packed in 16 bits");
};

// The body of one tile.  FULL = the tile has exactly TILE elements: every bounds check disappears.
template <class K, class V, class Op, int THREADS, int ITEMS, bool FULL, int RB>
__device__ __forceinline__ void radix_pass_tile(const RadixPassArgs<K, V, Op> &a, unsigned char *smem_raw, const uint32_t tile,
                                                const uint32_t tile_n) {
    typedef RadixPassCfg<K, V, THREADS, ITEMS, RB> Cfg;
    constexpr int TILE = Cfg::TILE, WARPS = Cfg::WARPS, NB = Cfg::NB, DPT = Cfg::DPT, OWN = Cfg::OWN;
    constexpr bool HAS_V = Cfg::HAS_V, PACK = Cfg::PACK;

    unsigned char *stage = smem_raw;
    uint32_t *whist = (uint32_t *)(smem_raw + Cfg::STAGE_BYTES); // [WARPS][NB] counters (u32, or u16 pairs when PACK)
    uint32_t *dstart = whist + Cfg::WHIST_WORDS;                 // [NB] first block rank of each digit
    uint32_t *gbase = dstart + NB;                               // [NB] global offset minus dstart
    uint32_t *misc = gbase + NB;                                 // [0] tile, [1..9] warp totals
    constexpr int WROW = NB / (PACK ? 2 : 1);                    // words per warp row

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t tile_begin = (uint64_t)tile * TILE;

    // ---- load keys, warp-striped: item k of lane l of warp w is tile element w*32*ITEMS + k*32 + l
    const K *kin = a.kin + tile_begin;
    const uint32_t base_i = warp * (32 * ITEMS) + lane;
    K key[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base_i + k * 32;
        if (FULL)
            key[k] = ld_stream(kin + i);
        else
            key[k] = (i < tile_n) ? ld_stream(kin + i) : (K) ~(K)0; // pads rank last within the tile
    }

    // ---- rank within the warp: match groups + one shared counter per (warp, digit) --------------
    // Three software-pipelined sweeps instead of one dependent chain per item: (1) all match
    // votes, (2) one shared-memory atomic per group leader (same-address atomics of a warp retire in
    // program order, so earlier items get the lower ranks: the sort stays stable), (3) the leaders'
    // old counter values are broadcast with shuffles.
    const Op op = a.op;
    const uint32_t dmax = op.max();
    uint32_t *wrow = whist + warp * WROW;
    const unsigned lt = lanemask_lt();
    unsigned mm[ITEMS];
    uint32_t rd[ITEMS]; // low 16 bits: rank, high 16 bits: digit
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = op(key[k]);
        mm[k] = match_digit<RB>(d);
        rd[k] = d;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        uint32_t old = 0;
        if ((mm[k] & lt) == 0) { // lowest lane of its group
            if constexpr (PACK) {
                const uint32_t sh = (rd[k] & 1u) << 4;
                old = (atomicAdd(&wrow[rd[k] >> 1], (uint32_t)__popc(mm[k]) << sh) >> sh) & 0xffffu;
            } else {
                old = atomicAdd(&wrow[rd[k]], (uint32_t)__popc(mm[k]));
            }
        }
        rd[k] |= old << 16; // parked in the high half until the broadcast below
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = rd[k] & 0xffffu;
        const uint32_t old = __shfl_sync(BP_FULL_MASK, rd[k] >> 16, __ffs(mm[k]) - 1);
        rd[k] = (old + __popc(mm[k] & lt)) | (d << 16);
    }

    __syncthreads();

    // ---- per digit: exclusive scan over the warps, publish the tile count ----------------------------
    const unsigned d0 = tid * DPT; // first of this thread's digits (tid < OWN)
    uint32_t count_full[DPT];
    uint32_t incl = 0, mine = 0;
    if (tid < (unsigned)OWN) {
        if constexpr (PACK) { // two digits per word: the packed halves add independently (a tile holds < 2^16 keys)
#pragma unroll
            for (int j = 0; j < DPT / 2; ++j) {
                uint32_t sum = 0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    const uint32_t c = whist[w * WROW + tid * (DPT / 2) + j];
                    whist[w * WROW + tid * (DPT / 2) + j] = sum;
                    sum += c;
                }
                count_full[2 * j] = sum & 0xffffu;
                count_full[2 * j + 1] = sum >> 16;
            }
        } else {
            uint32_t sum = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = whist[w * WROW + tid];
                whist[w * WROW + tid] = sum;
                sum += c;
            }
            count_full[0] = sum;
        }
        StatusVec<DPT> st;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            uint32_t count = count_full[j];
            if (!FULL && d0 + j == dmax) count -= (uint32_t)(TILE - tile_n); // pads carry the largest digit
            st.w[j] = (tile == 0 ? RS_FLAG_INC : RS_FLAG_AGG) | count;
            mine += count_full[j];
        }
        st_status<DPT>(a.status + (size_t)tile * NB + d0, st);
        incl = warp_inclusive_sum(mine);
        if (lane == 31) misc[1 + warp] = incl;
    }
    __syncthreads();
    if (tid < (unsigned)OWN) {
        uint32_t off = incl - mine;
        for (unsigned w = 0; w < warp; ++w) off += misc[1 + w];
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            dstart[d0 + j] = off;
            off += count_full[j];
        }
    }
    __syncthreads();

    // ---- stage the keys in digit order (their registers die here, before the look-back needs its own) ----
    K *skeys = (K *)stage;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = rd[k] >> 16;
        uint32_t wex; // keys of this digit in the warps before this one
        if constexpr (PACK)
            wex = (wrow[d >> 1] >> ((d & 1u) << 4)) & 0xffffu;
        else
            wex = wrow[d];
        const uint32_t r = (rd[k] & 0xffffu) + wex + dstart[d];
        rd[k] = r;
        skeys[r] = key[k];
    }
    // payload loads are issued now so that their latency overlaps the look-back
    V val[HAS_V ? ITEMS : 1];
    if constexpr (HAS_V) {
        const V *vin = a.vin + tile_begin;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = base_i + k * 32;
            if (FULL || i < tile_n) val[k] = ld_stream(vin + i);
        }
        if (a.vflags) { // the first pass of a record sort folds the cell flags into the IDs' spare top bits
            const uint8_t *fl = a.vflags + tile_begin;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t i = base_i + k * 32;
                if (FULL || i < tile_n) val[k] |= (V)fl[i] << (8 * sizeof(V) - 3);
            }
        }
    }

    // ---- look back: global offset of every digit run of this tile ------------------------------------
    if (tid < (unsigned)OWN) {
        uint32_t excl[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) excl[j] = 0;
        if (tile != 0) {
            // Walk back over the predecessors a batch of tiles at a time: the status loads of one batch are
            // independent, so a batch costs one L2 round trip.  The batch size sets how fast the first
            // wave of tiles (none of which has an inclusive prefix to offer yet) resolves: about
            // tile / (2 * LB_BATCH) round trips for tile number `tile`.  In the steady state an inclusive prefix
            // sits a few tiles back, so the first batch is short (the look-back loads were 12 % of the kernel's
            // instructions, profiles/r1_cfg3_sort_pass.txt).
            constexpr int LB_FIRST = 4, LB_BATCH = 16 / DPT;
            int64_t t = (int64_t)tile - 1;
            unsigned done = 0;
            constexpr unsigned ALL = (1u << DPT) - 1u;
            lookback_round<LB_FIRST, DPT>(a.status, NB, d0, t, excl, done, a.err);
            while (done != ALL && t >= 0) lookback_round<LB_BATCH, DPT>(a.status, NB, d0, t, excl, done, a.err);
            StatusVec<DPT> st;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                uint32_t count = count_full[j];
                if (!FULL && d0 + j == dmax) count -= (uint32_t)(TILE - tile_n);
                st.w[j] = RS_FLAG_INC | ((excl[j] + count) & RS_VALUE_MASK);
            }
            st_status<DPT>(a.status + (size_t)tile * NB + d0, st);
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j) gbase[d0 + j] = a.ghist_excl[d0 + j] + excl[j] - dstart[d0 + j];
    }
    __syncthreads();
    uint32_t dst[ITEMS];
    uint32_t dig[Op::SCATTER ? ITEMS : 1];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = k * THREADS + tid;
        if (FULL || i < tile_n) {
            const K kk = skeys[i];
            const uint32_t d = op(kk);
            dst[k] = gbase[d] + i;
            if constexpr (Op::SCATTER) {
                dig[k] = d;
                ((K *)a.op.kdst[d])[dst[k]] = kk; // possibly a store into a peer GPU's memory
            } else {
                a.kout[dst[k]] = kk;
            }
        }
    }
    if constexpr (HAS_V) {
        __syncthreads();
        V *svals = (V *)stage;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = base_i + k * 32;
            if (FULL || i < tile_n) svals[rd[k]] = val[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = k * THREADS + tid;
            if (FULL || i < tile_n) {
                if constexpr (Op::SCATTER)
                    ((V *)a.op.vdst[dig[k]])[dst[k]] = svals[i];
                else
                    a.vout[dst[k]] = svals[i];
            }
        }
    }
}

template <class K, class V, int THREADS, int ITEMS, int MINB = 1, class Op = ShiftMaskDigit<K>, int RB = 8>
__global__ void __launch_bounds__(THREADS, MINB) radix_pass_kernel(const RadixPassArgs<K, V, Op> a) {
    typedef RadixPassCfg<K, V, THREADS, ITEMS, RB> Cfg;
    constexpr int TILE = Cfg::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *whist = (uint32_t *)(smem_raw + Cfg::STAGE_BYTES);
    uint32_t *misc = whist + Cfg::WHIST_WORDS + 2 * Cfg::NB;

    const unsigned tid = threadIdx.x;
    if (tid == 0) misc[0] = atomicAdd(a.tile_counter, 1u);
    for (int i = tid; i < (int)Cfg::WHIST_WORDS; i += THREADS) whist[i] = 0;
    __syncthreads();
    const uint32_t tile = misc[0];
    const uint32_t n = a.n_dev ? *a.n_dev : a.n_host;
    const uint64_t tile_begin = (uint64_t)tile * TILE;
    if (tile_begin >= n) return;
    const uint32_t tile_n = (uint32_t)min((uint64_t)TILE, (uint64_t)n - tile_begin);
    if (tile_n == (uint32_t)TILE)
        radix_pass_tile<K, V, Op, THREADS, ITEMS, true, RB>(a, smem_raw, tile, tile_n);
    else
        radix_pass_tile<K, V, Op, THREADS, ITEMS, false, RB>(a, smem_raw, tile, tile_n);
}

// ---- record finish: the low key bits, ordered group by group ------------------------------------------------
// R records have only ~log2(R) bits of "position" in them: once a stable LSD sort has ordered the records by the TOP
// ~log2(R) - 3 varying key bits, every group of records that agree on those bits is a handful of neighbours, and
// ordering such a group by the remaining low bits is a few comparisons per record -- far cheaper than one more full
// radix pass per 8 low bits (multi-depth scenes: 43-48 varying key bits = 6 passes; 3 passes + this kernel instead).
// The same idea as pair_finish_kernel, without its compaction: the output position of a record is
//      group start + #(records of the group that sort before it),
// "before" = smaller key, or equal key and earlier in the input (the passes before were stable, so among equal keys the
// input order is the order the tie-break wants -- ascending IDs, see Impl::sort_range).
// A tile reads its records plus RFIN_HALO neighbours on either side; a group of more than RFIN_HALO records is "big" --
// every one of its records can tell, whichever tile it is in -- and is copied through unchanged while `*big` is raised:
// the output then still is a stable permutation of the input and the host runs the remaining radix passes over it.
constexpr int RFIN_THREADS = 256;
constexpr int RFIN_IPT = 8;
constexpr int RFIN_TILE = RFIN_THREADS * RFIN_IPT;
constexpr int RFIN_HALO = 256;
constexpr int RFIN_WIN = RFIN_TILE + 2 * RFIN_HALO;   // window of a tile: its records + RFIN_HALO neighbours on either side
constexpr int RFIN_WBITS = 12;                         // bits of a window position
constexpr int RFIN_ROWS = RFIN_WIN / 32;               // words of the group-head bitmap
static_assert(RFIN_HALO == RFIN_THREADS && RFIN_WIN <= (1 << RFIN_WBITS) && RFIN_HALO % 32 == 0, "window layout");

template <class K, class V> struct RecordFinishArgs {
    const K *kin;
    const V *vin;
    K *kout;
    V *vout;
    uint32_t n;
    uint32_t gshift;   // records with equal (key >> gshift) form a group
    unsigned int *big; // raised when a group exceeds RFIN_HALO records
    // the varying key bits below gshift as one or two bit-fields, (key >> shift) & mask, the second stacked above the
    // first's `bits0` bits -- what orders a group (used when they fit 32 - RFIN_WBITS bits)
    uint32_t shift0, mask0, bits0, shift1, mask1;
};

// The plain form of the rule, kept as the executable statement of it (BP_SORT_FINISH_WALK=1 selects it): every record walks
// its group in the staged keys.
template <class K, class V>
__global__ void __launch_bounds__(RFIN_THREADS) record_finish_walk_kernel(const RecordFinishArgs<K, V> a) {
    __shared__ K sk[RFIN_WIN];
    const unsigned tid = threadIdx.x;
    const uint32_t t0 = blockIdx.x * (uint32_t)RFIN_TILE;
    if (t0 >= a.n) return;
    const uint32_t w0 = t0 >= (uint32_t)RFIN_HALO ? t0 - (uint32_t)RFIN_HALO : 0u;
    const uint32_t w1 = (uint32_t)min((uint64_t)a.n, (uint64_t)t0 + RFIN_TILE + RFIN_HALO);
    const uint32_t wn = w1 - w0;
    for (uint32_t i = tid; i < wn; i += RFIN_THREADS) sk[i] = ld_stream(a.kin + w0 + i);
    // the payloads of this thread's records: in flight while the keys settle (striped: coalesced)
    V val[RFIN_IPT];
#pragma unroll
    for (int q = 0; q < RFIN_IPT; ++q) {
        const uint32_t i = t0 + q * RFIN_THREADS + tid;
        if (i < a.n) val[q] = ld_stream(a.vin + i);
    }
    __syncthreads();
    const uint32_t gs = a.gshift;
    bool any_big = false;
#pragma unroll
    for (int q = 0; q < RFIN_IPT; ++q) {
        const uint32_t i = t0 + q * RFIN_THREADS + tid;
        if (i >= a.n) break;
        const uint32_t li = i - w0;
        const K x = sk[li];
        const K g = x >> gs;
        uint32_t h = li, pos = 0;
        while (h > 0 && li - h <= (uint32_t)RFIN_HALO) { // records of the group before this one: smaller or equal keys come first
            const K y = sk[h - 1];
            if ((y >> gs) != g) break;
            --h;
            pos += (y <= x) ? 1u : 0u;
        }
        uint32_t e = li + 1;
        while (e < wn && e - li <= (uint32_t)RFIN_HALO) { // ... and after it: only strictly smaller keys come first
            const K y = sk[e];
            if ((y >> gs) != g) break;
            pos += (y < x) ? 1u : 0u;
            ++e;
        }
        // both ends seen <=> the neighbour on either side is a record of another group (or the array ends there)
        const bool closed_l = h == 0 ? w0 == 0 : (sk[h - 1] >> gs) != g;
        const bool closed_r = e == wn ? w1 == a.n : (sk[e] >> gs) != g;
        const bool big = !closed_l || !closed_r || e - h > (uint32_t)RFIN_HALO;
        any_big |= big;
        const uint32_t out = big ? i : w0 + h + pos;
        a.kout[out] = x;
        a.vout[out] = val[q];
    }
    if (any_big) *a.big = 1u;
}

// The fast form.  No key is compared twice and no group boundary is searched for:
//   * group heads ("my top bits differ from my predecessor's") go into a bitmap of the window, one ballot per 32 records;
//     a record finds the start and the end of its group with a count-leading-zeros / find-first-set on that bitmap;
//   * what orders a group is reduced to ONE word per record, C = the varying key bits below gshift squeezed together
//     (config 3: 4 depth bits + 18 origin bits), so the rank of a record is the number of words of its group before it that
//     are not larger plus the number after it that are smaller: one shared-memory load, one comparison, one add per member.
// C is 32 bits when those bits fit (two bit-fields around their widest gap), else the low gshift bits as they are.
// INNER: every position of the window is a record (all tiles but those at the two ends of the array).
template <class K, class V, class C, bool INNER>
__device__ __forceinline__ void record_finish_tile(const RecordFinishArgs<K, V> &a, C *sc, uint32_t *heads, const uint32_t t0) {
    constexpr int SLOTS = RFIN_WIN / RFIN_THREADS; // window positions per thread: slot j of thread t is position j * THREADS + t
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int32_t n = (int32_t)a.n;
    const int32_t wb = (int32_t)t0 - RFIN_HALO; // index of window position 0 (negative at the start of the array: empty positions)
    const uint32_t gs = a.gshift;
    K kk[SLOTS];
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) { // all the loads first: one memory round trip for the window
        const int32_t g = wb + j * RFIN_THREADS + (int32_t)tid;
        kk[j] = (INNER || (g >= 0 && g < n)) ? ld_stream(a.kin + g) : (K)0;
    }
    V val[RFIN_IPT];
#pragma unroll
    for (int q = 0; q < RFIN_IPT; ++q) {
        const int32_t g = (int32_t)t0 + q * RFIN_THREADS + (int32_t)tid;
        if (INNER || g < n) val[q] = ld_stream(a.vin + g);
    }
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        const uint32_t w = j * RFIN_THREADS + tid;
        const int32_t g = wb + (int32_t)w;
        const K k = kk[j];
        K prev = __shfl_up_sync(BP_FULL_MASK, k, 1);
        if (lane == 0 && (INNER || (g > 0 && g < n))) prev = a.kin[g - 1];
        // a head: the first record, a record whose predecessor belongs to another group, and the position after the last record
        bool head = ((k ^ prev) >> gs) != 0;
        if (!INNER) head = (g >= 0 && g < n) ? (g == 0 || head) : g == n;
        const uint32_t hb = __ballot_sync(BP_FULL_MASK, head);
        if (lane == 0) heads[j * (RFIN_THREADS / 32) + warp] = hb;
        if constexpr (sizeof(C) == 4)
            sc[w] = (C)(((uint32_t)(k >> a.shift0) & a.mask0) | (((uint32_t)(k >> a.shift1) & a.mask1) << a.bits0));
        else
            sc[w] = (C)((uint64_t)k & (gs >= 64 ? ~0ull : (1ull << gs) - 1ull));
    }
    __syncthreads();
    // The members of a group are visited by DISTANCE, d = 1 .. (the largest extent of a group around any of the warp's 32
    // records): the same trip count for every lane (one REDUX), lane l looking at positions w - d and w + d -- consecutive
    // lanes, consecutive words.  (A loop per thread over its own group made the warp run through every lane's loop
    // structure in turn: 164 thread instructions per record, 85 % issue-bound, profiles/r2_finish_kernel.txt.)
    bool any_big = false;
#pragma unroll
    for (int q = 0; q < RFIN_IPT; ++q) {
        const uint32_t w = (q + 1) * RFIN_THREADS + tid;
        const int32_t g = wb + (int32_t)w;
        const bool live = INNER || g < n;
        const uint32_t r = w >> 5, b = w & 31u;
        // start of the group: the last head at or before w
        uint32_t m = heads[r] & (0xffffffffu >> (31u - b));
        int rr = (int)r;
        while (m == 0 && rr > 0 && (int)r - rr <= RFIN_HALO / 32) m = heads[--rr];
        // end of the group: the first head after w
        uint32_t m2 = heads[r] & ((0xfffffffeu << b));
        int re = (int)r;
        while (m2 == 0 && re + 1 < RFIN_ROWS && re - (int)r <= RFIN_HALO / 32) m2 = heads[++re];
        const uint32_t hs = (uint32_t)rr * 32u + 31u - (uint32_t)__clz((int)m);
        const uint32_t he = (uint32_t)re * 32u + (uint32_t)__ffs((int)m2) - 1u;
        const bool big = live && (m == 0 || m2 == 0 || he - hs > (uint32_t)RFIN_HALO);
        const uint32_t nb = (live && !big) ? w - hs : 0u;      // members before this record ...
        const uint32_t na = (live && !big) ? he - 1u - w : 0u; // ... and after it
        const uint32_t dmax = __reduce_max_sync(BP_FULL_MASK, max(nb, na));
        const C mine = sc[w];
        uint32_t cnt = 0;
        for (uint32_t d = 1; d <= dmax; ++d) { // (w - d >= 0 and w + d < RFIN_WIN: d <= RFIN_HALO)
            const C cb = sc[w - d], ca = sc[w + d];
            cnt += (d <= nb && cb <= mine) ? 1u : 0u; // before it: smaller or equal words come first (stable)
            cnt += (d <= na && ca < mine) ? 1u : 0u;  // after it: only smaller ones
        }
        any_big |= big;
        if (live) {
            const int32_t out = big ? g : wb + (int32_t)(hs + cnt);
            a.kout[out] = kk[q + 1];
            a.vout[out] = val[q];
        }
    }
    if (any_big) *a.big = 1u;
}

template <class K, class V, class C>
__global__ void __launch_bounds__(RFIN_THREADS, 4) record_finish_kernel(const RecordFinishArgs<K, V> a) {
    __shared__ C sc[RFIN_WIN];
    __shared__ uint32_t heads[RFIN_ROWS + 1];
    const uint32_t t0 = blockIdx.x * (uint32_t)RFIN_TILE;
    if (t0 >= a.n) return;
    if (t0 >= (uint32_t)RFIN_HALO && (uint64_t)t0 + RFIN_TILE + RFIN_HALO <= (uint64_t)a.n)
        record_finish_tile<K, V, C, true>(a, sc, heads, t0);
    else
        record_finish_tile<K, V, C, false>(a, sc, heads, t0);
}

// ---- bucket counts for a splitter partition (one read of the keys) --------------------------------------
// With HALO (records of a multi-GPU shard exchange): a record whose cell reaches past later splitters is an
// ancestor of records other shards will own, so it is also counted as a halo copy for each of those
// shards: home = #splitters <= key, last = #splitters <= run_upper_key(key), halo copies for home+1..last.
template <class K, class T, bool HALO>
__global__ void __launch_bounds__(512) partition_hist_kernel(const K *__restrict__ keys, uint32_t n, SplitterDigit<K> op,
                                                              uint32_t *__restrict__ ghist, uint32_t *__restrict__ ghalo) {
    __shared__ uint32_t sh[2 * (MAX_SPLITTERS + 1)];
    if (threadIdx.x < 2 * (MAX_SPLITTERS + 1)) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t nb = op.n + 1; // buckets in use
    const unsigned lane = threadIdx.x & 31u;
    uint32_t mine = 0; // lane b of every warp accumulates bucket b
    constexpr int U = 4; // independent key loads in flight per thread
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nround = ((size_t)n + U * stride - 1) / (U * stride) * (U * stride); // whole warps stay in the loop together
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += U * stride) {
        K kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + (size_t)u * stride;
            kk[u] = i < n ? ld_stream(keys + i) : (K)0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool live = i0 + (size_t)u * stride < n;
            const K k = kk[u];
            const uint32_t d = live ? op(k) : 0xffffffffu;
            for (uint32_t b = 0; b < nb; ++b) { // one ballot per bucket in use (2..16), warp-uniform trip count
                const uint32_t c = __popc(__ballot_sync(BP_FULL_MASK, d == b));
                if (lane == b) mine += c;
            }
            if constexpr (HALO) {
                // common case: the cell ends before the next splitter -- one comparison
                if (live && d < op.n) {
                    const K hi = run_upper_key<T>(k);
                    if (((uint64_t)hi >> op.shift) >= op.spl[d]) {
                        const uint32_t last = op(hi);
                        for (uint32_t s = d + 1; s <= last; ++s) atomicAdd(&sh[MAX_SPLITTERS + 1 + s], 1u); // rare
                    }
                }
            }
        }
    }
    if (lane < nb && mine) atomicAdd(&sh[lane], mine);
    __syncthreads();
    if (threadIdx.x <= MAX_SPLITTERS && sh[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], sh[threadIdx.x]);
    if (HALO && threadIdx.x <= MAX_SPLITTERS && sh[MAX_SPLITTERS + 1 + threadIdx.x])
        atomicAdd(&ghalo[threadIdx.x], sh[MAX_SPLITTERS + 1 + threadIdx.x]);
}

// The bucket (and halo) counts of partition_hist_kernel as one row of 64-bit words for a peer-visible count
// matrix: [counts 0..nb) | halo 0..nb) (if any) | tag] -- written on the device so that no host round trip is needed.
constexpr int MAX_ROW_TAGS = 8;
struct RowTags {
    uint64_t v[MAX_ROW_TAGS];
    uint32_t n;
};
constexpr int MAX_ROW_COPIES = 16;
struct RowDst { // the same row goes to every rank's copy of the count matrix (peer stores: no copy engine, no extra launches)
    uint64_t *p[MAX_ROW_COPIES];
    uint32_t n;
};
__global__ void count_row_kernel(const uint32_t *__restrict__ hist, const uint32_t *__restrict__ halo, uint32_t nb, RowTags tags,
                                 RowDst dst) {
    const uint32_t i = threadIdx.x;
    const uint64_t c = i < nb ? hist[i] : 0, h = (halo && i < nb) ? halo[i] : 0;
    for (uint32_t r = 0; r < dst.n; ++r) {
        uint64_t *row = dst.p[r];
        if (i < nb) {
            row[i] = c;
            if (halo) row[nb + i] = h;
        }
        if (i < tags.n) row[(halo ? 2 * nb : nb) + i] = tags.v[i];
    }
}

// Writes the halo copies counted above into their destinations (slots handed out by atomics: the
// receiver sorts its records anyway).
template <class K, class V, class T>
__global__ void __launch_bounds__(256) halo_scatter_kernel(const K *__restrict__ keys, const V *__restrict__ vals, uint32_t n,
                                                            SplitterScatterDigit<K> op, uint32_t *__restrict__ cursor,
                                                            const uint8_t *__restrict__ vflags) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const K k = ld_stream(keys + i);
        const uint32_t d = op(k);
        if (d >= op.n) continue; // already in the last shard
        const K hi = run_upper_key<T>(k);
        if (((uint64_t)hi >> op.shift) < op.spl[d]) continue; // the cell ends before the next splitter
        const uint32_t last = op(hi);
        if (last > d) {
            V v = vals[i];
            if (vflags) v |= (V)vflags[i] << (8 * sizeof(V) - 3);
            for (uint32_t s = d + 1; s <= last; ++s) {
                const uint32_t slot = atomicAdd(&cursor[s], 1u);
                ((K *)op.kdst[s])[slot] = k;
                ((V *)op.vdst[s])[slot] = v;
            }
        }
    }
}

// equal ranges of query keys in a sorted key array (halo look-ups); one thread per query
template <class K>
__global__ void lookup_ranges_kernel(const K *__restrict__ keys, uint32_t n, const K *__restrict__ queries, uint32_t nq,
                                     uint32_t *__restrict__ lo_out, uint32_t *__restrict__ hi_out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const K key = queries[q];
    uint32_t lo = 0, hi = n;
    while (lo < hi) { // lower bound
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (keys[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    const uint32_t first = lo;
    hi = n;
    while (lo < hi) { // upper bound
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (keys[mid] <= key)
            lo = mid + 1;
        else
            hi = mid;
    }
    lo_out[q] = first;
    hi_out[q] = lo;
}

} // namespace bp
