// bp_scan.cuh -- K3/K4: the stack-free, load-balanced reformulation of Layer::scan.
//
// Replaces scan_impl (src/layer.rs:550-573) and the sort + dedup that follows it in scan_filtered /
// par_scan_filtered (src/layer.rs:473-474, 516-517).
//
// The reference sweeps the sorted tree with a stack of "open" ancestor cells.  In the sorted
// (Index, ID) order every cell's descendants-or-equals form one contiguous run that starts at the
// cell's own record, so the stack at record j is exactly the set of earlier *pushed* records i whose
// run contains j.  With
//     anc(i, j)   <=>  i < j  and  key_j <= key_i | ~level_mask(depth_i)
//     inactive(j) <=>  exists i: anc(i, j) and id_i == id_j     (the "same ID already on the stack" skip)
// the reference's output multiset is { (id_j, id_i) : anc(i, j), !inactive(i), !inactive(j),
// filter(id_j, id_i) }.  (DESIGN.md has the proof; tests/test_oracle.py checks the closed form
// against the literal stack sweep.)
//
// Kernels:
//   scan_runs_kernel    4096-record tiles, keys staged by one TMA bulk copy + a 32-key halo.  lcp[p] = whole levels on which
//                       keys p and p+1 agree; the run of record i ends at the first p >= i with lcp[p] < depth_i, found from
//                       per-depth bitmaps (a warp transposes the 32 x 32 bit matrix of `~0 << (lcp + 1)` columns) with two
//                       find-first-set steps, no key comparisons; runs that leave the tile use the halo, very long ones a
//                       warp-cooperative search.  Non-empty runs are compacted, their lengths prefix-summed; a tile takes
//                       its range of (sources, work items) with ONE packed 64-bit atomicAdd (the order of the sources is
//                       irrelevant: the pairs are sorted afterwards).
//   scan_chunks_kernel  every 4096-work-item chunk boundary is claimed by the run that contains it.
//   scan_emit_kernel    one CTA per chunk of 4096 (ancestor, descendant) work items however they are spread over the runs:
//                       an owner table + running maximum gives every item its run, the items are handed out striped
//                       (coalesced descendant loads, broadcast ancestor), the filter is a compile-time functor, with DEDUP
//                       a pair is emitted from the canonical one of the cells two objects share only; survivors are staged
//                       and written coalesced (slots by one atomicAdd per chunk; identity emission writes straight out).
//   count_scan_kernel + pair_scatter_kernel   counting sort of the pairs by later ID while they fit L2 (dense u32 IDs).
//   pair_finish_kernel  after the radix passes over the later ID: orders every (tiny) group of equal later IDs by the
//                       earlier ID, removes duplicates, compacts in order (look-back), writes the final (later, earlier)
//                       layout.  pair_unique_kernel: the adjacent-difference dedup behind the full-width fallback sort.
#pragma once

#include "bp_common.cuh"

namespace bp {

// ---------------------------------------------------------------------------------------------
// filter functors -- scan_filtered's F: FnMut(ID, ID) -> bool (src/layer.rs:456-460)
// ---------------------------------------------------------------------------------------------
struct FilterArgs {
    uint64_t arg;
    const uint32_t *table; // device, n_table x {cat, msk}
    uint64_t n_table;
};

template <int FK> struct FilterFn;
template <> struct FilterFn<BP_FILTER_NONE> {
    template <class IdT> __device__ __forceinline__ static bool pass(const FilterArgs &, IdT, IdT) { return true; }
};
template <> struct FilterFn<BP_FILTER_ID_PARITY> {
    template <class IdT> __device__ __forceinline__ static bool pass(const FilterArgs &, IdT a, IdT b) {
        return ((a ^ b) & (IdT)1) == (IdT)1;
    }
};
template <> struct FilterFn<BP_FILTER_XOR_MASK> {
    template <class IdT> __device__ __forceinline__ static bool pass(const FilterArgs &f, IdT a, IdT b) {
        return (((uint64_t)(a ^ b)) & f.arg) != 0;
    }
};
template <> struct FilterFn<BP_FILTER_CATEGORY> {
    template <class IdT> __device__ __forceinline__ static bool pass(const FilterArgs &f, IdT a, IdT b) {
        uint32_t ca = 0xffffffffu, ma = 0xffffffffu, cb = 0xffffffffu, mb = 0xffffffffu;
        if ((uint64_t)a < f.n_table) {
            const uint2 t = __ldg((const uint2 *)f.table + (uint64_t)a);
            ca = t.x;
            ma = t.y;
        }
        if ((uint64_t)b < f.n_table) {
            const uint2 t = __ldg((const uint2 *)f.table + (uint64_t)b);
            cb = t.x;
            mb = t.y;
        }
        return (ca & mb) != 0 && (cb & ma) != 0;
    }
};

template <> struct FilterFn<BP_FILTER_SPHERES> { // the narrow phase of examples/main.rs:461-479, one IEEE rounding per operation
    template <class IdT> __device__ __forceinline__ static bool pass(const FilterArgs &f, IdT a, IdT b) {
        if ((uint64_t)a >= f.n_table || (uint64_t)b >= f.n_table) return true;
        const float4 sa = __ldg((const float4 *)f.table + (uint64_t)a), sb = __ldg((const float4 *)f.table + (uint64_t)b);
        const float dx = __fsub_rn(sb.x, sa.x), dy = __fsub_rn(sb.y, sa.y), dz = __fsub_rn(sb.z, sa.z); // offset = pos1 - pos0
        const float m = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        return !(__fsqrt_rn(m) > __fadd_rn(sa.w, sb.w)); // `if dist > dist_min { None }`
    }
};

// ---------------------------------------------------------------------------------------------
// scan_runs_kernel
// ---------------------------------------------------------------------------------------------
// Tiles are 4096 elements: with smaller tiles the decoupled look-back is the bottleneck -- so many tiles
// are in flight that the nearest inclusive prefix is hundreds of tiles back (profiles/r1_scan_before.txt:
// 41 % / 31 % of the stall samples of the two kernels were the barrier behind the look-back).
constexpr int RUNS_THREADS = 512;
constexpr int RUNS_IPT = 8;
constexpr int RUNS_TILE = RUNS_THREADS * RUNS_IPT;
constexpr int EMIT_THREADS = 512;
constexpr int EMIT_IPT = 8;
constexpr int EMIT_MINB = 3; // CTAs per SM asked of the emission (latency-bound gathers: long_scoreboard 11 warps per issue slot at 2 CTAs)
constexpr int EMIT_CHUNK = EMIT_THREADS * EMIT_IPT;

struct ScanTotals {
    unsigned long long n_sources;   // records with a non-empty descendant run
    unsigned long long n_work;      // sum of run lengths = (ancestor, descendant) record pairs
    unsigned long long n_raw_pairs; // pairs emitted by scan_emit_kernel
    unsigned long long n_pairs;     // unique pairs
    unsigned int any_same_id;       // an (i, j) work item with id_i == id_j exists -> inactive records exist
    unsigned int pad;
};

template <class T> struct RunsArgs {
    const typename T::key_t *keys; // sorted
    uint32_t n;
    uint32_t *src_idx;   // [n]   compacted: record index of each non-empty run
    uint64_t *src_off;   // [n+1] compacted: exclusive prefix of run lengths (+ total at [n_sources])
    unsigned long long *packed_counter; // zeroed: (sources << RUNS_WORK_BITS) | work, bumped once per tile
    uint32_t *tile_counter; // zeroed; counts finished tiles
    ScanTotals *totals;
    int *err;
};

constexpr int RUNS_WORK_BITS = 34; // 30 bits of source count above 34 bits of work items

// Sources (records with a non-empty descendant run) may be listed in any order as long as src_off is
// the running sum of their lengths in that order: the pairs are sorted afterwards anyway.  So a tile does
// not need the prefix over all EARLIER tiles (a chained scan, whose look-back made every tile wait for
// its slowest predecessor) -- it only needs a private range, which one 64-bit atomicAdd hands out -- and
// inside the tile the sources are listed thread by thread (two warp scans per thread instead of two per item).
//
// Where a run ends.  j > i lies in cell(i) iff key_j shares the top DIM * depth_i origin bits with key_i,
// and in a sorted sequence two keys share a prefix iff every adjacent pair between them does.  With
// lcp[p] = number of whole levels on which keys p and p + 1 agree, the run of i therefore ends at the first
// p >= i with lcp[p] < depth_i ("next smaller value").  The tile answers that without searching the keys:
// one bitmap per depth d (bit p = lcp[p] < d; every lane's column ~0 << (lcp + 1) of 32 records is
// transposed across the warp so that lane d holds the word of row d), a 1-bit-per-word summary above
// it, and a find-next-set-bit per record -- instead of a galloping + bisecting search over 64-bit keys
// (the first version: 224 thread instructions per record, 70 % issue-bound, profiles/r1_scan_runs.txt).
// The tile's keys arrive by one TMA bulk copy.
template <class T> struct RunsSmem {
    typedef typename T::key_t K;
    static constexpr int WORDS = RUNS_TILE / 32;     // bitmap words per depth
    static constexpr int STRIDE = WORDS + 1;         // row stride in words: lane d stores row d, one bank each
    static constexpr int SWORDS = (WORDS + 31) / 32; // summary words per depth
    static constexpr int ROWS = T::AXIS_BITS + 1;    // depths 0 .. AXIS_BITS
    static constexpr int HALO = 32;                  // keys after the tile that are staged with it
    static constexpr size_t KEYS_BYTES = ((size_t)(RUNS_TILE + HALO) * sizeof(K) + 15) & ~(size_t)15;
    static constexpr size_t BITS_OFF = KEYS_BYTES;
    static constexpr size_t SUM_OFF = BITS_OFF + (size_t)ROWS * STRIDE * 4;
    static constexpr size_t MISC_OFF = (SUM_OFF + (size_t)ROWS * SWORDS * 4 + 15) & ~(size_t)15;
    static constexpr size_t BYTES = MISC_OFF + 16 /*mbarrier, sbase*/ + (RUNS_THREADS / 32) * 12;
    static_assert(RUNS_IPT == 8 && SWORDS * 4 == RUNS_THREADS / 32, "one summary byte per warp: 8 bitmap words each");
    static_assert(ROWS <= 32, "lane d holds bitmap row d");
};

// 32 x 32 bit-matrix transpose across a warp: lane r passes row r (bit c = column c) and gets column r
// (bit c = row c's bit r).  Five butterfly steps, each one shuffle + three logic ops: in step k the lanes
// r and r ^ k exchange the off-diagonal k x k blocks (the masked halves contain no bit that a rotation
// could wrap around, so one funnel shift moves them either way).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, unsigned lane) {
#pragma unroll
    for (int s = 4; s >= 0; --s) {
        const uint32_t k = 1u << s;
        const uint32_t lo = s == 4 ? 0x0000ffffu : s == 3 ? 0x00ff00ffu : s == 2 ? 0x0f0f0f0fu : s == 1 ? 0x33333333u : 0x55555555u;
        const bool upper = (lane & k) != 0;
        const uint32_t m = upper ? ~lo : lo;           // the half this lane keeps; the partner's same half moves across
        const uint32_t y = __shfl_xor_sync(BP_FULL_MASK, x, k) & m;
        x = (x & m) | __funnelshift_l(y, y, upper ? 32u - k : k);
    }
    return x;
}

// End of a run that leaves its tile: the whole warp probes [lo, n) 32 positions per step.
// Invariant: keys[lo] <= hi, (end == n or keys[end] > hi).  Returns the last index inside the run.
template <class K> __device__ __noinline__ uint32_t runs_far_search(const K *__restrict__ keys, uint32_t n, uint32_t lo, K hi) {
    const unsigned lane = lane_id();
    uint32_t end = n;
    while (end - lo > 1) {
        const uint32_t span = end - lo - 1;    // unknown positions lo+1 .. end-1
        const uint32_t step = (span + 31) / 32; // >= 1
        const uint64_t pos64 = (uint64_t)lo + (uint64_t)(lane + 1) * step;
        const bool in = pos64 < end;
        const bool gt = in ? (keys[(uint32_t)pos64] > hi) : true;
        const uint32_t b = __ballot_sync(BP_FULL_MASK, gt);
        const int f = b ? __ffs(b) - 1 : 32; // first probe beyond the run (32: all probes are inside)
        if (f < 32) {
            const uint64_t new_end = (uint64_t)lo + (uint64_t)(f + 1) * step;
            end = (uint32_t)(new_end < end ? new_end : end);
        }
        if (f > 0) lo += (uint32_t)f * step;
    }
    return lo;
}

template <class T>
__global__ void __launch_bounds__(RUNS_THREADS, 3) scan_runs_kernel(const RunsArgs<T> a) {
    typedef typename T::key_t K;
    typedef RunsSmem<T> S;
    constexpr int WARPS = RUNS_THREADS / 32;
    constexpr int WSPAN = 32 * RUNS_IPT; // records per warp
    constexpr int STRIDE = S::STRIDE, SWORDS = S::SWORDS, ROWS = S::ROWS;
    constexpr int TOTAL = T::DIM * T::AXIS_BITS + T::DEPTH_BITS;
    constexpr K ORIGIN = (K)((((uint64_t)1 << (T::DIM * T::AXIS_BITS)) - 1) << T::DEPTH_BITS);
    extern __shared__ __align__(16) unsigned char runs_smem[];
    K *skeys = (K *)runs_smem;                               // [tile_n + halo_n]
    uint32_t *sbits = (uint32_t *)(runs_smem + S::BITS_OFF); // [ROWS][STRIDE]
    uint32_t *ssum = (uint32_t *)(runs_smem + S::SUM_OFF);   // [ROWS][SWORDS]: bit w = bitmap word w is not zero
    uint64_t *mbar = (uint64_t *)(runs_smem + S::MISC_OFF);
    unsigned long long *sbase = (unsigned long long *)(mbar + 1);
    unsigned long long *sww = sbase + 1;
    uint32_t *swc = (uint32_t *)(sww + WARPS);

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t r0 = tile * RUNS_TILE;
    if (r0 >= a.n) return;
    const uint32_t tile_n = min((uint32_t)RUNS_TILE, a.n - r0);
    const bool bulk = tile_n == (uint32_t)RUNS_TILE && ((uintptr_t)a.keys & 15u) == 0;
    if (tid == 0) {
        if (bulk) {
            mbar_init(mbar, 1);
            mbar_expect_tx(mbar, (uint32_t)(RUNS_TILE * sizeof(K)));
            bulk_load(skeys, a.keys + r0, (uint32_t)(RUNS_TILE * sizeof(K)), mbar);
        }
    }
    // the first keys after the tile: the successor of the last record, and where most runs that leave the tile end
    const uint32_t halo_n = min((uint32_t)S::HALO, a.n - (r0 + tile_n));
    if (tid < halo_n) skeys[tile_n + tid] = a.keys[r0 + tile_n + tid];
    if (!bulk)
        for (uint32_t i = tid; i < tile_n; i += RUNS_THREADS) skeys[i] = ld_stream(a.keys + r0 + i);
    __syncthreads();
    if (bulk) mbar_wait(mbar, 0);

    // warp-striped: item q of lane l of warp w is tile record w*WSPAN + q*32 + l, i.e. bit l of bitmap word w*IPT + q.
    // dl = depth | (lcp + 1) << 8;  lcp = levels shared with the next record, -1: there is no next record
    // (ends every run), 254: not a record.  The lane's column of the bit matrix "lcp < d" is ~0 << (lcp + 1);
    // the transpose hands lane d the 32 records' bits of row d.
    uint32_t dl[RUNS_IPT];
    uint32_t nz8 = 0; // bit q: this lane's row has a set bit in word q of the warp
#pragma unroll
    for (int q = 0; q < RUNS_IPT; ++q) {
        const uint32_t li = warp * WSPAN + q * 32 + lane;
        uint32_t d = 0, l1 = 255u;
        if (li < tile_n) {
            const K k = skeys[li];
            d = min(key_depth<T>(k), (uint32_t)T::AXIS_BITS);
            l1 = 0;
            if (r0 + li + 1 < a.n) {
                const K x = (K)((k ^ skeys[li + 1]) & ORIGIN);
                l1 = 1u + (x ? (uint32_t)(TOTAL - 64 + __clzll((long long)(uint64_t)x)) / (uint32_t)T::DIM : (uint32_t)T::AXIS_BITS);
            }
        }
        dl[q] = d | (l1 << 8);
        const uint32_t row = warp_transpose32(l1 < 32u ? 0xffffffffu << l1 : 0u, lane);
        if (lane < (unsigned)ROWS) sbits[lane * STRIDE + warp * RUNS_IPT + q] = row;
        nz8 |= (row != 0u ? 1u : 0u) << q;
    }
    if (lane < (unsigned)ROWS) ((unsigned char *)ssum)[lane * (SWORDS * 4) + warp] = (unsigned char)nz8;
    __syncthreads();

    uint32_t len[RUNS_IPT];
    uint32_t far_bits = 0; // bit q: this thread's item q has a run that leaves the tile
#pragma unroll
    for (int q = 0; q < RUNS_IPT; ++q) {
        const uint32_t li = warp * WSPAN + q * 32 + lane;
        len[q] = 0;
        const uint32_t d = dl[q] & 0xffu, l1 = dl[q] >> 8;
        if (l1 != 255u && l1 > d) { // lcp >= d: the next record is still inside this record's cell
            const uint32_t *row = sbits + d * STRIDE;
            uint32_t w = li >> 5;
            uint32_t m = row[w] & (0xfffffffeu << (li & 31u));
            if (m == 0) {
                const uint32_t *srow = ssum + d * SWORDS;
                uint32_t sw = w >> 5;
                uint32_t sm = srow[sw] & ((w & 31u) == 31u ? 0u : (0xfffffffeu << (w & 31u)));
                while (sm == 0 && ++sw < (uint32_t)SWORDS) sm = srow[sw];
                if (sm) {
                    w = (sw << 5) + (uint32_t)__ffs(sm) - 1u;
                    m = row[w];
                }
            }
            if (m)
                len[q] = (w << 5) + (uint32_t)__ffs(m) - 1u - li;
            else
                far_bits |= 1u << q; // every remaining record of the tile is inside; the end is further on
        }
    }
    // Runs that leave the tile (a few per tile: the cells that straddle its end), one at a time by the whole warp.
    // Most of them end within the next few records -- the staged halo answers those without a global load;
    // only cells holding more than HALO records past the tile end take the search over the rest of the tree.
    if (__any_sync(BP_FULL_MASK, far_bits != 0)) {
#pragma unroll
        for (int q = 0; q < RUNS_IPT; ++q) {
            uint32_t m = __ballot_sync(BP_FULL_MASK, (far_bits >> q) & 1u);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t li = warp * WSPAN + q * 32 + src;
                const K hi = run_upper_key<T>(skeys[li]);
                // lane h looks at the h-th record after the tile; past the end of the tree counts as "beyond the run"
                const uint32_t beyond = __ballot_sync(BP_FULL_MASK, lane < halo_n ? skeys[tile_n + lane] > hi : true);
                uint32_t last; // last record inside the run
                if (beyond)
                    last = r0 + tile_n + (uint32_t)__ffs(beyond) - 2u;
                else
                    last = runs_far_search<K>(a.keys, a.n, r0 + tile_n + (uint32_t)S::HALO - 1u, hi);
                if ((int)lane == src) len[q] = last - (r0 + li);
            }
        }
    }

    // list positions: thread by thread (the order of the sources is free), then across warps, then one atomic for the tile
    uint32_t tc = 0;
    unsigned long long ts = 0;
#pragma unroll
    for (int q = 0; q < RUNS_IPT; ++q) {
        tc += len[q] != 0 ? 1u : 0u;
        ts += len[q];
    }
    const uint32_t ci = warp_inclusive_sum(tc);
    const unsigned long long wi = warp_inclusive_sum(ts);
    if (lane == 31) {
        swc[warp] = ci;
        sww[warp] = wi;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t c = lane < WARPS ? swc[lane] : 0;
        unsigned long long w = lane < WARPS ? sww[lane] : 0;
        const uint32_t cinc = warp_inclusive_sum(c);
        const unsigned long long winc = warp_inclusive_sum(w);
        if (lane < WARPS) {
            swc[lane] = cinc - c;
            sww[lane] = winc - w;
        }
        if (lane == 31) {
            unsigned long long add = ((unsigned long long)cinc << RUNS_WORK_BITS) + winc;
            if (winc >= (1ull << (RUNS_WORK_BITS - 1))) { // cannot be packed: the host reports BP_ERR_TOO_LARGE
                a.totals->pad = 1u;
                add = 0;
            }
            const unsigned long long old = atomicAdd(a.packed_counter, add);
            // a carry out of the work field would corrupt the source count: exactly one tile sees it happen
            if ((old & ((1ull << RUNS_WORK_BITS) - 1)) + winc >= (1ull << RUNS_WORK_BITS)) a.totals->pad = 1u;
            *sbase = old;
        }
    }
    __syncthreads();
    uint32_t c = (uint32_t)(*sbase >> RUNS_WORK_BITS) + swc[warp] + ci - tc;
    unsigned long long w = (*sbase & ((1ull << RUNS_WORK_BITS) - 1)) + sww[warp] + wi - ts;
#pragma unroll
    for (int q = 0; q < RUNS_IPT; ++q) {
        if (len[q] != 0) {
            a.src_idx[c] = r0 + warp * WSPAN + q * 32 + lane;
            a.src_off[c] = w;
            ++c;
            w += len[q];
        }
    }
    // the last tile to finish publishes the totals and the sentinel offset
    __syncthreads();
    if (tid == 0) {
        __threadfence(); // the tile's stores (made visible to this thread by the barrier) before the ticket
        const uint32_t done = atomicAdd(a.tile_counter, 1u);
        if (done == gridDim.x - 1) {
            const unsigned long long tot = atomicAdd(a.packed_counter, 0ull);
            const unsigned long long ns = tot >> RUNS_WORK_BITS, nw = tot & ((1ull << RUNS_WORK_BITS) - 1);
            a.totals->n_sources = ns;
            a.totals->n_work = nw;
            a.src_off[ns] = nw;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// scan_chunks_kernel: one thread per non-empty run; every chunk boundary m*EMIT_CHUNK that falls
// inside the run's work range [off, off + len) records this run as the chunk's first source.  The
// ranges tile [0, n_work), so every chunk_src[m] is written exactly once and the emit CTAs need no
// global binary search.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scan_chunks_kernel(const uint64_t *__restrict__ src_off, uint32_t n_sources,
                                                          uint32_t *__restrict__ chunk_src) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_sources) return;
    const uint64_t w = src_off[c], e = src_off[c + 1];
    for (uint64_t m = (w + EMIT_CHUNK - 1) / EMIT_CHUNK; m * EMIT_CHUNK < e; ++m) chunk_src[m] = c;
}

// ---------------------------------------------------------------------------------------------
// scan_emit_kernel
// ---------------------------------------------------------------------------------------------
enum { EMIT_MODE_FIRST = 0, EMIT_MODE_FLAG = 1, EMIT_MODE_ACTIVE = 2 };

template <class IdT> struct EmitArgs {
    const IdT *ids; // sorted tree IDs (with the cell flags in their top 3 bits when id_mask != ~0)
    const void *keys; // sorted tree keys (dedup-at-source only)
    IdT id_mask;      // removes the cell flags from a loaded ID
    const uint32_t *src_idx;
    const uint64_t *src_off;
    const uint32_t *chunk_src;
    uint64_t n_work;
    uint32_t n_sources;
    uint32_t first_owned;   // records below this index only act as ancestors (multi-GPU halo): never the later record
    int mode;               // FIRST: emit assuming no inactive records, report same-ID items;
                            // FLAG: only mark inactive[j]; ACTIVE: emit, skipping inactive records
    unsigned char *inactive; // [n records] (FLAG / ACTIVE)
    uint64_t *out_packed;   // u32 IDs: (later << 32) | earlier
    uint64_t *out_a;        // u64 IDs: later
    uint64_t *out_b;        //          earlier
    uint64_t capacity;      // pairs the output arrays can hold
    unsigned long long *pair_counter; // zeroed: running count of emitted pairs (output slots are handed out by
                                      // atomicAdd: the order of the raw pairs is irrelevant, they are sorted next)
    uint32_t *later_count;  // optional, zeroed, [max ID + 1]: pairs emitted per later ID (counting sort of the pairs)
    ScanTotals *totals;
    FilterArgs filter;
    int *err;
};

template <class IdT> struct EmitSmem {
    static constexpr bool WIDE = sizeof(IdT) == 8;
    static constexpr size_t OFF_WORDS = EMIT_CHUNK + 2;                               // u64: run offsets, later the staged output
    static constexpr size_t IDX_WORDS = WIDE ? EMIT_CHUNK : ((EMIT_CHUNK + 2) / 2 + 1) & ~(size_t)1; // u64: run record indices (u32), later `earlier`; even, so that the owner table stays 16-byte aligned
    static constexpr size_t OWNER_WORDS = EMIT_CHUNK / 4;                             // u64: source of every work item (u16 each)
    static constexpr size_t BYTES = (OFF_WORDS + IDX_WORDS + OWNER_WORDS + EMIT_THREADS / 32 + 4) * sizeof(uint64_t);
    static_assert(EMIT_CHUNK < 65536 && EMIT_IPT == 8, "owner entries are u16; a thread scans 8 of them as one uint4");
};

// Dedup at the source.  Two objects that share several cells meet once per shared cell, so the raw pairs
// carry every ID pair up to 2^DIM times (config 3: 3.4x).  The shared cells of a pair form a box; the pair
// is emitted only from the box's minimum corner.  With fX = "object X also covers the previous cell along
// this axis" (3 bits per record, written by encode_kernel) and s = depth(later) - depth(earlier), the
// later record's cell c is the corner along an axis iff
//      !fLater  ||  (!fEarlier && the s low bits of c's coordinate are zero)
// (c-1 is either not the later object's, or it falls into the previous cell of the earlier object's depth
// which that object does not cover).  Only used while no same-ID (inactive) record has been seen; the
// pair sort + dedup that follows stays in place (IDs owning several bounds still produce duplicates).
template <class T> __device__ __forceinline__ bool canonical_cell(typename T::key_t key_i, typename T::key_t key_j, uint32_t fi, uint32_t fj) {
    if (fj == 0) return true;
    constexpr int TOTAL = T::DIM * T::AXIS_BITS + T::DEPTH_BITS;
    constexpr uint64_t ONE_AXIS = T::DIM == 2 ? 0x5555555555555555ull : 0x9249249249249249ull; // every DIM-th bit
    const uint32_t da = key_depth<T>(key_i), db = key_depth<T>(key_j);
    // the Morton digits of the levels da+1 .. db of the later record's cell, lowest level at bit 0: bit DIM*t + ax is
    // bit t of the cell coordinate along ax (below the earlier record's cell size)
    const uint32_t width = (uint32_t)T::DIM * (db - da);
    const uint64_t digits = ((uint64_t)key_j >> (TOTAL - (int)(T::DIM * db))) & (((uint64_t)1 << width) - 1u);
    uint32_t nonzero = 0; // bit ax: the coordinate along ax has a non-zero bit among those levels
#pragma unroll
    for (int ax = 0; ax < T::DIM; ++ax) nonzero |= ((digits & (ONE_AXIS << ax)) != 0 ? 1u : 0u) << ax;
    return ((fi | nonzero) & fj) == 0;
}

template <class IdT, int FK, class T = IndexTraits<BP_INDEX64_3D>, bool DEDUP = false>
__global__ void __launch_bounds__(EMIT_THREADS, EMIT_MINB) scan_emit_kernel(const EmitArgs<IdT> a) {
    constexpr bool WIDE = sizeof(IdT) == 8;
    constexpr int FLAG_SHIFT = 8 * sizeof(IdT) - 3;
    typedef EmitSmem<IdT> S;
    extern __shared__ __align__(16) unsigned char emit_smem[];
    uint64_t *soff = (uint64_t *)emit_smem;
    // sidx holds the source record indices during the walk; for u64 IDs it is sized so that it can
    // stage the `earlier` half of the output afterwards
    uint64_t *sidx_raw = soff + S::OFF_WORDS;
    uint16_t *sowner = (uint16_t *)(sidx_raw + S::IDX_WORDS); // [EMIT_CHUNK] run (index into soff / sidx, + 1) of every work item
    uint64_t *sscratch = sidx_raw + S::IDX_WORDS + S::OWNER_WORDS;
    uint64_t *sbase_p = sscratch + EMIT_THREADS / 32 + 2;
    uint32_t *sidx = (uint32_t *)sidx_raw;
    // the staged output reuses soff / sidx: every thread has finished its walk before the barrier that
    // precedes the staging
    uint64_t *spa = soff;     // packed pair, or `later`
    uint64_t *spb = sidx_raw; // `earlier` (u64 IDs only)

    const unsigned tid = threadIdx.x;
    const uint32_t chunk = blockIdx.x;
    const uint64_t w0 = (uint64_t)chunk * EMIT_CHUNK;
    if (w0 >= a.n_work) return;
    const uint32_t chunk_n = (uint32_t)min((uint64_t)EMIT_CHUNK, a.n_work - w0);

    // the sources whose runs intersect this chunk: at most chunk_n of them (every run is non-empty)
    const uint32_t c0 = a.chunk_src[chunk];
    const uint32_t nsrc = min(a.n_sources - c0, chunk_n);
    for (uint32_t i = tid; i < nsrc + 1; i += EMIT_THREADS) soff[i] = a.src_off[c0 + i]; // src_off[n_sources] = n_work
    for (uint32_t i = tid; i < nsrc; i += EMIT_THREADS) sidx[i] = a.src_idx[c0 + i];
    __syncthreads();

    // Without a filter and while no same-ID item has been seen, every work item yields exactly one pair:
    // the output position is the work-item index and no compaction (block scan + look-back) is needed.
    // If a same-ID item does turn up, the host discards this emission and re-emits in ACTIVE mode.
    const bool identity = !DEDUP && FK == BP_FILTER_NONE && a.mode == EMIT_MODE_FIRST && a.first_owned == 0;

    // Which run owns which work item: every run stamps its number on its first work item of the chunk, a
    // running maximum spreads it over the rest (thread t scans its 8 consecutive entries as one 16-byte
    // word, then warps, then the block).  After that the work items are handed out STRIPED -- item q of thread
    // t is w0 + q * THREADS + t -- so that the lanes of a warp read consecutive records: the descendant's ID
    // and key are coalesced loads, the ancestor's a broadcast.  (With 8 consecutive items per thread every
    // load instruction touched 32 different sectors; profiles/r1_emit_blocked.txt: 65 % of the LSU wavefront
    // peak, long-scoreboard stalls on top.)
    for (uint32_t i = tid; i < EMIT_CHUNK / 8; i += EMIT_THREADS) ((uint4 *)sowner)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (uint32_t i = tid; i < nsrc; i += EMIT_THREADS) {
        const uint64_t o = soff[i];
        if (o < w0 + chunk_n) sowner[o > w0 ? (uint32_t)(o - w0) : 0u] = (uint16_t)(i + 1); // run starts are distinct
    }
    __syncthreads();
    {
        uint4 v = ((const uint4 *)sowner)[tid];
        uint32_t e[8] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16, v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16};
#pragma unroll
        for (int k = 1; k < 8; ++k) e[k] = max(e[k], e[k - 1]);
        uint32_t incl = e[7];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(BP_FULL_MASK, incl, o);
            if ((tid & 31u) >= (unsigned)o) incl = max(incl, t);
        }
        uint32_t *swmax = (uint32_t *)sscratch;
        if ((tid & 31u) == 31u) swmax[tid >> 5] = incl;
        __syncthreads();
        uint32_t before = __shfl_up_sync(BP_FULL_MASK, incl, 1); // maximum over the earlier threads of this warp ...
        if ((tid & 31u) == 0u) before = 0;
        for (unsigned w = 0; w < (tid >> 5); ++w) before = max(before, swmax[w]); // ... and over the earlier warps
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = max(e[k], before);
        v.x = e[0] | (e[1] << 16);
        v.y = e[2] | (e[3] << 16);
        v.z = e[4] | (e[5] << 16);
        v.w = e[6] | (e[7] << 16);
        ((uint4 *)sowner)[tid] = v;
    }
    __syncthreads();

    uint32_t emit_bits = 0; // bit q: item q of this thread yields a pair (pa[q], pb[q])
    uint64_t pa[EMIT_IPT], pb[EMIT_IPT];
    {
        typedef typename T::key_t K;
        constexpr int BATCH = 4; // items whose loads are in flight together: one memory round trip per batch, not per item
        bool same_seen = false;
#pragma unroll
        for (int h = 0; h < EMIT_IPT; h += BATCH) {
            uint32_t ii[BATCH], jj[BATCH];
            IdT ri[BATCH], rj[BATCH];
            K ki[BATCH], kj[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const uint32_t w = (h + u) * EMIT_THREADS + tid;
                const uint32_t wc = w < chunk_n ? w : 0u; // (item 0 exists: the loads below stay in bounds)
                const uint32_t s = (uint32_t)sowner[wc] - 1u;
                ii[u] = sidx[s];
                jj[u] = ii[u] + 1u + (uint32_t)(w0 + wc - soff[s]);
                ri[u] = a.ids[ii[u]];
                rj[u] = a.ids[jj[u]];
                if constexpr (DEDUP) {
                    const K *keys = (const K *)a.keys;
                    ki[u] = keys[ii[u]];
                    kj[u] = keys[jj[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int q = h + u;
                const uint32_t w = q * EMIT_THREADS + tid;
                if (w >= chunk_n) continue;
                const uint32_t i = ii[u], j = jj[u];
                const IdT id_i = ri[u] & a.id_mask, id_j = rj[u] & a.id_mask;
                bool emit;
                if (a.mode == EMIT_MODE_FIRST) {
                    const bool same = id_i == id_j;
                    same_seen |= same;
                    emit = identity || (!same && j >= a.first_owned && FilterFn<FK>::pass(a.filter, id_j, id_i));
                    if constexpr (DEDUP) {
                        if (emit) emit = canonical_cell<T>(ki[u], kj[u], (uint32_t)(ri[u] >> FLAG_SHIFT), (uint32_t)(rj[u] >> FLAG_SHIFT));
                    }
                } else if (a.mode == EMIT_MODE_FLAG) {
                    if (id_i == id_j) a.inactive[j] = 1;
                    emit = false;
                } else {
                    emit = j >= a.first_owned && !a.inactive[i] && !a.inactive[j] && FilterFn<FK>::pass(a.filter, id_j, id_i);
                }
                if (identity) { // one pair per work item, at the work item's own index: written straight out, coalesced
                    const uint64_t g = w0 + w;
                    if (g < a.capacity) {
                        if (WIDE) {
                            a.out_a[g] = (uint64_t)id_j;
                            a.out_b[g] = (uint64_t)id_i;
                        } else {
                            a.out_packed[g] = ((uint64_t)id_j << 32) | (uint64_t)id_i;
                        }
                    }
                    if (a.later_count) atomicAdd(a.later_count + (size_t)id_j, 1u);
                } else if (emit) {
                    pa[q] = (uint64_t)id_j; // (later, earlier) -- src/layer.rs:567
                    pb[q] = (uint64_t)id_i;
                    emit_bits |= 1u << q;
                    if (a.later_count) atomicAdd(a.later_count + (size_t)id_j, 1u);
                }
            }
        }
        if (same_seen) a.totals->any_same_id = 1u;
    }
    if (identity) {
        if (w0 + chunk_n == a.n_work && tid == 0) *a.pair_counter = a.n_work; // identity: one pair per work item
        return;
    }
    if (a.mode == EMIT_MODE_FLAG) return;

    __syncthreads(); // every thread is done with soff / sidx before the staged output overwrites them
    uint32_t chunk_pass;
    const uint32_t ex = block_exclusive_sum<EMIT_THREADS, uint32_t>((uint32_t)__popc(emit_bits), (uint32_t *)sscratch, &chunk_pass);
    if (tid == 0) *sbase_p = atomicAdd(a.pair_counter, (unsigned long long)chunk_pass);
    // stage, then write coalesced
#pragma unroll
    for (int q = 0; q < EMIT_IPT; ++q) {
        if (emit_bits & (1u << q)) {
            const uint32_t slot = ex + (uint32_t)__popc(emit_bits & ((1u << q) - 1u));
            if (WIDE) {
                spa[slot] = pa[q];
                spb[slot] = pb[q];
            } else {
                spa[slot] = (pa[q] << 32) | pb[q];
            }
        }
    }
    __syncthreads();
    const uint64_t base = *sbase_p;
    for (uint32_t i = tid; i < chunk_pass; i += EMIT_THREADS) {
        const uint64_t g = base + i;
        if (g < a.capacity) {
            if (WIDE) {
                a.out_a[g] = spa[i];
                a.out_b[g] = spb[i];
            } else {
                a.out_packed[g] = spa[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// scan_groups_kernel -- the whole sweep of scan_impl (src/layer.rs:550-573) in ONE kernel for trees whose records all
// sit at the same depth (uniform object sizes: BASELINE configs 2, 4, 5 -- the host sees it in the varying-bit mask of
// the keys: no depth bit differs).  Then cell(i) contains cell(j) iff key_i == key_j, so the stack at record j holds
// exactly the earlier records with j's key: the run of equal keys that ends at j.  No run search, no list of sources, no
// work-item chunks, no gathers: a tile stages the IDs of its records (+ GRP_HALO records before it), marks the first
// record of every equal-key group in a bitmap (one ballot per 32 records), and every record pairs itself with the
// records between the head of its group and itself -- counted first (one atomicAdd per tile hands out the output range),
// then written.  Same filter functors, same dedup at the source (equal depths: the pair is canonical iff the two records'
// cell flags share no axis), same-ID items only detected (the host then takes the general path with its inactive-record
// passes), records below first_owned only act as ancestors.  A group of more than GRP_HALO records (a crowded cell)
// raises `pad`: general path as well.  Output past `capacity` is counted, not written: the host grows the buffer and
// runs the kernel again (it sizes the buffer from the last scan, so a steady frame loop never does).
// ---------------------------------------------------------------------------------------------
constexpr int GRP_THREADS = 256;
constexpr int GRP_IPT = 8;
constexpr int GRP_TILE = GRP_THREADS * GRP_IPT;
constexpr int GRP_HALO = GRP_THREADS;            // one more window slot per thread
constexpr int GRP_WIN = GRP_TILE + GRP_HALO;
constexpr int GRP_ROWS = GRP_WIN / 32;
constexpr int GRP_STAGE = 2048;                  // pairs a tile stages for coalesced stores (more than that: written directly)

template <class K, class IdT> struct GroupScanArgs {
    const K *keys;  // sorted tree keys, all of one depth
    const IdT *ids; // sorted tree IDs (cell flags in their top 3 bits when id_mask != ~0)
    uint32_t n;
    IdT id_mask;
    uint32_t first_owned;
    uint64_t *out_packed; // u32 IDs: (later << 32) | earlier
    uint64_t *out_a;      // u64 IDs: later
    uint64_t *out_b;      //          earlier
    uint64_t capacity;
    unsigned long long *pair_counter; // zeroed: pairs emitted
    unsigned long long *work_counter; // zeroed: (ancestor, descendant) record pairs visited
    uint32_t *later_count;            // optional, zeroed: pairs per later ID (counting sort of the pairs)
    ScanTotals *totals;               // any_same_id, pad
    FilterArgs filter;
};

// INNER: every position of the window is a record (all tiles but the first and the last few).
// The walk over a record's ancestors is a loop over the DISTANCE d = 1 .. (largest ancestor count in the warp), the same
// trip count for all 32 lanes (one REDUX), lane l looking at position p - d: consecutive lanes, consecutive shared-memory
// words.  A loop per thread over its own ancestors made the warp execute every lane's loop structure one after the
// other (7.2 warp instructions per record, 76 % issue-bound; profiles/r2_scan_groups.txt).
template <class K, class IdT, int FK, bool DEDUP, bool INNER>
__device__ __forceinline__ void scan_groups_tile(const GroupScanArgs<K, IdT> &a, IdT *sid, uint64_t *stage, uint32_t *heads,
                                                 uint32_t *sscratch, unsigned long long *sbase, uint32_t *swork, const uint32_t t0) {
    constexpr bool WIDE = sizeof(IdT) == 8;
    constexpr int FLAG_SHIFT = 8 * sizeof(IdT) - 3;
    constexpr int SLOTS = GRP_WIN / GRP_THREADS;
    constexpr int STAGE_N = WIDE ? GRP_STAGE / 2 : GRP_STAGE; // (u64 IDs stage two words per pair)
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const int32_t n = (int32_t)a.n;
    const int32_t wb = (int32_t)t0 - GRP_HALO; // record index of window position 0
    if (tid == 0) *swork = 0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        const uint32_t p = j * GRP_THREADS + tid;
        const int32_t g = wb + (int32_t)p;
        const bool valid = INNER || (g >= 0 && g < n);
        const K k = valid ? ld_stream(a.keys + g) : (K)0;
        K prev = __shfl_up_sync(BP_FULL_MASK, k, 1);
        if (lane == 0 && valid && (INNER || g > 0)) prev = a.keys[g - 1];
        const bool head = valid && ((!INNER && g == 0) || k != prev);
        const uint32_t hb = __ballot_sync(BP_FULL_MASK, head);
        if (lane == 0) heads[j * (GRP_THREADS / 32) + warp] = hb;
        sid[p] = valid ? ld_stream(a.ids + g) : (IdT)0;
    }
    __syncthreads();

    // ---- count: every record's ancestors (the records between the head of its group and itself), the verdict on each
    // kept as one bit (up to 32 of them: always, in practice; beyond that they are evaluated again when written) ------
    uint32_t cnt[GRP_IPT], em[GRP_IPT];
    uint32_t mine = 0, work = 0;
    bool big = false, same_seen = false;
#pragma unroll
    for (int q = 0; q < GRP_IPT; ++q) {
        const uint32_t p = (q + 1) * GRP_THREADS + tid;
        const int32_t g = wb + (int32_t)p;
        const uint32_t r = p >> 5, b = p & 31u;
        uint32_t m = heads[r] & (0xffffffffu >> (31u - b));
        int rr = (int)r;
        while (m == 0 && rr > 0 && (int)r - rr <= GRP_HALO / 32) m = heads[--rr];
        uint32_t c = p - ((uint32_t)rr * 32u + 31u - (uint32_t)__clz((int)m));
        if (!INNER && g >= n) {
            c = 0;
        } else if (m == 0 || c > (uint32_t)GRP_HALO) { // the group starts before the window: not for this kernel
            big = true;
            c = 0;
        }
        cnt[q] = c;
        work += c;
        const IdT rj = sid[p];
        const IdT id_j = rj & a.id_mask;
        const uint32_t fj = (uint32_t)(rj >> FLAG_SHIFT);
        const bool owned = (uint32_t)g >= a.first_owned;
        const uint32_t dmax = __reduce_max_sync(BP_FULL_MASK, c);
        uint32_t bits = 0;
        for (uint32_t d = 1; d <= dmax; ++d) {
            const IdT ri = sid[p - d]; // (d <= dmax <= GRP_HALO <= p: inside the window for every lane)
            const IdT id_i = ri & a.id_mask;
            const bool live = d <= c;
            const bool same = id_i == id_j;
            same_seen |= live && same;
            bool e = live && !same && owned && FilterFn<FK>::pass(a.filter, id_j, id_i);
            if (DEDUP) e = e && ((uint32_t)(ri >> FLAG_SHIFT) & fj) == 0u;
            if (d <= 32u)
                bits |= (e ? 1u : 0u) << (d - 1u);
            else
                mine += e ? 1u : 0u;
        }
        em[q] = bits;
        mine += (uint32_t)__popc(bits);
    }
    if (same_seen) a.totals->any_same_id = 1u;
    if (big) a.totals->pad = 1u;
    work = warp_sum(work);
    if (lane == 0 && work) atomicAdd(swork, work);
    uint32_t tile_pairs;
    const uint32_t ex = block_exclusive_sum<GRP_THREADS, uint32_t>(mine, sscratch, &tile_pairs);
    if (tid == 0) {
        *sbase = tile_pairs ? atomicAdd(a.pair_counter, (unsigned long long)tile_pairs) : 0ull;
        if (*swork) atomicAdd(a.work_counter, (unsigned long long)*swork);
    }
    __syncthreads();
    if (tile_pairs == 0) return;

    // ---- write: every surviving pair into this thread's part of the tile's output range, through a staging buffer
    // (coalesced stores) whenever the tile's pairs fit it -----------------------------------------------------------
    const bool staged = tile_pairs <= (uint32_t)STAGE_N;
    const uint64_t base = *sbase;
    uint32_t slot = ex;
    auto put = [&](IdT id_j, IdT id_i) {
        if (staged) {
            if (WIDE) {
                stage[slot] = (uint64_t)id_j;
                stage[STAGE_N + slot] = (uint64_t)id_i;
            } else {
                stage[slot] = ((uint64_t)id_j << 32) | (uint64_t)id_i;
            }
        } else if (base + slot < a.capacity) {
            if (WIDE) {
                a.out_a[base + slot] = (uint64_t)id_j; // (later, earlier) -- src/layer.rs:567
                a.out_b[base + slot] = (uint64_t)id_i;
            } else {
                a.out_packed[base + slot] = ((uint64_t)id_j << 32) | (uint64_t)id_i;
            }
        }
        ++slot;
        if (a.later_count) atomicAdd(a.later_count + (size_t)id_j, 1u);
    };
#pragma unroll
    for (int q = 0; q < GRP_IPT; ++q) {
        const uint32_t p = (q + 1) * GRP_THREADS + tid;
        if (cnt[q] == 0) continue;
        const IdT rj = sid[p];
        const IdT id_j = rj & a.id_mask;
        uint32_t bits = em[q];
        while (bits) {
            const uint32_t d = (uint32_t)__ffs((int)bits);
            bits &= bits - 1u;
            put(id_j, sid[p - d] & a.id_mask);
        }
        if (cnt[q] > 32u) { // (a crowded cell of up to GRP_HALO records)
            const bool owned = (uint32_t)(wb + (int32_t)p) >= a.first_owned;
            const uint32_t fj = (uint32_t)(rj >> FLAG_SHIFT);
#pragma unroll 1
            for (uint32_t d = 33; d <= cnt[q]; ++d) {
                const IdT ri = sid[p - d];
                const IdT id_i = ri & a.id_mask;
                bool e = id_i != id_j && owned && FilterFn<FK>::pass(a.filter, id_j, id_i);
                if (DEDUP) e = e && ((uint32_t)(ri >> FLAG_SHIFT) & fj) == 0u;
                if (e) put(id_j, id_i);
            }
        }
    }
    if (staged) {
        __syncthreads();
        for (uint32_t i = tid; i < tile_pairs; i += GRP_THREADS) {
            if (base + i >= a.capacity) break;
            if (WIDE) {
                a.out_a[base + i] = stage[i];
                a.out_b[base + i] = stage[STAGE_N + i];
            } else {
                a.out_packed[base + i] = stage[i];
            }
        }
    }
}

template <class K, class IdT, int FK, bool DEDUP>
__global__ void __launch_bounds__(GRP_THREADS) scan_groups_kernel(const GroupScanArgs<K, IdT> a) {
    __shared__ IdT sid[GRP_WIN];
    __shared__ uint64_t stage[GRP_STAGE];
    __shared__ uint32_t heads[GRP_ROWS];
    __shared__ uint32_t sscratch[GRP_THREADS / 32 + 2];
    __shared__ unsigned long long sbase;
    __shared__ uint32_t swork;
    const uint32_t t0 = blockIdx.x * (uint32_t)GRP_TILE;
    if (t0 >= a.n) return;
    if (t0 >= (uint32_t)GRP_HALO + 1u && (uint64_t)t0 + GRP_TILE <= (uint64_t)a.n)
        scan_groups_tile<K, IdT, FK, DEDUP, true>(a, sid, stage, heads, sscratch, &sbase, &swork, t0);
    else
        scan_groups_tile<K, IdT, FK, DEDUP, false>(a, sid, stage, heads, sscratch, &sbase, &swork, t0);
}

// ---------------------------------------------------------------------------------------------
// Counting sort of the raw pairs by their later ID -- the first half of `collisions.sort_unstable()`
// (src/layer.rs:473, :516) when the IDs are dense 32-bit numbers (the usual 0..N): scan_emit_kernel
// counts the pairs of every later ID as it emits them, one exclusive scan over the ID range turns the
// counts into group offsets, and one pass drops every pair into its group (slots by atomicAdd: the order
// inside a group is settled by pair_finish_kernel anyway).  One read + one write of the pairs instead of
// one histogram read + three radix passes; the two atomics per pair hit an array that lives in L2.
// ---------------------------------------------------------------------------------------------
constexpr int CSCAN_THREADS = 256;
constexpr int CSCAN_IPT = 16;
constexpr int CSCAN_TILE = CSCAN_THREADS * CSCAN_IPT;

// in-place exclusive prefix sum of counts[0, n) (n a multiple of 4, array 16-byte aligned)
__global__ void __launch_bounds__(CSCAN_THREADS) count_scan_kernel(uint32_t *__restrict__ counts, uint32_t n, uint64_t *status,
                                                                    uint32_t *tile_counter, int *err) {
    __shared__ uint32_t swt[CSCAN_THREADS / 32 + 1];
    __shared__ uint32_t stile;
    __shared__ uint32_t sexcl;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) stile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = stile;
    const uint32_t t0 = tile * CSCAN_TILE;
    if (t0 >= n) return;
    uint4 v[CSCAN_IPT / 4];
    uint32_t sum = 0;
    uint4 *p4 = (uint4 *)(counts + t0) + tid * (CSCAN_IPT / 4); // thread t owns CSCAN_IPT consecutive counts
#pragma unroll
    for (int k = 0; k < CSCAN_IPT / 4; ++k) {
        const uint32_t e = t0 + tid * CSCAN_IPT + k * 4;
        v[k] = e < n ? p4[k] : make_uint4(0, 0, 0, 0);
        sum += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    uint32_t tile_total;
    uint32_t ex = block_exclusive_sum<CSCAN_THREADS, uint32_t>(sum, swt, &tile_total);
    if (warp == 0) {
        const uint64_t e = lookback_exclusive(status, tile, (uint64_t)tile_total, err);
        if (lane == 0) sexcl = (uint32_t)e;
    }
    __syncthreads();
    ex += sexcl;
#pragma unroll
    for (int k = 0; k < CSCAN_IPT / 4; ++k) {
        const uint32_t e = t0 + tid * CSCAN_IPT + k * 4;
        uint4 o;
        o.x = ex;
        o.y = o.x + v[k].x;
        o.z = o.y + v[k].y;
        o.w = o.z + v[k].z;
        ex = o.w + v[k].w;
        if (e < n) p4[k] = o;
    }
}

// every pair into the group of its later ID; cursor[] holds the group offsets and is consumed
__global__ void __launch_bounds__(256) pair_scatter_kernel(const uint64_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ cursor,
                                                            uint64_t *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t p = ld_stream(in + i);
        const uint32_t slot = atomicAdd(cursor + (size_t)(p >> 32), 1u);
        out[slot] = p;
    }
}

// ---------------------------------------------------------------------------------------------
// pair_unique_kernel -- `collisions.dedup()` (src/layer.rs:474, :517) + the (ID, ID) memory layout
// ---------------------------------------------------------------------------------------------
constexpr int UNIQ_THREADS = 256;
constexpr int UNIQ_IPT = 8;
constexpr int UNIQ_TILE = UNIQ_THREADS * UNIQ_IPT;

template <class IdT> struct UniqueArgs {
    const uint64_t *in_packed; // u32 IDs: sorted packed pairs
    const uint64_t *in_a;      // u64 IDs: sorted (a, b) as two arrays
    const uint64_t *in_b;
    uint32_t n_host;
    const unsigned long long *n_dev; // optional device-side count (n_raw_pairs)
    IdT *out;                        // [2 * n] (later, earlier) interleaved
    uint64_t *status;                // zeroed
    uint32_t *tile_counter;          // zeroed
    ScanTotals *totals;
    int *err;
};

template <class IdT>
__global__ void __launch_bounds__(UNIQ_THREADS) pair_unique_kernel(const UniqueArgs<IdT> a) {
    constexpr bool WIDE = sizeof(IdT) == 8;
    constexpr int WARPS = UNIQ_THREADS / 32;
    __shared__ uint32_t swtot[WARPS + 1];
    __shared__ uint64_t sbase;
    __shared__ uint32_t stile;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) stile = atomicAdd(a.tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = stile;
    const uint64_t n = a.n_dev ? (uint64_t)*a.n_dev : (uint64_t)a.n_host;
    const uint64_t t0 = (uint64_t)tile * UNIQ_TILE;
    if (n == 0) {
        if (tile == 0 && tid == 0) a.totals->n_pairs = 0;
        return;
    }
    if (t0 >= n) return;
    const uint32_t tile_n = (uint32_t)min((uint64_t)UNIQ_TILE, n - t0);

    // warp-striped: warp w owns tile elements [w*32*IPT, (w+1)*32*IPT); item q of lane l is element
    // w*32*IPT + q*32 + l, so loads are coalesced and the order inside a warp is (q, lane)
    const uint32_t wseg = warp * (32 * UNIQ_IPT);
    const unsigned lt = lanemask_lt();
    uint64_t xa[UNIQ_IPT], xb[UNIQ_IPT];
    uint32_t pos[UNIQ_IPT];
    uint32_t keepbits = 0, cnt = 0;
    uint64_t last_a = 0, last_b = 0; // lane 31's element of the previous q
#pragma unroll
    for (int q = 0; q < UNIQ_IPT; ++q) {
        const uint32_t li = wseg + q * 32 + lane;
        const bool valid = li < tile_n;
        const uint64_t g = t0 + li;
        xa[q] = valid ? (WIDE ? a.in_a[g] : a.in_packed[g]) : 0;
        xb[q] = (valid && WIDE) ? a.in_b[g] : 0;
        uint64_t pa = __shfl_up_sync(BP_FULL_MASK, xa[q], 1);
        uint64_t pb = __shfl_up_sync(BP_FULL_MASK, xb[q], 1);
        if (lane == 0) {
            if (q == 0) {
                if (valid && g > 0) {
                    pa = WIDE ? a.in_a[g - 1] : a.in_packed[g - 1];
                    pb = WIDE ? a.in_b[g - 1] : 0;
                }
            } else {
                pa = last_a;
                pb = last_b;
            }
        }
        const bool keep = valid && (g == 0 || xa[q] != pa || (WIDE && xb[q] != pb));
        const unsigned m = __ballot_sync(BP_FULL_MASK, keep);
        pos[q] = cnt + __popc(m & lt);
        cnt += __popc(m);
        if (keep) keepbits |= 1u << q;
        last_a = __shfl_sync(BP_FULL_MASK, xa[q], 31);
        last_b = __shfl_sync(BP_FULL_MASK, xb[q], 31);
    }
    if (lane == 0) swtot[warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        const uint32_t t = lane < WARPS ? swtot[lane] : 0;
        const uint32_t ti = warp_inclusive_sum(t);
        if (lane < WARPS) swtot[lane] = ti - t;
        const uint32_t tile_keep = __shfl_sync(BP_FULL_MASK, ti, 31);
        if (lane == 0) swtot[WARPS] = tile_keep;
        const uint64_t e = lookback_exclusive(a.status, tile, (uint64_t)tile_keep, a.err);
        if (lane == 0) sbase = e;
    }
    __syncthreads();
    const uint64_t base = sbase + swtot[warp];
#pragma unroll
    for (int q = 0; q < UNIQ_IPT; ++q) {
        if (keepbits & (1u << q)) {
            const uint64_t g = base + pos[q];
            if (WIDE) {
                ulonglong2 p;
                p.x = xa[q]; // later
                p.y = xb[q]; // earlier
                ((ulonglong2 *)a.out)[g] = p;
            } else {
                uint2 p;
                p.x = (uint32_t)(xa[q] >> 32); // later
                p.y = (uint32_t)xa[q];         // earlier
                ((uint2 *)a.out)[g] = p;
            }
        }
    }
    if (t0 + tile_n == n && tid == 0) a.totals->n_pairs = sbase + swtot[WARPS];
}

// ---------------------------------------------------------------------------------------------
// pair_finish_kernel -- the second half of `collisions.sort_unstable(); collisions.dedup()`
// (src/layer.rs:473-474, 516-517) when the radix passes only ordered the pairs by their LATER ID.
//
// A later ID has a handful of partners (config 2: 3 raw pairs per object on average), so after a
// stable radix sort on the later ID alone every group of equal later IDs is tiny.  Ordering the
// earlier IDs inside such a group and dropping duplicates is quadratic work on 3-element groups --
// far cheaper than the three more full radix passes over the earlier ID's bits it replaces.
//   keep(e)  <=>  no element before e in its group has the same earlier ID
//   rank(e)   =   number of kept elements of the group with a smaller earlier ID
// A tile owns the groups that START in it; its window extends FIN_HALO elements past the tile so those
// groups are complete.  A group that does not fit sets `overflow` and the host falls back to the
// full-width sort + pair_unique_kernel (no result is lost: the input is only read).
// ---------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;
constexpr int FIN_HALO = 512;
template <class IdT> struct FinCfg { // u64 IDs stage two arrays twice: a smaller tile keeps the kernel under 48 KB of shared memory
    static constexpr int TILE = sizeof(IdT) == 8 ? 512 : 2048;
    static constexpr int WIN = TILE + FIN_HALO;
    static constexpr int IPT = WIN / FIN_THREADS;
};

template <class IdT> struct FinishArgs {
    const uint64_t *in_packed; // u32 IDs: packed pairs sorted by the later ID (the high half)
    const uint64_t *in_a;      // u64 IDs: later IDs (sorted) ...
    const uint64_t *in_b;      //          ... and their partners
    uint32_t n;
    uint32_t gshift;        // u32 IDs: pairs with equal (packed >> gshift) form a group; 32 = one group per later ID, more when
                            // the radix passes left the lowest bits of the later ID to this kernel (a group then holds a few
                            // later IDs and is ordered by the whole packed pair)
    IdT *out;               // [2 * n] (later, earlier) interleaved
    uint64_t *status;       // look-back, one per tile, zeroed
    uint32_t *tile_counter; // zeroed
    ScanTotals *totals;     // n_pairs; pad = 1 on overflow
    int *err;
};

// Per element: one walk over its group gives its stable position inside the group ordered by the
// earlier ID (#smaller + #equal-before); the group is rewritten in that order in a second shared buffer,
// where duplicates are adjacent, so dedup + ordered compaction + coalesced stores are the usual
// adjacent-difference / scan / look-back.
// LOWBITS (u32 IDs): a group is the pairs with equal (packed >> gshift), gshift > 32 -- a few later IDs -- ordered by the
// whole packed pair; otherwise a group is one later ID (the high word), ordered by the earlier ID (the low word).
template <class IdT, bool LOWBITS = false>
__global__ void __launch_bounds__(FIN_THREADS) pair_finish_kernel(const FinishArgs<IdT> a) {
    constexpr bool WIDE = sizeof(IdT) == 8;
    static_assert(!(WIDE && LOWBITS), "u64 IDs are grouped by the later ID");
    constexpr int FIN_TILE = FinCfg<IdT>::TILE, FIN_WIN = FinCfg<IdT>::WIN, FIN_IPT = FinCfg<IdT>::IPT;
    constexpr uint64_t HOLE = ~0ull; // marks window slots this tile does not own
    __shared__ uint64_t sa[FIN_WIN + 1];            // packed pair, or later ID (+1: the element after the window, for the overflow test)
    __shared__ uint64_t sb[WIDE ? FIN_WIN + 1 : 1]; // earlier ID (u64 IDs only)
    __shared__ uint64_t ta[FIN_WIN];                // the owned groups, each ordered by the earlier ID
    __shared__ uint64_t tb[WIDE ? FIN_WIN : 1];
    __shared__ uint32_t sscratch[FIN_THREADS / 32 + 2];
    __shared__ uint64_t sbase;
    __shared__ uint32_t stile;
    constexpr int FIN_ROWS = FIN_WIN / 32;
    __shared__ uint32_t heads[FIN_ROWS];
    static_assert(FIN_WIN % FIN_THREADS == 0, "whole rows of the head bitmap");

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) stile = atomicAdd(a.tile_counter, 1u);
    __syncthreads();
    const uint32_t tile = stile;
    const uint64_t t0 = (uint64_t)tile * FIN_TILE;
    if (t0 >= a.n) return;
    const uint32_t tile_n = (uint32_t)min((uint64_t)FIN_TILE, (uint64_t)a.n - t0);
    const uint32_t win_n = (uint32_t)min((uint64_t)FIN_WIN, (uint64_t)a.n - t0);
    const bool more = t0 + win_n < a.n; // an element exists right after the window

    for (uint32_t i = tid; i < win_n + (more ? 1u : 0u); i += FIN_THREADS) {
        sa[i] = WIDE ? a.in_a[t0 + i] : a.in_packed[t0 + i];
        if (WIDE) sb[i] = a.in_b[t0 + i];
    }
    for (uint32_t i = tid; i < (uint32_t)FIN_WIN; i += FIN_THREADS) {
        ta[i] = HOLE;
        if (WIDE) tb[i] = HOLE; // (max, max) is not a pair, so both halves equal to HOLE can only be a hole
    }
    // does the first group of the window continue a group of the previous tile?
    const uint32_t gs = LOWBITS ? a.gshift : 32u;
    const uint64_t prev = t0 > 0 ? (WIDE ? a.in_a[t0 - 1] : (a.in_packed[t0 - 1] >> gs)) : 0;
    __syncthreads();
    const bool cont0 = t0 > 0 && prev == (WIDE ? sa[0] : (sa[0] >> gs));

    // ---- group heads of the window as a bitmap (one ballot per 32 elements) -----------------------------------
    auto grp = [&](uint64_t y) -> uint64_t { return WIDE ? y : (y >> gs); };
#pragma unroll
    for (int k = 0; k < FIN_IPT; ++k) {
        const uint32_t i = k * FIN_THREADS + tid;
        bool head = false;
        if (i < win_n) head = i == 0 ? !cont0 : grp(sa[i]) != grp(sa[i - 1]);
        const uint32_t hb = __ballot_sync(BP_FULL_MASK, head);
        if (lane == 0) heads[k * (FIN_THREADS / 32) + warp] = hb;
    }
    __syncthreads();

    // ---- per element: the ends of its group from the bitmap, then its stable position inside the group: the number of
    // members before it that are not larger + the number after it that are smaller.  The members are visited by
    // DISTANCE, d = 1 .. (largest extent of a group around any of the warp's 32 elements): one trip count for all lanes (a
    // REDUX), lane l looking at elements i - d and i + d -- consecutive lanes, consecutive words.
#pragma unroll 1
    for (int k = 0; k < FIN_IPT; ++k) {
        const uint32_t i = k * FIN_THREADS + tid;
        const bool live = i < win_n;
        const uint32_t r = i >> 5, b = i & 31u;
        uint32_t m = heads[r] & (0xffffffffu >> (31u - b));
        int rr = (int)r;
        while (m == 0 && rr > 0) m = heads[--rr];
        const uint32_t hs = (uint32_t)rr * 32u + 31u - (uint32_t)__clz((int)m);
        uint32_t m2 = heads[r] & (0xfffffffeu << b);
        int re = (int)r;
        while (m2 == 0 && re + 1 < FIN_ROWS) m2 = heads[++re];
        const uint32_t he = m2 ? (uint32_t)re * 32u + (uint32_t)__ffs((int)m2) - 1u : win_n; // no head after it: up to the end of the window
        // owned: the group starts in this tile (m == 0: it continues a group of the previous tile)
        const bool owned = live && m != 0 && hs < tile_n;
        const uint64_t x = sa[live ? i : 0u];
        // an owned group that runs past the window cannot be finished here
        if (owned && he == win_n && more && grp(sa[win_n]) == grp(x)) a.totals->pad = 1u;
        const uint32_t nb = owned ? i - hs : 0u, na = owned ? he - 1u - i : 0u;
        const uint32_t dmax = __reduce_max_sync(BP_FULL_MASK, max(nb, na));
        // what orders the group: the earlier ID, or the whole packed pair
        const uint64_t eb = WIDE ? sb[live ? i : 0u] : (LOWBITS ? x : (x & 0xffffffffull));
        uint32_t pos = 0;
        for (uint32_t d = 1; d <= dmax; ++d) {
            const uint32_t jb = i - min(d, i), ja = min(i + d, (uint32_t)FIN_WIN);
            const uint64_t ob = WIDE ? sb[jb] : (LOWBITS ? sa[jb] : (sa[jb] & 0xffffffffull));
            const uint64_t oa = WIDE ? sb[ja] : (LOWBITS ? sa[ja] : (sa[ja] & 0xffffffffull));
            pos += (d <= nb && ob <= eb) ? 1u : 0u; // before it: smaller or equal ones come first (stable)
            pos += (d <= na && oa < eb) ? 1u : 0u;  // after it: only strictly smaller ones
        }
        if (owned) {
            ta[hs + pos] = x;
            if (WIDE) tb[hs + pos] = sb[i];
        }
    }
    __syncthreads();

    // ---- duplicates are adjacent now: keep the first of each run, compact in order ---------------------
    uint32_t keepbits = 0, mine = 0;
#pragma unroll
    for (int q = 0; q < FIN_IPT; ++q) { // blocked: thread t owns window slots [t*IPT, t*IPT + IPT)
        const uint32_t i = tid * FIN_IPT + q;
        bool keep = false;
        if (i < win_n && (ta[i] != HOLE || (WIDE && tb[i] != HOLE))) {
            keep = i == 0 || ta[i - 1] != ta[i] || (WIDE && tb[i - 1] != tb[i]);
        }
        if (keep) {
            keepbits |= 1u << q;
            ++mine;
        }
    }
    uint32_t tile_keep;
    uint32_t ex = block_exclusive_sum<FIN_THREADS, uint32_t>(mine, sscratch, &tile_keep);
    if (warp == 0) {
        const uint64_t e = lookback_exclusive(a.status, tile, (uint64_t)tile_keep, a.err);
        if (lane == 0) sbase = e;
    }
    // kept elements move to the front of the (now dead) input buffers, in order
    uint64_t ka[FIN_IPT], kb[WIDE ? FIN_IPT : 1];
#pragma unroll
    for (int q = 0; q < FIN_IPT; ++q) {
        const uint32_t i = tid * FIN_IPT + q;
        if (keepbits & (1u << q)) {
            ka[q] = ta[i];
            if (WIDE) kb[q] = tb[i];
        }
    }
#pragma unroll
    for (int q = 0; q < FIN_IPT; ++q) {
        if (keepbits & (1u << q)) {
            sa[ex] = ka[q];
            if (WIDE) sb[ex] = kb[q];
            ++ex;
        }
    }
    __syncthreads();
    const uint64_t base = sbase;
    for (uint32_t i = tid; i < tile_keep; i += FIN_THREADS) {
        if (WIDE) {
            ulonglong2 p;
            p.x = sa[i]; // later
            p.y = sb[i]; // earlier
            ((ulonglong2 *)a.out)[base + i] = p;
        } else {
            uint2 p;
            p.x = (uint32_t)(sa[i] >> 32); // later
            p.y = (uint32_t)sa[i];         // earlier
            ((uint2 *)a.out)[base + i] = p;
        }
    }
    if (t0 + tile_n == a.n && tid == 0) a.totals->n_pairs = base + tile_keep;
}

} // namespace bp
