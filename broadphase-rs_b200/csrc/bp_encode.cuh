// bp_encode.cuh -- K1: fused contains + quantise + depth choice + cell enumeration + Morton encode.
//
// Replaces the sequential loop of Layer::extend (src/layer.rs:94-121) and everything it calls:
// Bounds::contains (src/geom.rs:121-128), SystemBounds::to_local (src/geom.rs:148-163),
// IndexGenerator::indices / indices_at_depth (src/geom.rs:189-304), set_depth / set_origin /
// encode_axis (src/index.rs:106-112, 155-172, 193-207, 230-250).
//
// One pass over the objects: a CTA stages a tile of AABBs in shared memory with 16-byte loads,
// every thread quantises its objects with explicitly rounded IEEE ops (sub, div, mul, add -- never
// fused, never a reciprocal), counts its cells, the tile's record range comes from a block scan + a
// decoupled look-back over the preceding tiles, and the records are staged in shared memory and
// written in the reference's order (object order; z outermost, then y, x innermost) with coalesced
// stores.  The kernel also folds in the reductions the sort planner needs (OR / AND of all keys
// and IDs, "IDs ascending" check), so no extra pass over the records is required.
#pragma once

#include "bp_common.cuh"

namespace bp {

struct ExtendResult {
    unsigned long long total_records; // records this call produced (all of them, even if capacity ran out)
    unsigned long long n_invalid;     // objects rejected by contains()
    unsigned long long key_or, key_and, id_or, id_and;
    unsigned int nonmono;  // 1 if an ID smaller than its predecessor was seen
    unsigned int too_many; // 1 if one object wanted more than ENCODE_MAX_CELLS cells
    unsigned long long id_first, id_last; // IDs of the first / last object of the call (valid or not): with !nonmono, the ID range
};

constexpr uint32_t ENCODE_MAX_CELLS = 1u << 20;
constexpr int ENCODE_THREADS = 256;
constexpr int ENCODE_OPT = 4; // objects per thread
constexpr int ENCODE_TILE = ENCODE_THREADS * ENCODE_OPT;
constexpr int ENCODE_ROUND = 8192; // records generated per round (= ENCODE_TILE * 2^3: one round at natural depth)
constexpr int ENCODE_LUT_BITS = 10;

// Optional by-product for the multi-GPU frame: while the splitters of the previous frame are still good, the records are
// counted per destination shard (and per halo copy) as they are generated -- the separate counting pass over the keys
// (partition_hist_kernel) and its launch disappear.  Same definition: home = #splitters <= key, halo copies for the
// shards home+1 .. #splitters <= run_upper_key(key).
constexpr int ENCODE_MAX_SPLITTERS = 15;
struct EncodeCount {
    uint64_t spl[ENCODE_MAX_SPLITTERS];
    uint32_t n_spl;
    uint32_t *cnt; // [0, 16): records per home shard, [16, 32): halo copies per shard; zeroed by the host
};

template <class T, class IdT> struct EncodeArgs {
    const float *bounds; // n x 2*DIM
    const IdT *ids;      // n
    uint32_t n;
    float sys_min[3], sys_max[3], sys_size[3];
    uint32_t min_depth;
    typename T::key_t *keys_out; // tree arrays
    IdT *ids_out;
    uint8_t *cell_flags_out; // per record: bit a set = the object also covers the previous cell along axis a (or null)
    uint64_t out_base;     // records already in the tree
    uint64_t capacity;     // records the tree arrays can hold
    uint64_t *status;      // look-back status, one per tile, zeroed
    uint32_t *tile_counter; // zeroed
    ExtendResult *result;  // initialised by the host (sums 0, and-masks ~0)
    const IdT *prev_last_id; // last ID of the previous extend since the tail began (or null)
    IdT *next_last_id;
    int *err;
    const uint32_t *lut;   // [2^ENCODE_LUT_BITS] Morton spread table of the layer's dimension (built once by the host)
    EncodeCount count; // used by encode_kernel<.., COUNT = true> only
};

// Shared-memory layout.  The staged AABBs are dead once every thread has quantised its objects, so the
// per-round owner table reuses their space.
template <class T, class IdT> struct EncodeSmem {
    static constexpr size_t BOUNDS_BYTES = (size_t)ENCODE_TILE * 2 * T::DIM * sizeof(float);
    static constexpr size_t OWNER_BYTES = (size_t)ENCODE_ROUND * sizeof(uint16_t);
    static constexpr size_t UNION_BYTES = BOUNDS_BYTES > OWNER_BYTES ? BOUNDS_BYTES : OWNER_BYTES;
    static constexpr size_t DESC_OFF = UNION_BYTES;                                         // uint4 per object
    static constexpr size_t ID_OFF = DESC_OFF + (size_t)ENCODE_TILE * sizeof(uint4);        // IdT per object
    static constexpr size_t CNT_OFF = ID_OFF + (size_t)ENCODE_TILE * sizeof(IdT);           // u32 per object (+1)
    static constexpr size_t LUT_OFF = CNT_OFF + (size_t)(ENCODE_TILE + 4) * sizeof(uint32_t); // u32 x 2^LUT_BITS
    static constexpr size_t RED_OFF = LUT_OFF + ((size_t)4 << ENCODE_LUT_BITS);
    static constexpr size_t BYTES = RED_OFF + 64 * sizeof(uint64_t);
};

// SystemBounds::to_local for one scalar (src/geom.rs:148-156): ((g - min) / size * RANGE + 0) as u32
// with one IEEE rounding per operation; the cast truncates, saturates and maps NaN to 0 exactly like
// Rust's `as u32` (cvt.rzi.u32.f32).
__device__ __forceinline__ uint32_t quantise(float g, float mn, float size) {
    const float t = __fsub_rn(g, mn);
    const float q = __fdiv_rn(t, size);
    const float m = __fmul_rn(q, 4294967040.0f);
    const float r = __fadd_rn(m, 0.0f);
    return __float2uint_rz(r);
}

// Morton spread of a cell coordinate through a 2^10-entry shared table (bit i -> bit DIM*i): two
// (3D, <= 19 bits) or three (2D, <= 29 bits) look-ups instead of a 5-step 64-bit mask cascade.
template <int DIM> __device__ __forceinline__ uint64_t spread_lut(const uint32_t *lut, uint32_t c) {
    constexpr uint32_t M = (1u << ENCODE_LUT_BITS) - 1u;
    uint64_t r = (uint64_t)lut[c & M] | ((uint64_t)lut[(c >> ENCODE_LUT_BITS) & M] << (DIM * ENCODE_LUT_BITS));
    if (DIM == 2) r |= (uint64_t)lut[(c >> (2 * ENCODE_LUT_BITS)) & M] << (2 * DIM * ENCODE_LUT_BITS);
    return r;
}

template <class T, class IdT, bool COUNT = false>
__global__ void __launch_bounds__(ENCODE_THREADS) encode_kernel(const EncodeArgs<T, IdT> a) {
    typedef typename T::key_t K;
    __shared__ uint32_t scount[COUNT ? 32 : 1];
    __shared__ uint64_t sspl[COUNT ? ENCODE_MAX_SPLITTERS + 1 : 1];
    if (COUNT && threadIdx.x < 32) scount[threadIdx.x] = 0; // (published by the barriers below, long before its first use)
    if (COUNT && threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < ENCODE_MAX_SPLITTERS; ++i) sspl[i] = a.count.spl[i]; // static indices: from the constant bank
    }
    // COUNT: records per home shard, one 8-bit field per shard in two packed words (a thread generates at most
    // 8 cells x 4 objects = 32 records of a tile), reduced over the warp once per tile
    unsigned long long pc0 = 0, pc1 = 0;
    constexpr int DIM = T::DIM;
    constexpr int FPO = 2 * DIM; // floats per object
    typedef EncodeSmem<T, IdT> S;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sbounds = (float *)smem_raw;
    uint16_t *sowner = (uint16_t *)smem_raw;               // [ENCODE_ROUND] object of every record of the round
    uint4 *sdesc = (uint4 *)(smem_raw + S::DESC_OFF);      // {cx0, cy0, cz0, depth | nx << 8 | ny << 16 (or general marker)}
    IdT *sid = (IdT *)(smem_raw + S::ID_OFF);
    uint32_t *scnt = (uint32_t *)(smem_raw + S::CNT_OFF);  // [ENCODE_TILE + 1] exclusive record offsets
    uint32_t *slut = (uint32_t *)(smem_raw + S::LUT_OFF);
    uint64_t *sred = (uint64_t *)(smem_raw + S::RED_OFF);  // scratch

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) *(uint32_t *)sred = atomicAdd(a.tile_counter, 1u);
    // the spread table (entry v = bits of v moved to positions DIM * i) is copied, not computed: 2^10 five-step mask
    // cascades per 1024-object tile were 13 % of the kernel's instructions (profiles/r2_cfg5_frame_kernels.txt)
    for (uint32_t v = tid; v < (1u << ENCODE_LUT_BITS); v += ENCODE_THREADS) slut[v] = a.lut[v];
    __syncthreads();
    const uint32_t tile = *(uint32_t *)sred;
    __syncthreads();
    const uint32_t obj0 = tile * ENCODE_TILE;
    if (obj0 >= a.n) return;
    const uint32_t tile_objs = min((uint32_t)ENCODE_TILE, a.n - obj0);

    // ---- stage the tile's AABBs with 16-byte loads ------------------------------------------------
    {
        const float *src = a.bounds + (size_t)obj0 * FPO;
        const uint32_t nfl = tile_objs * FPO;
        if ((((uintptr_t)src) & 15u) == 0) {
            const uint32_t nv = nfl >> 2;
            const float4 *src4 = (const float4 *)src;
            float4 *dst4 = (float4 *)sbounds;
            for (uint32_t i = tid; i < nv; i += ENCODE_THREADS) dst4[i] = __ldcs(src4 + i);
            for (uint32_t i = (nv << 2) + tid; i < nfl; i += ENCODE_THREADS) sbounds[i] = src[i];
        } else {
            for (uint32_t i = tid; i < nfl; i += ENCODE_THREADS) sbounds[i] = src[i];
        }
    }
    __syncthreads();

    // ---- per object: contains, quantise, depth, cell grid ----------------------------------------
    uint32_t count[ENCODE_OPT];
    uint32_t n_invalid = 0, nonmono = 0, too_many = 0;
    unsigned long long id_or = 0, id_and = ~0ull;
#pragma unroll
    for (int k = 0; k < ENCODE_OPT; ++k) {
        const uint32_t o = k * ENCODE_THREADS + tid; // striped: conflict-light shared reads, coalesced ID loads
        count[k] = 0;
        if (o >= tile_objs) continue;
        const uint32_t g = obj0 + o;
        const IdT id = a.ids[g];
        sid[o] = id;
        // IDs ascending in object order => records enter the sort in ascending ID order, so a stable
        // sort on the key alone yields the (Index, ID) order of src/layer.rs:146-165.
        if (g > 0) {
            if (a.ids[g - 1] > id) nonmono = 1;
        } else if (a.prev_last_id) {
            if (*a.prev_last_id > id) nonmono = 1;
        }
        const float *b = sbounds + o * FPO;
        bool valid = true;
        uint32_t lmin[3] = {0, 0, 0}, lmax[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const float bmin = b[i], bmax = b[DIM + i];
            // Bounds::contains -- src/geom.rs:121-128 (NaN passes both tests)
            if (a.sys_min[i] > bmin || a.sys_max[i] < bmax) valid = false;
            lmin[i] = quantise(bmin, a.sys_min[i], a.sys_size[i]);
            lmax[i] = quantise(bmax, a.sys_min[i], a.sys_size[i]);
        }
        if (!valid) {
            ++n_invalid;
            continue;
        }
        id_or |= (unsigned long long)id;
        id_and &= (unsigned long long)id;
        // indices(): depth = leading_zeros(max_axis(sizei) - 1), raised to min_depth, clamped
        uint32_t max_axis = 0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) max_axis = max(max_axis, lmax[i] - lmin[i] + 1u);
        uint32_t d = (uint32_t)__clz((int)(max_axis - 1u));
        d = max(d, a.min_depth);
        d = min(d, (uint32_t)T::AXIS_BITS);
        uint32_t c0[3] = {0, 0, 0}, nc[3] = {1, 1, 1};
        uint64_t cells = 1;
        if (d != 0) { // indices_at_depth(): cell coordinates at depth d = the top d bits of the local coordinate
            const uint32_t sh = 32u - d;
#pragma unroll
            for (int i = 0; i < DIM; ++i) {
                const uint32_t lo = lmin[i] >> sh, hi = lmax[i] >> sh;
                c0[i] = lo;
                nc[i] = hi > lo ? hi - lo + 1u : 1u;
                cells *= nc[i];
            }
        }
        if (cells > ENCODE_MAX_CELLS) {
            too_many = 1;
            cells = 0;
        }
        count[k] = (uint32_t)cells;
        // nx, ny <= 255 fit the packed descriptor; larger grids (min_depth far below the natural depth) carry
        // nx in .w's upper bits and recompute ny from the count -- kept simple: nx in bits 8..19, ny in 20..31
        uint4 dsc;
        dsc.x = c0[0];
        dsc.y = c0[1];
        dsc.z = c0[2];
        dsc.w = d | (min(nc[0], 0xfffu) << 8) | (min(nc[1], 0xfffu) << 20);
        if (nc[0] > 0xfffu || nc[1] > 0xfffu) { // cannot be described: refuse like an oversized object
            too_many = 1;
            count[k] = 0;
        }
        sdesc[o] = dsc;
    }
    __syncthreads(); // sbounds is dead from here on (its space becomes the owner table)

    // ---- exclusive scan of the counts in object order ---------------------------------------------
#pragma unroll
    for (int k = 0; k < ENCODE_OPT; ++k) scnt[k * ENCODE_THREADS + tid] = count[k];
    __syncthreads();
    uint32_t c4[ENCODE_OPT], tsum = 0;
#pragma unroll
    for (int k = 0; k < ENCODE_OPT; ++k) {
        c4[k] = scnt[tid * ENCODE_OPT + k];
        tsum += c4[k];
    }
    uint32_t tile_total;
    uint32_t ex = block_exclusive_sum<ENCODE_THREADS, uint32_t>(tsum, (uint32_t *)sred, &tile_total);
#pragma unroll
    for (int k = 0; k < ENCODE_OPT; ++k) {
        scnt[tid * ENCODE_OPT + k] = ex;
        ex += c4[k];
    }
    if (tid == 0) scnt[ENCODE_TILE] = tile_total;
    __syncthreads();

    // ---- tile offset: decoupled look-back -----------------------------------------------------------
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive(a.status, tile, (uint64_t)tile_total, a.err);
        if (lane == 0) sred[32] = excl;
    }

    // ---- generate: one record per thread and step ---------------------------------------------------
    // Every object first stamps its index on the records it owns (a loop of its cell count, one 2-byte
    // store each); then the threads walk the records in order: record p looks up its owner, turns its
    // cell number into (ix, iy, iz), spreads the three cell coordinates through the table and stores the
    // key and the ID straight to their final place -- consecutive threads, consecutive records.
    unsigned long long key_or = 0, key_and = ~0ull;
    uint32_t off[ENCODE_OPT];
#pragma unroll
    for (int k = 0; k < ENCODE_OPT; ++k) off[k] = scnt[k * ENCODE_THREADS + tid];
    uint64_t gout = 0;
    for (uint32_t w0 = 0; w0 < tile_total; w0 += ENCODE_ROUND) {
        const uint32_t w1 = min(tile_total, w0 + (uint32_t)ENCODE_ROUND);
#pragma unroll
        for (int k = 0; k < ENCODE_OPT; ++k) {
            const uint32_t lo = max(off[k], w0), hi = min(off[k] + count[k], w1);
            for (uint32_t p = lo; p < hi; ++p) sowner[p - w0] = (uint16_t)(k * ENCODE_THREADS + tid);
        }
        __syncthreads(); // also publishes sred[32] in the first round
        if (w0 == 0) gout = a.out_base + sred[32];
        for (uint32_t p = w0 + tid; p < w1; p += ENCODE_THREADS) {
            const uint32_t o = sowner[p - w0];
            const uint4 dsc = sdesc[o];
            const uint32_t d = dsc.w & 0xffu, nx = (dsc.w >> 8) & 0xfffu, ny = dsc.w >> 20;
            uint32_t c = p - scnt[o];
            uint32_t ix, iy, iz;
            if (nx <= 2 && ny <= 2) { // the natural-depth case: at most two cells per axis
                ix = c & (nx - 1u);
                c >>= (nx - 1u);
                iy = c & (ny - 1u);
                iz = c >> (ny - 1u);
            } else {
                const uint32_t nxy = nx * ny;
                iz = c / nxy;
                c -= iz * nxy;
                iy = c / nx;
                ix = c - iy * nx;
            }
            uint64_t origin = spread_lut<DIM>(slut, dsc.x + ix) | (spread_lut<DIM>(slut, dsc.y + iy) << 1);
            if (DIM == 3) origin |= spread_lut<DIM>(slut, dsc.z + iz) << 2;
            // the spread of a coordinate shifted to the top of its AXIS_BITS field is the spread of the cell
            // coordinate shifted by DIM * (AXIS_BITS - d)
            const K key = d ? make_key<T>(d, origin << (DIM * (T::AXIS_BITS - (int)d))) : (K)0; // depth 0 -> Index::default()
            key_or |= (unsigned long long)key;
            key_and &= (unsigned long long)key;
            const uint64_t gi = gout + p;
            if (gi < a.capacity) {
                a.keys_out[gi] = key;
                a.ids_out[gi] = sid[o];
                if (a.cell_flags_out) a.cell_flags_out[gi] = (uint8_t)((ix ? 1u : 0u) | (iy ? 2u : 0u) | (iz ? 4u : 0u));
            }
            if constexpr (COUNT) {
                const uint32_t ns = a.count.n_spl;
                const uint32_t home = splitter_rank15(sspl, (uint64_t)key);
                if (home < 8)
                    pc0 += 1ull << (8 * home);
                else
                    pc1 += 1ull << (8 * (home - 8));
                if (home < ns) { // common case: the cell ends before the next splitter -- one comparison
                    const uint64_t hi = (uint64_t)run_upper_key<T>(key);
                    if (hi >= sspl[home]) {
                        const uint32_t last = splitter_rank15(sspl, hi);
                        for (uint32_t s2 = home + 1; s2 <= last; ++s2) atomicAdd(&scount[16 + s2], 1u); // rare
                    }
                }
            }
        }
        __syncthreads();
    }
    if (tile_total == 0) __syncthreads(); // no round ran: still wait for the look-back result
    const uint64_t tile_base = sred[32];

    // ---- block-level reductions for the sort planner ------------------------------------------------
    key_or = warp_or(key_or);
    key_and = warp_and(key_and);
    id_or = warp_or(id_or);
    id_and = warp_and(id_and);
    n_invalid = warp_sum(n_invalid);
    nonmono = warp_or(nonmono);
    too_many = warp_or(too_many);
    if (lane == 0) {
        if (key_or) atomicOr(&a.result->key_or, key_or);
        if (key_and != ~0ull) atomicAnd(&a.result->key_and, key_and);
        if (id_or) atomicOr(&a.result->id_or, id_or);
        if (id_and != ~0ull) atomicAnd(&a.result->id_and, id_and);
        if (n_invalid) atomicAdd(&a.result->n_invalid, (unsigned long long)n_invalid);
        if (nonmono) atomicOr(&a.result->nonmono, 1u);
        if (too_many) atomicOr(&a.result->too_many, 1u);
    }
    if constexpr (COUNT) {
        for (uint32_t b = 0; b <= a.count.n_spl; ++b) { // (warp-uniform trip count; every thread is here)
            const uint32_t c = warp_sum((uint32_t)((b < 8 ? pc0 >> (8 * b) : pc1 >> (8 * (b - 8))) & 0xffull));
            if (lane == 0 && c) atomicAdd(&scount[b], c);
        }
        __syncthreads();
        if (tid < 32 && scount[tid]) atomicAdd(&a.count.cnt[tid], scount[tid]);
    }
    if (obj0 + tile_objs == a.n && tid == 0) { // last tile: totals + the tail's last ID
        a.result->total_records = tile_base + tile_total;
        a.result->id_last = (unsigned long long)a.ids[a.n - 1];
        if (a.next_last_id) *a.next_last_id = a.ids[a.n - 1];
    }
    if (obj0 == 0 && tid == 0) a.result->id_first = (unsigned long long)a.ids[0];
}

// The cell flags ride through the sort in the unused top 3 bits of the ID payload (the host checks that
// the IDs leave them free) and are removed again before the IDs are shown to anybody.
template <class IdT> __global__ void __launch_bounds__(256) flags_merge_kernel(IdT *__restrict__ ids, const uint8_t *__restrict__ flags, uint32_t n) {
    constexpr int SH = 8 * sizeof(IdT) - 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        ids[i] |= (IdT)flags[i] << SH;
}
template <class IdT> __global__ void __launch_bounds__(256) flags_strip_kernel(IdT *__restrict__ ids, uint32_t n) {
    constexpr IdT MASK = (IdT)(((IdT)1 << (8 * sizeof(IdT) - 3)) - 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) ids[i] &= MASK;
}

// OR / AND of keys and IDs + "IDs ascending" for a tree that did not come from encode_kernel
// (bp_layer_set_records).  result must be pre-initialised like for encode_kernel.
template <class K, class IdT>
__global__ void __launch_bounds__(256) record_masks_kernel(const K *__restrict__ keys, const IdT *__restrict__ ids,
                                                           uint32_t n, IdT id_mask /* removes cell flags */, ExtendResult *result) {
    unsigned long long key_or = 0, key_and = ~0ull, id_or = 0, id_and = ~0ull;
    unsigned nonmono = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = (unsigned long long)keys[i], v = (unsigned long long)(ids[i] & id_mask);
        key_or |= k;
        key_and &= k;
        id_or |= v;
        id_and &= v;
        if (i > 0 && (ids[i - 1] & id_mask) > (ids[i] & id_mask)) nonmono = 1;
    }
    key_or = warp_or(key_or);
    key_and = warp_and(key_and);
    id_or = warp_or(id_or);
    id_and = warp_and(id_and);
    nonmono = warp_or(nonmono);
    if ((threadIdx.x & 31) == 0) {
        if (key_or) atomicOr(&result->key_or, key_or);
        if (key_and != ~0ull) atomicAnd(&result->key_and, key_and);
        if (id_or) atomicOr(&result->id_or, id_or);
        if (id_and != ~0ull) atomicAnd(&result->id_and, id_and);
        if (nonmono) atomicOr(&result->nonmono, 1u);
    }
}

} // namespace bp
