// bp_exchange.cuh -- the fused partition + all-to-all pass of the multi-GPU path.
//
// No reference counterpart (the reference is one process): this is how the records -- and later the raw pairs -- reach
// the GPU that owns their Morton range.  One read of the (key, payload) arrays; every tile is stably partitioned into
// <= 16 buckets (destination shards, found by comparing the key with the g - 1 splitters) and each (tile, bucket) run is
// written straight into the destination GPU's receive buffer through its NVLink peer mapping, at
// [start of this rank's chunk in that buffer] + [keys of the bucket in earlier tiles] (decoupled look-back over one
// 16-word status row per tile).  The partition is stable, so a receive buffer holds every source's records in their
// original order -- which is what lets the receiver sort on the key alone when the IDs ascend (dist.sort_plan).
//
// Round 1 used the generic onesweep pass with a splitter digit for this; its per-item destination look-up indexed two
// 16-entry kernel-parameter arrays dynamically (a 560-byte stack frame, ~1.3 KB of spill traffic per thread) and its time
// was local ranking PLUS NVLink drain.  Here the runs are long (a tile of 4608 records over <= 16 buckets), so the drain is
// bucket-major: the destination of a run is a warp-uniform scalar, and -- template parameter TMA -- the 16-byte aligned
// body of every run leaves the SM as ONE cp.async.bulk shared -> global copy (keys) plus one for the payload: the staging
// area places each bucket's run at the 16-byte phase of its destination, the at most 1 key / 3 IDs before and after the
// aligned body go by ordinary stores.  (tools/scatter_tma_probe, profiles/r2_scatter_tma_probe.log: per-thread stores and
// bulk copies drain into a peer at the same ~710 GB/s and both overlap with the ranking of the other resident CTAs; the
// bulk path frees the issue slots of 4608 x 2 store instructions per tile.)
#pragma once

#include "bp_radix.cuh"

namespace bp {

constexpr int XCH_BUCKETS = 16; // MAX_SPLITTERS + 1

template <class K, class V> struct ExchangeArgs {
    const K *kin;
    const V *vin;          // unused for V = NoVal
    const uint8_t *vflags; // optional: 3 cell-flag bits per input record, OR-ed into the top bits of the payload
    uint32_t n;
    uint32_t n_spl; // buckets in use = n_spl + 1
    uint32_t shift; // bucket = number of splitters <= (key >> shift)
    uint64_t spl[MAX_SPLITTERS];
    uint64_t kdst[XCH_BUCKETS]; // device address at which this rank's chunk starts in every bucket's key destination
    uint64_t vdst[XCH_BUCKETS]; // ... and payload destination
    uint32_t *status;           // [tiles][XCH_BUCKETS], zeroed; bits 31..30 flag, 29..0 count
    uint32_t *tile_counter;     // zeroed
    int *err;
};

template <class K, class V, int THREADS, int ITEMS> struct ExchangeCfg {
    static constexpr bool HAS_V = !std::is_same<V, NoVal>::value;
    static constexpr int TILE = THREADS * ITEMS, WARPS = THREADS / 32;
    static constexpr int KPER = 16 / sizeof(K);                                  // keys per 16 bytes
    static constexpr int VSIZE = HAS_V ? (int)sizeof(V) : 4;
    static constexpr int VPER = 16 / VSIZE;                                      // payloads per 16 bytes
    static constexpr size_t KEY_BYTES = ((size_t)(TILE + XCH_BUCKETS * KPER) * sizeof(K) + 15) / 16 * 16;
    static constexpr size_t VAL_BYTES = HAS_V ? ((size_t)(TILE + XCH_BUCKETS * VPER) * VSIZE + 15) / 16 * 16 : 0;
    static constexpr size_t SMEM_BYTES = KEY_BYTES + VAL_BYTES + (size_t)(WARPS * XCH_BUCKETS + 10 * XCH_BUCKETS + 16) * sizeof(uint32_t);
    static_assert(TILE < 65536, "ranks are packed in 16 bits");
};

__device__ __forceinline__ void bulk_store_s2g(uint64_t gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

// Lanes of the warp in the same bucket (<= 16 buckets: four ballots).
__device__ __forceinline__ unsigned match_bucket(uint32_t d) { return match_digit<4>(d); }

template <class K, class V, int THREADS, int ITEMS, bool TMA, bool FULL>
__device__ __forceinline__ void exchange_tile(const ExchangeArgs<K, V> &a, unsigned char *smem_raw, const uint32_t tile, const uint32_t tile_n) {
    typedef ExchangeCfg<K, V, THREADS, ITEMS> Cfg;
    constexpr int TILE = Cfg::TILE, WARPS = Cfg::WARPS, NB = XCH_BUCKETS;
    constexpr bool HAS_V = Cfg::HAS_V;
    typedef typename std::conditional<HAS_V, V, uint32_t>::type VT;

    K *skeys = (K *)smem_raw;
    VT *svals = (VT *)(smem_raw + Cfg::KEY_BYTES);
    uint32_t *whist = (uint32_t *)(smem_raw + Cfg::KEY_BYTES + Cfg::VAL_BYTES); // [WARPS][NB]
    uint32_t *cnt = whist + WARPS * NB;                                        // [NB] keys of the tile per bucket (without pads)
    uint32_t *gofs = cnt + NB;                                                 // [NB] keys of the bucket in earlier tiles
    uint32_t *kstart = gofs + NB;                                              // [NB] where the bucket's run starts in skeys
    uint32_t *vstart = kstart + NB;                                            // [NB] ... in svals
    uint64_t *sdst = (uint64_t *)(vstart + NB);                                // [2 * NB] kdst | vdst (8-byte aligned: 4 * NB words precede)
    uint64_t *sspl = sdst + 2 * NB;                                            // [NB] splitters (unused ones ~0)
    uint32_t *misc = (uint32_t *)(sspl + NB);                                  // [0] tile, [1] poison

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint64_t tile_begin = (uint64_t)tile * TILE;
    const uint32_t nb = a.n_spl + 1;

    // ---- load keys, warp-striped: item k of lane l of warp w is tile element w*32*ITEMS + k*32 + l ----
    const K *kin = a.kin + tile_begin;
    const uint32_t base_i = warp * (32 * ITEMS) + lane;
    K key[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t i = base_i + k * 32;
        if (FULL)
            key[k] = ld_stream(kin + i);
        else
            key[k] = (i < tile_n) ? ld_stream(kin + i) : (K) ~(K)0; // pads fall into the last bucket, behind its real keys
    }
    // payload loads in flight across the ranking
    VT val[HAS_V ? ITEMS : 1];
    if constexpr (HAS_V) {
        const V *vin = a.vin + tile_begin;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = base_i + k * 32;
            if (FULL || i < tile_n) val[k] = ld_stream(vin + i);
        }
        if (a.vflags) { // dedup at the source across the exchange: the cell flags leave in the IDs' spare top bits
            const uint8_t *fl = a.vflags + tile_begin;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t i = base_i + k * 32;
                if (FULL || i < tile_n) val[k] |= (VT)fl[i] << (8 * sizeof(VT) - 3);
            }
        }
    }

    // ---- bucket + stable rank inside (warp, bucket) ----
    const unsigned lt = lanemask_lt();
    uint32_t *wrow = whist + warp * NB;
    uint32_t rd[ITEMS]; // low 16 bits: rank inside the (warp, bucket) group, high 16 bits: bucket
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = min(splitter_rank15(sspl, (uint64_t)key[k] >> a.shift), a.n_spl); // (an all-ones pad passes the ~0 fillers too)
        const unsigned m = match_bucket(d);
        uint32_t old = 0;
        if ((m & lt) == 0) old = atomicAdd(&wrow[d], (uint32_t)__popc(m)); // one shared atomic per group, in item order: stable
        old = __shfl_sync(BP_FULL_MASK, old, __ffs(m) - 1);
        rd[k] = (old + __popc(m & lt)) | (d << 16);
    }
    __syncthreads();

    // ---- warp 0: per-bucket totals, publish, look back, lay the staging area out ----
    if (warp == 0) {
        uint32_t count = 0, excl = 0;
        if (lane < NB) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t c = whist[w * NB + lane];
                whist[w * NB + lane] = run;
                run += c;
            }
            count = run;
            if (!FULL && lane == nb - 1) count -= (uint32_t)(TILE - tile_n);
            uint32_t *st = a.status + (size_t)tile * NB + lane;
            st_volatile_u32(st, (tile == 0 ? RS_FLAG_INC : RS_FLAG_AGG) | count);
            if (tile != 0) {
                constexpr int B = 8; // predecessor tiles per round trip
                int64_t t = (int64_t)tile - 1;
                bool done = false;
                while (!done && t >= 0) {
                    uint32_t sv[B];
#pragma unroll
                    for (int j = 0; j < B; ++j) sv[j] = (t - j >= 0) ? ld_volatile_u32(a.status + (size_t)(t - j) * NB + lane) : RS_FLAG_INC;
#pragma unroll
                    for (int j = 0; j < B; ++j) {
                        if (done) break;
                        uint32_t sb = sv[j];
                        if ((sb >> 30) == 0) {
                            const uint32_t *ps = a.status + (size_t)(t - j) * NB + lane;
                            uint32_t spins = 0;
                            do {
                                if (++spins > BP_SPIN_LIMIT) { // a predecessor never published: do not write anywhere
                                    *a.err = 1;
                                    misc[1] = 1;
                                    sb = RS_FLAG_INC;
                                    break;
                                }
                                __nanosleep(20);
                                sb = ld_volatile_u32(ps);
                            } while ((sb >> 30) == 0);
                        }
                        excl += sb & RS_VALUE_MASK;
                        if ((sb >> 30) == 2) done = true;
                    }
                    t -= B;
                }
                st_volatile_u32(st, RS_FLAG_INC | ((excl + count) & RS_VALUE_MASK));
            }
            cnt[lane] = count;
            gofs[lane] = excl;
        }
        // Staging layout: bucket b's run starts at the 16-byte phase of its destination, so that the aligned body of the
        // run is aligned in shared memory too (cp.async.bulk needs both).  Sequential over <= 16 buckets, by shuffles.
        const uint32_t kph = lane < NB ? (uint32_t)((sdst[lane] / sizeof(K) + excl) & (Cfg::KPER - 1)) : 0;
        const uint32_t vph = (HAS_V && lane < NB) ? (uint32_t)((sdst[NB + lane] / sizeof(VT) + excl) & (Cfg::VPER - 1)) : 0;
        uint32_t kpos = 0, vpos = 0, my_k = 0, my_v = 0;
        for (uint32_t b = 0; b < nb; ++b) {
            const uint32_t cb = __shfl_sync(BP_FULL_MASK, count, b);
            const uint32_t kp = __shfl_sync(BP_FULL_MASK, kph, b), vp = __shfl_sync(BP_FULL_MASK, vph, b);
            kpos += (kp - kpos) & (Cfg::KPER - 1);
            vpos += (vp - vpos) & (Cfg::VPER - 1);
            if (lane == b) {
                my_k = kpos;
                my_v = vpos;
            }
            const uint32_t span = (!FULL && b == nb - 1) ? cb + (uint32_t)(TILE - tile_n) : cb; // pads are staged too (never drained)
            kpos += span;
            vpos += span;
        }
        if (lane < NB) {
            kstart[lane] = my_k;
            vstart[lane] = my_v;
        }
    }
    __syncthreads();

    // ---- stage keys and payloads bucket by bucket ----
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = rd[k] >> 16;
        const uint32_t r = (rd[k] & 0xffffu) + wrow[d];
        skeys[kstart[d] + r] = key[k];
        if constexpr (HAS_V) svals[vstart[d] + r] = val[k];
    }
    if constexpr (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the TMA engine
    __syncthreads();
    if (misc[1]) return; // look-back timed out (error already raised): write nothing rather than to a wrong place

    // ---- drain ----
    if constexpr (TMA) {
        if (tid < nb) {
            const uint32_t b = tid, c = cnt[b], e = gofs[b];
            if (c) {
                {
                    constexpr uint32_t PER = Cfg::KPER;
                    const uint64_t base = sdst[b];
                    const K *src = skeys + kstart[b];
                    const uint32_t head = min((PER - (uint32_t)((base / sizeof(K) + e) & (PER - 1))) & (PER - 1), c);
                    const uint32_t mid = (c - head) / PER * PER;
                    for (uint32_t j = 0; j < head; ++j) ((K *)base)[e + j] = src[j];
                    if (mid) bulk_store_s2g(base + (uint64_t)(e + head) * sizeof(K), src + head, mid * (uint32_t)sizeof(K));
                    for (uint32_t j = head + mid; j < c; ++j) ((K *)base)[e + j] = src[j];
                }
                if constexpr (HAS_V) {
                    constexpr uint32_t PER = Cfg::VPER;
                    const uint64_t base = sdst[NB + b];
                    const VT *src = svals + vstart[b];
                    const uint32_t head = min((PER - (uint32_t)((base / sizeof(VT) + e) & (PER - 1))) & (PER - 1), c);
                    const uint32_t mid = (c - head) / PER * PER;
                    for (uint32_t j = 0; j < head; ++j) ((VT *)base)[e + j] = src[j];
                    if (mid) bulk_store_s2g(base + (uint64_t)(e + head) * sizeof(VT), src + head, mid * (uint32_t)sizeof(VT));
                    for (uint32_t j = head + mid; j < c; ++j) ((VT *)base)[e + j] = src[j];
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the CTA's shared memory may go once it has been read
        }
    } else {
        for (uint32_t b = 0; b < nb; ++b) { // warp-uniform: the destination of a run is a scalar
            const uint32_t c = cnt[b], e = gofs[b];
            K *kd = (K *)sdst[b] + e;
            const K *ks = skeys + kstart[b];
            for (uint32_t i = tid; i < c; i += THREADS) kd[i] = ks[i];
            if constexpr (HAS_V) {
                VT *vd = (VT *)sdst[NB + b] + e;
                const VT *vs = svals + vstart[b];
                for (uint32_t i = tid; i < c; i += THREADS) vd[i] = vs[i];
            }
        }
    }
}

template <class K, class V, int THREADS, int ITEMS, int MINB, bool TMA>
__global__ void __launch_bounds__(THREADS, MINB) exchange_pass_kernel(const ExchangeArgs<K, V> a) {
    typedef ExchangeCfg<K, V, THREADS, ITEMS> Cfg;
    constexpr int TILE = Cfg::TILE, WARPS = Cfg::WARPS, NB = XCH_BUCKETS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *whist = (uint32_t *)(smem_raw + Cfg::KEY_BYTES + Cfg::VAL_BYTES);
    uint64_t *sdst = (uint64_t *)(whist + WARPS * NB + 4 * NB);
    uint64_t *sspl = sdst + 2 * NB;
    uint32_t *misc = (uint32_t *)(sspl + NB);

    const unsigned tid = threadIdx.x;
    if (tid == 0) {
        misc[0] = atomicAdd(a.tile_counter, 1u);
        misc[1] = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) { // static indices: straight from the constant bank
            sdst[b] = a.kdst[b];
            sdst[NB + b] = a.vdst[b];
            sspl[b] = b < MAX_SPLITTERS ? a.spl[b < MAX_SPLITTERS ? b : 0] : ~0ull; // (the host pads unused splitters with ~0)
        }
    }
    for (int i = tid; i < WARPS * NB; i += THREADS) whist[i] = 0;
    __syncthreads();
    const uint32_t tile = misc[0];
    const uint64_t tile_begin = (uint64_t)tile * TILE;
    if (tile_begin >= a.n) return;
    const uint32_t tile_n = (uint32_t)min((uint64_t)TILE, (uint64_t)a.n - tile_begin);
    if (tile_n == (uint32_t)TILE)
        exchange_tile<K, V, THREADS, ITEMS, TMA, true>(a, smem_raw, tile, tile_n);
    else
        exchange_tile<K, V, THREADS, ITEMS, TMA, false>(a, smem_raw, tile, tile_n);
}

} // namespace bp
