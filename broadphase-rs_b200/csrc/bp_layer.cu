// bp_layer.cu -- host runtime + C ABI (include/bp.h) of libbroadphase_b200.so.
//
// Mirrors the reference's Layer<Index, ID> (src/layer.rs:42-68): a device-resident tree of
// (Index, ID) records in SoA form (keys[], ids[]) with the reference's "sorted" flag, plus the
// scratch the kernels need.  All work is enqueued on one CUDA stream per layer; the host only
// synchronises where a size has to come back (record count after extend, work-item / pair counts
// in scan).  There is no CPU fallback: every entry point fails if CUDA does.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "bp_common.cuh"
#include "bp_encode.cuh"
#include "bp_merge.cuh"
#include "bp_query.cuh"
#include "bp_radix.cuh"
#include "bp_exchange.cuh"
#include "bp_scan.cuh"

using namespace bp;

namespace {

// ---- tunables -------------------------------------------------------------------------------------
template <class K, class V> struct PassTune { // threads, items per thread, min CTAs/SM of radix_pass_kernel
    static constexpr int THREADS = 384, ITEMS = 12, MINB = 2;
};
// tuned with tools/sort_bench on B200 (profiles/r1_sort_tuning.md)
template <> struct PassTune<uint64_t, uint32_t> { static constexpr int THREADS = 384, ITEMS = 12, MINB = 3; };
template <> struct PassTune<uint32_t, uint32_t> { static constexpr int THREADS = 512, ITEMS = 16, MINB = 2; };
template <> struct PassTune<uint64_t, NoVal> { static constexpr int THREADS = 384, ITEMS = 12, MINB = 3; };
template <> struct PassTune<uint64_t, uint64_t> { static constexpr int THREADS = 256, ITEMS = 12, MINB = 3; };
// Shapes per digit width.  A 9-bit pass has half as many keys per (tile, digit) run as an 8-bit pass, and the write-out is
// what it pays for (profiles/r2_sortbench_*.log: 384 x 12 x 3: 1.34 ms against 1.07 ms per pass on 146 M records); the larger
// tile of 384 x 16 (2 CTAs / SM, 80 registers: fewer spills) brings it to 1.24 ms, so a 9-bit plan wins whenever it saves
// a pass out of at most seven.  10-bit passes (1.6-1.8 ms) never pay.
template <class K, class V, int RB> struct PassTuneRB : PassTune<K, V> {};
template <> struct PassTuneRB<uint64_t, uint32_t, 9> { static constexpr int THREADS = 384, ITEMS = 16, MINB = 2; };
template <> struct PassTuneRB<uint64_t, NoVal, 9> { static constexpr int THREADS = 384, ITEMS = 16, MINB = 2; };

constexpr uint64_t MAX_RECORDS = (1ull << 30) - 1;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ProfEvent {
    cudaEvent_t start, stop;
    int cls;
};

// How a count gets back to the host.  A cudaMemcpyAsync + stream/event synchronise costs ~22 us of idle GPU per
// round trip at config 2 (two copy operations, the wake-up of the waiting thread, then the launch latency of what
// follows: three such gaps were 14 % of the frame, profiles/r1_timeline_cfg2.txt).  Instead a one-thread kernel
// stores the few words into pinned, device-mapped host memory, fences, and stores a sequence number; the host
// spins on that number (checking the stream now and then, so a failed kernel cannot hang it).
struct Mailbox {
    unsigned long long words[12];
    unsigned int err;
    unsigned int seq;
};

__global__ void post_kernel(const unsigned long long *__restrict__ src, int nwords, const int *__restrict__ err, Mailbox *mb,
                            unsigned int seq) { // one warp: the loads of the words overlap
    const int lane = (int)threadIdx.x;
    if (lane < nwords) mb->words[lane] = src[lane];
    if (lane == 31) mb->err = (unsigned int)*err;
    __threadfence_system();
    __syncwarp();
    if (lane == 0) *(volatile unsigned int *)&mb->seq = seq;
}

} // namespace

// The count row of a counting extend (Impl::extend_count_rows): per-shard counts from encode_kernel<.., COUNT>, then the tag
// words broadphase-rs_b200/dist.py expects (N_TAGS = 7), built from the extend's result block without leaving the device:
// id_or | (bit 63: the IDs leave their top 3 bits free and the caller allows folding), key_or, key_and, id_and, first ID,
// last ID, IDs ascending -- an empty tree reports first = 2^64 - 1, last = 0, ascending = 1 like bp_layer_id_order.
__global__ void count_row_result_kernel(const uint32_t *__restrict__ cnt, uint32_t nb, const bp::ExtendResult *__restrict__ res,
                                        int allow_fold, int id_bits, bp::RowDst dst) {
    const uint32_t i = threadIdx.x;
    const bool empty = res->total_records == 0;
    uint64_t tag = 0;
    switch (i) {
    case 0: tag = res->id_or | ((allow_fold && (res->id_or >> (id_bits - 3)) == 0) ? (1ull << 63) : 0ull); break;
    case 1: tag = res->key_or; break;
    case 2: tag = res->key_and; break;
    case 3: tag = res->id_and; break;
    case 4: tag = empty ? ~0ull : res->id_first; break;
    case 5: tag = empty ? 0ull : res->id_last; break;
    case 6: tag = (empty || !res->nonmono) ? 1ull : 0ull; break;
    default: break;
    }
    const uint64_t c = i < nb ? cnt[i] : 0, h = i < nb ? cnt[16 + i] : 0;
    for (uint32_t r = 0; r < dst.n; ++r) {
        uint64_t *row = dst.p[r];
        if (i < nb) {
            row[i] = c;
            row[nb + i] = h;
        }
        if (i < 7) row[2 * nb + i] = tag;
    }
}

struct bp_layer {
    bp_layer_config cfg;
    int kind, id_bytes, key_bytes, dim, device;
    uint32_t min_depth;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    // tree: ping-pong SoA buffers
    DevBuf keys[2], ids[2];
    int cur = 0;
    size_t cap_records = 0;
    uint64_t n_records = 0;
    bool dirty = false;       // !sorted flag of the reference
    uint64_t prefix = 0;      // records [0, prefix) are sorted (meaningful while dirty)
    bool tail_sorted = false; // records [prefix, n) form one sorted run
    bool tail_nonmono = false; // tail IDs not known to ascend in record order
    bool tail_has_last = false;
    int last_slot = 0;
    uint64_t key_or = 0, key_and = ~0ull, id_or = 0, id_and = ~0ull;
    uint64_t id_first = 0, id_last = 0; // IDs of the first / last object extended since the last clear (bp_layer_id_order)
    uint64_t n_invalid = 0;
    uint64_t n_halo = 0; // records [0, n_halo) only act as ancestors in scan (multi-GPU halos)
    uint64_t last_raw_pairs = 0; // raw pairs of the last scan that took the one-kernel path (sizes the next scan's output)
    // Layer::merge of a sorted layer into a sorted layer is deferred: the other layer's records are not copied behind ours,
    // the next sort merges the two runs straight out of the two layers' buffers (one read of the other tree instead of
    // copy + read).  Every other access to this layer's records, and every call that may change the other layer,
    // materialises the copy first (resolve_pending), so nothing observable differs from the eager append.
    bp_layer *lazy_src = nullptr;           // the layer whose records logically sit at [n_records - lazy_n, n_records)
    uint64_t lazy_n = 0;
    std::vector<bp_layer *> lazy_readers;   // layers holding a deferred merge from this one
    int radix_bits_cap = 0; // BP_RADIX_BITS (tuning aid): widest radix digit the sorts may use; 0 = no cap
    uint64_t sort_finish_min = 1ull << 18; // record sorts of at least this many records may take the "top bits + finish" plan
                                           // (it ends with a host round trip); BP_SORT_FINISH_MIN, 0 = never
    uint32_t sort_finish_cooldown = 0;     // sorts to go before the plan is tried again after a group overflowed its window
    bool scan_dedup = true; // bp_layer_set_scan_dedup: the scan may emit every ID pair from its canonical shared cell only
    // dedup at the source: encode writes 3 cell flags per record (cell_flags); a full sort of a tree that
    // only holds encoded records moves them into the top 3 bits of the IDs (ids_flagged) so that they
    // travel with the records; any other mutation strips them again
    DevBuf cell_flags;
    DevBuf spread_lut; // 2^10-entry Morton spread table of the layer's dimension (encode_kernel)
    bool flags_valid = true; // every record of the tree has its flags in cell_flags
    bool ids_flagged = false;

    // pending extend result
    bool pending = false;
    uint64_t pending_base = 0;
    cudaEvent_t ev_sync = nullptr;
    cudaStream_t copy_stream = nullptr;     // bp_layer_extend_host: chunked H2D copies run here, the encodes trail behind them
    cudaEvent_t ev_chunk[16] = {};
    ExtendResult *h_res = nullptr; // pinned
    ExtendResult *d_res = nullptr;
    uint32_t *d_cnt = nullptr;     // 32 words: per-shard record / halo counts taken by encode_kernel<.., COUNT> (multi-GPU)
    void *d_last = nullptr;        // 2 x u64 slots
    ScanTotals *h_tot = nullptr;   // pinned
    ScanTotals *d_tot = nullptr;
    int *d_err = nullptr;
    int *h_err = nullptr; // pinned
    Mailbox *h_mail[2] = {nullptr, nullptr}; // pinned + mapped: [0] extend results, [1] scan / query totals
    Mailbox *d_mail[2] = {nullptr, nullptr}; // the same memory as the device sees it
    unsigned int mail_seq = 0, pending_seq = 0;

    DevBuf scratch;                // look-back status words, tile counters, histograms
    DevBuf stage_bounds, stage_ids; // extend_host staging
    DevBuf src_idx, src_off, chunk_src, inactive;
    DevBuf praw[2], praw_b[2];     // raw pairs, ping-pong (u32 IDs: packed; u64 IDs: later / earlier)
    DevBuf pout;                   // final pairs
    DevBuf pick_shapes, pick_out;  // pick_ray: staged shape table (host callers), one bp_pick_result per ray
    void *h_pick = nullptr;        // pinned mirror of pick_out
    size_t h_pick_cap = 0;
    DevBuf query_params, query_counts, query_offsets; // batched queries: geometry parameters, per-query counts, CSR offsets
    void *h_offsets = nullptr;     // pinned mirror of query_offsets
    size_t h_offsets_cap = 0;
    void *pairs_src = nullptr;     // finish_pairs: the raw pairs live here instead of praw[0] (bp_layer_unique_pairs_inplace_device)
    bool pairs_grouped = false;    // finish_pairs: the raw pairs are already grouped by their first ID, in order
    uint64_t pair_later_fixed = 0; // bp_layer_set_pair_later_fixed: bits of the later ID that agree in every pair of the next call
    DevBuf pair_cnt;               // pairs per later ID (counting sort of the pairs; dense 32-bit IDs only)
    bool want_pair_counts = false; // the caller of scan_raw will finish the pairs itself (scan), not hand them out raw
    uint64_t pair_cnt_n = 0;       // > 0: the last emission counted its pairs per later ID into pair_cnt[0, pair_cnt_n)
    DevBuf filter_table;
    void *h_pairs = nullptr;
    size_t h_pairs_cap = 0;
    void *h_keys = nullptr, *h_ids = nullptr;
    size_t h_keys_cap = 0, h_ids_cap = 0;
    uint64_t n_pairs = 0;

    bool profiling = false;
    std::vector<ProfEvent> prof_events;
    std::vector<ProfEvent> prof_pool;
    bp_stats stats;
    std::string last_error;
};

namespace {

int fail(bp_layer *L, int status, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (L) L->last_error = buf;
    return status;
}

#define CU(L, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            cudaGetLastError();                                                                              \
            return fail((L), e_ == cudaErrorMemoryAllocation ? BP_ERR_OOM : BP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                         \
        }                                                                                                    \
    } while (0)

#define TRY(expr)                   \
    do {                            \
        int s_ = (expr);            \
        if (s_ != BP_OK) return s_; \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ensure(bp_layer *L, DevBuf &b, size_t bytes, bool preserve = false, size_t preserve_bytes = 0) {
    if (bytes <= b.cap) return BP_OK;
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    want = (want + 255) & ~(size_t)255;
    void *np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = (bytes + 255) & ~(size_t)255;
        e = cudaMalloc(&np, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(L, BP_ERR_OOM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
    }
    if (preserve && b.p && preserve_bytes) {
        CU(L, cudaMemcpyAsync(np, b.p, preserve_bytes, cudaMemcpyDeviceToDevice, L->stream));
    }
    if (b.p) {
        CU(L, cudaStreamSynchronize(L->stream));
        CU(L, cudaFree(b.p));
    }
    b.p = np;
    b.cap = want;
    return BP_OK;
}

void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ---- launch bookkeeping -------------------------------------------------------------------------------
struct LaunchScope {
    bp_layer *L;
    int cls;
    ProfEvent pe;
    bool timed;
    LaunchScope(bp_layer *L_, int cls_, double bytes) : L(L_), cls(cls_), timed(false) {
        L->stats.launches[cls] += 1;
        L->stats.launches_total += 1;
        L->stats.algo_bytes[cls] += bytes;
        if (L->profiling) {
            if (!L->prof_pool.empty()) {
                pe = L->prof_pool.back();
                L->prof_pool.pop_back();
            } else {
                cudaEventCreate(&pe.start);
                cudaEventCreate(&pe.stop);
            }
            pe.cls = cls;
            cudaEventRecord(pe.start, L->stream);
            timed = true;
        }
    }
    ~LaunchScope() {
        if (timed) {
            cudaEventRecord(pe.stop, L->stream);
            L->prof_events.push_back(pe);
        }
    }
};

void detach_lazy(bp_layer *L);
int materialize_lazy(bp_layer *L);
int resolve_pending(bp_layer *L, bool keep_lazy = false);

int check_launch(bp_layer *L, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(L, BP_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return BP_OK;
}

// Enqueues the post of `nwords` 64-bit words at d_src (+ the error flag) to mailbox `box`; returns the sequence number.
int post_mail(bp_layer *L, int box, const void *d_src, int nwords, unsigned int *out_seq) {
    const unsigned int seq = ++L->mail_seq;
    {
        LaunchScope ls(L, BP_K_MISC, 0);
        post_kernel<<<1, 32, 0, L->stream>>>((const unsigned long long *)d_src, nwords, L->d_err, L->d_mail[box], seq);
    }
    *out_seq = seq;
    return check_launch(L, "post_kernel");
}

// Waits until mailbox `box` carries `seq`, then copies `bytes` of it to `dst`.
int wait_mail(bp_layer *L, int box, unsigned int seq, void *dst, size_t bytes) {
    volatile Mailbox *mb = L->h_mail[box];
    for (unsigned int spins = 1; mb->seq != seq; ++spins) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause(); // a polite spin: the sibling hyper-thread and the memory system are not hammered
#endif
        if ((spins & 0xffffu) == 0) { // the stream must still be busy (or have just finished); anything else is an error
            const cudaError_t e = cudaStreamQuery(L->stream);
            if (e == cudaSuccess) {
                if (mb->seq == seq) break;
                return fail(L, BP_ERR_INTERNAL, "the stream drained without posting its result");
            }
            if (e != cudaErrorNotReady) {
                cudaGetLastError();
                return fail(L, BP_ERR_CUDA, "waiting for a device result: %s", cudaGetErrorString(e));
            }
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    memcpy(dst, (const void *)L->h_mail[box]->words, bytes);
    if (L->h_mail[box]->err) {
        cudaMemsetAsync(L->d_err, 0, sizeof(int), L->stream);
        return fail(L, BP_ERR_INTERNAL, "a kernel reported a look-back time-out");
    }
    return BP_OK;
}

void collect_profile(bp_layer *L) {
    // BP_TIMELINE=1: print where every launch of the collected interval started and ended (ms after the first one),
    // i.e. the gaps the stream spent on memsets, small copies and host round trips (a tuning aid)
    static const bool timeline = getenv("BP_TIMELINE") != nullptr;
    if (timeline && !L->prof_events.empty()) {
        static const char *names[BP_K_COUNT] = {"encode", "sort_hist", "sort_pass", "merge", "scan_runs", "scan_emit",
                                                "pair_hist", "pair_pass", "pair_unique", "misc", "query", "partition", "sort_finish"};
        float prev_end = 0.f;
        for (ProfEvent &pe : L->prof_events) {
            float t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&t0, L->prof_events[0].start, pe.start);
            cudaEventElapsedTime(&t1, L->prof_events[0].start, pe.stop);
            fprintf(stderr, "[bp timeline] %-11s start %8.4f  dur %7.4f  gap before %7.4f\n", names[pe.cls], t0, t1 - t0, t0 - prev_end);
            prev_end = t1;
        }
    }
    for (ProfEvent &pe : L->prof_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pe.start, pe.stop) == cudaSuccess) L->stats.kernel_ms[pe.cls] += ms;
        L->prof_pool.push_back(pe);
    }
    L->prof_events.clear();
}

template <class T> inline T *ptr(DevBuf &b, size_t off = 0) { return (T *)b.p + off; }

// ---- radix planning --------------------------------------------------------------------------------------
// Digits for an LSD sort over the bits set in `mask` (bits that differ between keys).  Each pass
// takes whichever covers more of the remaining bits: one W-bit window starting at the lowest
// remaining bit, or two runs of consecutive set bits (the lowest run, then the next one after a
// gap) of W bits in total.  W = 8 by default; the record sort takes 9- or 10-bit digits when that saves a
// whole pass (27 varying bits: 3 passes instead of 4; 43: 5 instead of 6).
int plan_passes(uint64_t mask, RadixPlan &plan, int first = 0, int W = 8) {
    int np = first;
    const uint64_t wmask = (1ull << W) - 1ull;
    auto run_len = [](uint64_t m, int from, int cap) {
        int n = 0;
        while (n < cap && from + n < 64 && ((m >> (from + n)) & 1ull)) ++n;
        return n;
    };
    while (mask && np < RADIX_MAX_PASSES) {
        const int s0 = __builtin_ctzll(mask);
        // option A: one window
        const uint64_t window = (mask >> s0) & wmask;
        const int bitsA = 64 - __builtin_clzll(window);
        const int coverA = __builtin_popcountll(window);
        // option B: two runs
        const int len0 = run_len(mask, s0, W);
        int s1 = 0, len1 = 0;
        if (len0 < W) {
            const uint64_t rest = (s0 + len0 >= 64) ? 0 : (mask >> (s0 + len0)) << (s0 + len0);
            if (rest) {
                s1 = __builtin_ctzll(rest);
                len1 = run_len(mask, s1, W - len0);
            }
        }
        if (len0 + len1 > coverA) {
            plan.shift[np] = (unsigned char)s0;
            plan.bits[np] = (unsigned char)len0;
            plan.shift2[np] = (unsigned char)s1;
            plan.bits2[np] = (unsigned char)len1;
            for (int i = 0; i < len0; ++i) mask &= ~(1ull << (s0 + i));
            for (int i = 0; i < len1; ++i) mask &= ~(1ull << (s1 + i));
        } else {
            plan.shift[np] = (unsigned char)s0;
            plan.bits[np] = (unsigned char)bitsA;
            plan.shift2[np] = 0;
            plan.bits2[np] = 0;
            mask &= ~(wmask << s0);
        }
        ++np;
    }
    plan.npasses = np;
    return np;
}

// "Top bits + finish" plan of a record sort (record_finish_kernel, bp_radix.cuh): radix passes over the highest varying key
// bits only -- whole 8-bit passes, enough of them that a group of records agreeing on those bits is expected to hold at most
// ~8 records (2^bits >= cnt / 8) -- and one finish pass that orders every group by the rest.  Taken when it replaces at
// least two radix passes (the finish pass costs about one).  *top = the bits the radix passes sort on; records with equal
// (key >> *gshift) form a group.
bool plan_top_bits(uint64_t kmask, uint64_t cnt, uint64_t *top, uint32_t *gshift) {
    const int nv = __builtin_popcountll(kmask);
    RadixPlan full;
    memset(&full, 0, sizeof full);
    const int np_full = plan_passes(kmask, full, 0, 8);
    if (np_full < 3 || cnt < 2) return false;
    const int lg = 64 - __builtin_clzll(cnt - 1); // ceil(log2(cnt))
    const int want = std::max(lg - 3, 8);
    const int nbits = std::min(8 * ((want + 7) / 8), nv);
    if (nbits >= nv) return false;
    uint64_t m = kmask;
    int low = 0;
    for (int i = 0; i < nbits; ++i) {
        low = 63 - __builtin_clzll(m);
        m &= ~(1ull << low);
    }
    const uint64_t t = kmask & ~((1ull << low) - 1ull);
    RadixPlan tp;
    memset(&tp, 0, sizeof tp);
    if (plan_passes(t, tp, 0, 8) + 2 > np_full) return false;
    *top = t;
    *gshift = (uint32_t)low;
    return true;
}

// scratch layout for one radix sort: [hist 16*256 u32][counters 16 u32][status passes*tiles*256 u32]
struct RadixScratch {
    uint32_t *hist, *counters, *status;
    size_t bytes;
};

// Which (key, payload) sorts may use digits wider than 8 bits (each width is another set of kernel instantiations).
template <class K, class V> struct WideDigits { static constexpr int MAX_BITS = 8; };
template <> struct WideDigits<uint64_t, uint32_t> { static constexpr int MAX_BITS = 9; }; // records
template <> struct WideDigits<uint64_t, NoVal> { static constexpr int MAX_BITS = 9; };    // packed pairs
// Measured on B200 (profiles/r2_sortbench_*.log, r2_bench_digit_width.txt): a 9-bit pass costs 1.24-1.32 ms where the 8-bit pass
// costs 1.02 ms (146 M records), so 3 x 9 bits beat 4 x 8 bits by only 3 % on the 2^25-object frame while the pass falls from
// 52 % to 41 % of the HBM roofline; the 8-bit plan stays the default, BP_RADIX_BITS=9 switches the wide digits on.
constexpr int DEFAULT_RADIX_BITS = 8;

template <class K, class V, int RB>
int radix_sort_rb(bp_layer *L, const RadixPlan &plan, K *k0, V *v0, K *k1, V *v1, uint32_t n, const uint32_t *n_dev, int cls_hist,
                  int cls_pass, int *out_passes, bool *out_in_alt, size_t elem_bytes, const uint8_t *first_pass_vflags, const K *k_src,
                  const V *v_src);

template <class K, class V>
int radix_sort(bp_layer *L, K *k0, V *v0, K *k1, V *v1, uint32_t n, const uint32_t *n_dev, uint64_t mask, int cls_hist,
               int cls_pass, int *out_passes, bool *out_in_alt, size_t elem_bytes, const uint8_t *first_pass_vflags = nullptr,
               const K *k_src = nullptr, const V *v_src = nullptr) {
    // k_src / v_src: the first pass (and the histograms) read these arrays instead of k0 / v0 and write k1 / v1 as usual --
    // sorting straight out of a buffer the layer does not own (a multi-GPU receive buffer) without a staging copy
    *out_in_alt = false;
    *out_passes = 0;
    if (n < 2) return BP_OK;
    // the narrowest digit that reaches the smallest number of passes
    RadixPlan plan;
    memset(&plan, 0, sizeof plan);
    int rb = 8, np = plan_passes(mask, plan, 0, 8);
    if (np == 0) return BP_OK;
    const int max_bits = std::min(L->radix_bits_cap ? L->radix_bits_cap : DEFAULT_RADIX_BITS, WideDigits<K, V>::MAX_BITS);
    for (int w = 9; w <= max_bits; ++w) {
        RadixPlan pw;
        memset(&pw, 0, sizeof pw);
        const int npw = plan_passes(mask, pw, 0, w);
        if (npw < np && npw * 7 <= np * 6) { // a wider pass costs about 1.17 x an 8-bit pass
            np = npw;
            rb = w;
            plan = pw;
        }
    }
    if constexpr (WideDigits<K, V>::MAX_BITS >= 9) {
        if (rb == 9)
            return radix_sort_rb<K, V, 9>(L, plan, k0, v0, k1, v1, n, n_dev, cls_hist, cls_pass, out_passes, out_in_alt, elem_bytes,
                                          first_pass_vflags, k_src, v_src);
        // Narrower digits when they need no extra pass (27 varying bits: 4 x 7 instead of 8 + 8 + 8 + 3): a 7-bit pass votes
        // once less per key and writes digit runs twice as long -- 0.92 against 1.00 ms per pass on 146 M records, 0.58 of
        // the HBM roofline (profiles/r2_sortbench_narrow_digits.log); a 6-bit pass 0.86 ms.
        if (rb == 8 && !(L->radix_bits_cap == 8)) {
            for (int w = 6; w <= 7; ++w) {
                RadixPlan pn;
                memset(&pn, 0, sizeof pn);
                if (plan_passes(mask, pn, 0, w) != np) continue;
                if (w == 6)
                    return radix_sort_rb<K, V, 6>(L, pn, k0, v0, k1, v1, n, n_dev, cls_hist, cls_pass, out_passes, out_in_alt, elem_bytes,
                                                  first_pass_vflags, k_src, v_src);
                return radix_sort_rb<K, V, 7>(L, pn, k0, v0, k1, v1, n, n_dev, cls_hist, cls_pass, out_passes, out_in_alt, elem_bytes,
                                              first_pass_vflags, k_src, v_src);
            }
        }
    }
    return radix_sort_rb<K, V, 8>(L, plan, k0, v0, k1, v1, n, n_dev, cls_hist, cls_pass, out_passes, out_in_alt, elem_bytes,
                                  first_pass_vflags, k_src, v_src);
}

template <class K, class V, int RB>
int radix_sort_rb(bp_layer *L, const RadixPlan &plan, K *k0, V *v0, K *k1, V *v1, uint32_t n, const uint32_t *n_dev, int cls_hist,
                  int cls_pass, int *out_passes, bool *out_in_alt, size_t elem_bytes, const uint8_t *first_pass_vflags, const K *k_src,
                  const V *v_src) {
    typedef PassTuneRB<K, V, RB> Tune;
    typedef RadixPassCfg<K, V, Tune::THREADS, Tune::ITEMS, RB> Cfg;
    constexpr int NB = Cfg::NB;
    const int np = plan.npasses;
    const uint32_t tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
    const size_t hist_bytes = (size_t)RADIX_MAX_PASSES * NB * sizeof(uint32_t);
    const size_t ctr_bytes = 64 * sizeof(uint32_t);
    const size_t status_bytes = (size_t)np * tiles * NB * sizeof(uint32_t);
    const size_t total = hist_bytes + ctr_bytes + status_bytes;
    TRY(ensure(L, L->scratch, total));
    uint32_t *hist = (uint32_t *)L->scratch.p;
    uint32_t *counters = hist + RADIX_MAX_PASSES * NB;
    uint32_t *status = counters + 64;
    CU(L, cudaMemsetAsync(L->scratch.p, 0, total, L->stream));
    {
        LaunchScope ls(L, cls_hist, (double)n * sizeof(K));
        const int blocks = (int)std::min<size_t>((n + 512 * 8 - 1) / (512 * 8), 148 * 8);
        radix_hist_kernel<K><<<std::max(blocks, 1), 512, (size_t)np * NB * sizeof(uint32_t), L->stream>>>(k_src ? k_src : k0, n, n_dev, plan,
                                                                                                         hist, NB);
    }
    TRY(check_launch(L, "radix_hist_kernel"));
    {
        LaunchScope ls(L, BP_K_MISC, 0);
        radix_scan_hist_kernel<NB><<<np, NB, 0, L->stream>>>(hist);
    }
    TRY(check_launch(L, "radix_scan_hist_kernel"));
    auto kern1 = radix_pass_kernel<K, V, Tune::THREADS, Tune::ITEMS, Tune::MINB, OneFieldDigit<K>, RB>;
    auto kern2 = radix_pass_kernel<K, V, Tune::THREADS, Tune::ITEMS, Tune::MINB, ShiftMaskDigit<K>, RB>;
    CU(L, cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    CU(L, cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    const K *kin = k_src ? k_src : k0;
    const V *vin = k_src ? v_src : v0;
    K *kout = k1;
    V *vout = v1;
    for (int p = 0; p < np; ++p) {
        LaunchScope ls(L, cls_pass, 2.0 * (double)n * (double)elem_bytes);
        if (plan.bits2[p] == 0) {
            RadixPassArgs<K, V, OneFieldDigit<K>> a;
            a.kin = kin;
            a.kout = kout;
            a.vin = vin;
            a.vout = vout;
            a.n_host = n;
            a.n_dev = n_dev;
            a.ghist_excl = hist + (size_t)p * NB;
            a.status = status + (size_t)p * tiles * NB;
            a.tile_counter = counters + p;
            a.op.shift = plan.shift[p];
            a.op.mask = (1u << plan.bits[p]) - 1u;
            a.vflags = p == 0 ? first_pass_vflags : nullptr;
            a.err = L->d_err;
            kern1<<<tiles, Tune::THREADS, Cfg::SMEM_BYTES, L->stream>>>(a);
        } else {
            RadixPassArgs<K, V, ShiftMaskDigit<K>> a;
            a.kin = kin;
            a.kout = kout;
            a.vin = vin;
            a.vout = vout;
            a.n_host = n;
            a.n_dev = n_dev;
            a.ghist_excl = hist + (size_t)p * NB;
            a.status = status + (size_t)p * tiles * NB;
            a.tile_counter = counters + p;
            a.op.shift = plan.shift[p];
            a.op.mask = (1u << plan.bits[p]) - 1u;
            a.op.shift2 = plan.shift2[p];
            a.op.mask2 = (1u << plan.bits2[p]) - 1u;
            a.op.bits = plan.bits[p];
            a.vflags = p == 0 ? first_pass_vflags : nullptr;
            a.err = L->d_err;
            kern2<<<tiles, Tune::THREADS, Cfg::SMEM_BYTES, L->stream>>>(a);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(L, BP_ERR_CUDA, "launch of radix_pass_kernel failed: %s", cudaGetErrorString(e));
        kin = kout;
        vin = vout;
        kout = (p & 1) ? k1 : k0;
        vout = (p & 1) ? v1 : v0;
    }
    *out_passes = np;
    *out_in_alt = (np & 1) != 0;
    return BP_OK;
}

// ---- per (index kind, id type) implementation ------------------------------------------------------------
template <int KIND, class IdT> struct Impl {
    typedef IndexTraits<KIND> T;
    typedef typename T::key_t K;

    static K *keys(bp_layer *L, int which) { return (K *)L->keys[which].p; }
    static IdT *ids(bp_layer *L, int which) { return (IdT *)L->ids[which].p; }

    static int ensure_tree(bp_layer *L, uint64_t records) {
        if (records <= L->cap_records) return BP_OK;
        if (records > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "tree would exceed 2^30 records");
        uint64_t want = std::max<uint64_t>(records, L->cap_records + L->cap_records / 2);
        want = std::min<uint64_t>(want, MAX_RECORDS);
        const size_t live = (size_t)L->n_records;
        TRY(ensure(L, L->keys[L->cur], want * sizeof(K), true, live * sizeof(K)));
        TRY(ensure(L, L->ids[L->cur], want * sizeof(IdT), true, live * sizeof(IdT)));
        TRY(ensure(L, L->keys[L->cur ^ 1], want * sizeof(K)));
        TRY(ensure(L, L->ids[L->cur ^ 1], want * sizeof(IdT)));
        TRY(ensure(L, L->cell_flags, want, true, live));
        L->cap_records = want;
        return BP_OK;
    }

    // Removes the cell flags from the IDs (before anything but the scan looks at them).
    static int strip_flags(bp_layer *L) {
        if (!L->ids_flagged) return BP_OK;
        if (L->n_records) {
            LaunchScope ls(L, BP_K_MISC, 2.0 * (double)L->n_records * sizeof(IdT));
            const int blocks = (int)std::min<uint64_t>((L->n_records + 1023) / 1024, 148 * 8);
            flags_strip_kernel<IdT><<<blocks, 256, 0, L->stream>>>(ids(L, L->cur), (uint32_t)L->n_records);
        }
        TRY(check_launch(L, "flags_strip_kernel"));
        L->ids_flagged = false;
        L->flags_valid = false;
        return BP_OK;
    }

    // ---- extend ------------------------------------------------------------------------------------------
    static int launch_encode(bp_layer *L, const float *sysb, const float *d_bounds, const IdT *d_ids, uint32_t n,
                             const EncodeCount *count = nullptr) {
        EncodeArgs<T, IdT> a;
        memset(&a.count, 0, sizeof a.count);
        if (count) a.count = *count;
        a.bounds = d_bounds;
        a.ids = d_ids;
        a.n = n;
        for (int i = 0; i < 3; ++i) a.sys_min[i] = a.sys_max[i] = a.sys_size[i] = 0.f;
        for (int i = 0; i < T::DIM; ++i) {
            a.sys_min[i] = sysb[i];
            a.sys_max[i] = sysb[T::DIM + i];
            volatile float s = sysb[T::DIM + i] - sysb[i]; // sizef -- src/geom.rs:97-102, one f32 rounding
            a.sys_size[i] = s;
        }
        a.min_depth = L->min_depth;
        a.keys_out = keys(L, L->cur);
        a.ids_out = ids(L, L->cur);
        a.cell_flags_out = L->flags_valid ? (uint8_t *)L->cell_flags.p : nullptr;
        a.out_base = L->n_records;
        a.capacity = L->cap_records;
        const uint32_t tiles = (n + ENCODE_TILE - 1) / ENCODE_TILE;
        const size_t sbytes = (size_t)tiles * sizeof(uint64_t) + 64;
        TRY(ensure(L, L->scratch, sbytes));
        CU(L, cudaMemsetAsync(L->scratch.p, 0, sbytes, L->stream));
        a.tile_counter = (uint32_t *)L->scratch.p;
        a.lut = (const uint32_t *)L->spread_lut.p;
        a.status = (uint64_t *)((char *)L->scratch.p + 64);
        // result accumulators: sums 0, and-masks all ones
        ExtendResult init;
        memset(&init, 0, sizeof init);
        init.key_and = ~0ull;
        init.id_and = ~0ull;
        *L->h_res = init;
        CU(L, cudaMemcpyAsync(L->d_res, L->h_res, sizeof init, cudaMemcpyHostToDevice, L->stream));
        a.result = L->d_res;
        IdT *slots = (IdT *)L->d_last; // two 8-byte slots
        a.prev_last_id = L->tail_has_last ? (const IdT *)((char *)slots + 8 * L->last_slot) : nullptr;
        a.next_last_id = (IdT *)((char *)slots + 8 * (L->last_slot ^ 1));
        a.err = L->d_err;
        typedef EncodeSmem<T, IdT> S;
        auto kern = count ? encode_kernel<T, IdT, true> : encode_kernel<T, IdT, false>;
        CU(L, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES));
        {
            // algorithmic bytes: input AABBs + IDs; the record bytes are added when the count is known
            LaunchScope ls(L, BP_K_ENCODE, (double)n * (2 * T::DIM * 4 + sizeof(IdT)));
            kern<<<tiles, ENCODE_THREADS, S::BYTES, L->stream>>>(a);
        }
        TRY(check_launch(L, "encode_kernel"));
        TRY(post_mail(L, 0, L->d_res, (int)(sizeof(ExtendResult) / 8), &L->pending_seq));
        L->pending = true;
        L->pending_base = L->n_records;
        return BP_OK;
    }

    static int extend_device(bp_layer *L, const float *sysb, const float *d_bounds, const void *d_ids, size_t n,
                             const EncodeCount *count = nullptr) {
        if (n == 0) return BP_OK;
        TRY(strip_flags(L));
        if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "too many objects in one extend");
        const uint64_t per_obj = 1ull << T::DIM;
        TRY(ensure_tree(L, std::min<uint64_t>(L->n_records + n * per_obj, MAX_RECORDS)));
        for (int attempt = 0; attempt < 2; ++attempt) {
            TRY(launch_encode(L, sysb, d_bounds, (const IdT *)d_ids, (uint32_t)n, count));
            if (L->min_depth == 0 && L->n_records + n * per_obj <= L->cap_records) return BP_OK; // cannot overflow: stays async
            if (count) return fail(L, BP_ERR_INTERNAL, "counting extend needs min_depth 0 (the counts of a repeated call would add up)");
            // min_depth can push objects below their natural depth: the record count is only
            // known after the kernel, so check it now and redo the call once with enough room
            TRY(wait_mail(L, 0, L->pending_seq, L->h_res, sizeof(ExtendResult)));
            const uint64_t total = L->h_res->total_records;
            if (L->pending_base + total <= L->cap_records) return BP_OK;
            L->pending = false;
            TRY(ensure_tree(L, L->pending_base + total));
        }
        return fail(L, BP_ERR_INTERNAL, "extend did not converge");
    }

    // clear + extend with the per-shard counts taken by the encode kernel itself, then the count row (counts, halo counts
    // and the 7 tag words of the multi-GPU frame, all read from device memory) stored to every rank's matrix: the first
    // host round trip of the frame is the one that fetches the finished matrix.
    static int extend_count_rows(bp_layer *L, const float *sysb, const float *d_bounds, const void *d_ids, size_t n,
                                 const uint64_t *spl, int n_spl, bool allow_fold, const uint64_t *d_rows, int n_rows) {
        EncodeCount ec;
        for (int i = 0; i < ENCODE_MAX_SPLITTERS; ++i) ec.spl[i] = i < n_spl ? spl[i] : ~0ull;
        ec.n_spl = (uint32_t)n_spl;
        ec.cnt = L->d_cnt;
        CU(L, cudaMemsetAsync(L->d_cnt, 0, 32 * sizeof(uint32_t), L->stream));
        if (n == 0) { // no kernel will initialise the result block: an empty tree
            ExtendResult init;
            memset(&init, 0, sizeof init);
            init.key_and = ~0ull;
            init.id_and = ~0ull;
            *L->h_res = init;
            CU(L, cudaMemcpyAsync(L->d_res, L->h_res, sizeof init, cudaMemcpyHostToDevice, L->stream));
        }
        TRY(extend_device(L, sysb, d_bounds, d_ids, n, &ec));
        RowDst rd;
        rd.n = (uint32_t)std::min(n_rows, MAX_ROW_COPIES);
        for (uint32_t i = 0; i < (uint32_t)MAX_ROW_COPIES; ++i) rd.p[i] = i < rd.n ? (uint64_t *)d_rows[i] : nullptr;
        {
            LaunchScope ls(L, BP_K_MISC, 0);
            count_row_result_kernel<<<1, 32, 0, L->stream>>>(L->d_cnt, (uint32_t)n_spl + 1, L->d_res, allow_fold ? 1 : 0,
                                                            (int)(8 * sizeof(IdT)), rd);
        }
        return check_launch(L, "count_row_result_kernel");
    }

    // ---- sort --------------------------------------------------------------------------------------------------
    // Sorts records [off, off + cnt) of the current buffer by (key, id); the result is left in the
    // current buffer.
    // src_k / src_v (whole-tree sorts only): the records are read from these arrays instead of the current buffer; if no
    // pass runs at all they are copied into it.
    static int sort_range(bp_layer *L, uint64_t off, uint64_t cnt, bool need_id_passes, const uint8_t *fold_flags = nullptr,
                          bool *flags_folded = nullptr, const K *src_k = nullptr, const IdT *src_v = nullptr) {
        const int c = L->cur, o = c ^ 1;
        K *k0 = keys(L, c) + off, *k1 = keys(L, o) + off;
        IdT *v0 = ids(L, c) + off, *v1 = ids(L, o) + off;
        const uint64_t kmask = L->key_or & ~L->key_and, imask = L->id_or & ~L->id_and;
        bool in_alt = false;
        int passes = 0, total_passes = 0;
        if (need_id_passes && imask) { // secondary key first (LSD): IDs as the sort key, Index as the payload
            TRY((radix_sort<IdT, K>(L, v0, k0, v1, k1, (uint32_t)cnt, nullptr, imask, BP_K_SORT_HIST, BP_K_SORT_PASS,
                                    &passes, &in_alt, sizeof(K) + sizeof(IdT), nullptr, src_v, src_k)));
            total_passes += passes;
            if (passes) src_k = nullptr, src_v = nullptr; // consumed
            if (in_alt) {
                std::swap(k0, k1);
                std::swap(v0, v1);
            }
        }
        bool in_alt2 = false;
        // (the flags can only ride along when the key passes are the first thing that touches the records)
        const uint8_t *vf = (fold_flags && !(need_id_passes && imask)) ? fold_flags + off : nullptr;
        // Multi-depth scenes have 40-60 varying key bits but only ~log2(cnt) bits of position: radix passes over the top
        // bits, then one finish pass that orders the (tiny) groups of records agreeing on them (plan_top_bits).
        uint64_t top = 0;
        uint32_t gshift = 0;
        bool finish = false;
        if (L->sort_finish_min && cnt >= L->sort_finish_min) {
            if (L->sort_finish_cooldown)
                --L->sort_finish_cooldown; // a recent sort met a group beyond the finish window: plain passes for a while
            else
                finish = plan_top_bits(kmask, cnt, &top, &gshift);
        }
        TRY((radix_sort<K, IdT>(L, k0, v0, k1, v1, (uint32_t)cnt, nullptr, finish ? top : kmask, BP_K_SORT_HIST, BP_K_SORT_PASS,
                                &passes, &in_alt2, sizeof(K) + sizeof(IdT), vf, src_k, src_v)));
        if (flags_folded) *flags_folded = vf != nullptr && passes > 0;
        total_passes += passes;
        if (finish && passes) {
            RecordFinishArgs<K, IdT> fa;
            fa.kin = in_alt2 ? k1 : k0;
            fa.vin = in_alt2 ? v1 : v0;
            fa.kout = in_alt2 ? k0 : k1;
            fa.vout = in_alt2 ? v0 : v1;
            fa.n = (uint32_t)cnt;
            fa.gshift = gshift;
            fa.big = &L->d_tot->pad;
            // what orders a group: the varying bits below gshift, as one bit-field or two around their widest gap
            const uint64_t lowmask = kmask & ((1ull << gshift) - 1ull);
            int f0 = 0, l0 = 0, f1 = 0, l1 = 0;
            if (lowmask) {
                const int lo = __builtin_ctzll(lowmask), hi = 63 - __builtin_clzll(lowmask);
                int gap_at = -1, gap_len = 0;
                for (int b = lo, run = 0; b <= hi; ++b) {
                    run = ((lowmask >> b) & 1ull) ? 0 : run + 1;
                    if (run > gap_len) gap_len = run, gap_at = b - run + 1;
                }
                f0 = lo;
                l0 = (gap_len ? gap_at : hi + 1) - lo;
                if (gap_len) f1 = gap_at + gap_len, l1 = hi + 1 - f1;
            }
            fa.shift0 = (uint32_t)f0;
            fa.bits0 = (uint32_t)l0;
            fa.mask0 = l0 >= 32 ? 0xffffffffu : (1u << l0) - 1u;
            fa.shift1 = (uint32_t)f1;
            fa.mask1 = l1 >= 32 ? 0xffffffffu : (1u << l1) - 1u;
            CU(L, cudaMemsetAsync(L->d_tot, 0, sizeof(ScanTotals), L->stream));
            {
                LaunchScope ls(L, BP_K_SORT_FINISH, 2.0 * (double)cnt * (sizeof(K) + sizeof(IdT)));
                const uint32_t ftiles = (uint32_t)((cnt + RFIN_TILE - 1) / RFIN_TILE);
                static const bool walk = getenv("BP_SORT_FINISH_WALK") && atoi(getenv("BP_SORT_FINISH_WALK")) != 0;
                if (walk)
                    record_finish_walk_kernel<K, IdT><<<ftiles, RFIN_THREADS, 0, L->stream>>>(fa);
                else if (l0 + l1 <= 32)
                    record_finish_kernel<K, IdT, uint32_t><<<ftiles, RFIN_THREADS, 0, L->stream>>>(fa);
                else
                    record_finish_kernel<K, IdT, uint64_t><<<ftiles, RFIN_THREADS, 0, L->stream>>>(fa);
            }
            TRY(check_launch(L, "record_finish_kernel"));
            in_alt2 = !in_alt2;
            TRY(fetch_totals(L));
            if (L->h_tot->pad) {
                // Some group is larger than the finish window (many records in one small region of space): those groups were
                // copied through unchanged, so the buffer still is a stable permutation -- sort it the long way.
                L->sort_finish_cooldown = 32;
                K *fk0 = in_alt2 ? k1 : k0, *fk1 = in_alt2 ? k0 : k1;
                IdT *fv0 = in_alt2 ? v1 : v0, *fv1 = in_alt2 ? v0 : v1;
                bool in_alt3 = false;
                TRY((radix_sort<K, IdT>(L, fk0, fv0, fk1, fv1, (uint32_t)cnt, nullptr, kmask, BP_K_SORT_HIST, BP_K_SORT_PASS,
                                        &passes, &in_alt3, sizeof(K) + sizeof(IdT))));
                total_passes += passes;
                in_alt2 = in_alt2 != in_alt3;
            }
        }
        if (src_k && !passes) { // nothing to sort by: the records still have to arrive in the tree
            CU(L, cudaMemcpyAsync(k0, src_k, cnt * sizeof(K), cudaMemcpyDeviceToDevice, L->stream));
            CU(L, cudaMemcpyAsync(v0, src_v, cnt * sizeof(IdT), cudaMemcpyDeviceToDevice, L->stream));
        }
        const bool final_in_alt = in_alt != in_alt2;
        L->stats.sort_passes = (uint32_t)total_passes;
        if (final_in_alt) {
            if (off == 0 && cnt == L->n_records) {
                L->cur = o; // whole tree: just flip the buffers
            } else {
                CU(L, cudaMemcpyAsync(keys(L, c) + off, keys(L, o) + off, cnt * sizeof(K), cudaMemcpyDeviceToDevice, L->stream));
                CU(L, cudaMemcpyAsync(ids(L, c) + off, ids(L, o) + off, cnt * sizeof(IdT), cudaMemcpyDeviceToDevice, L->stream));
            }
        }
        return BP_OK;
    }

    static int merge_runs(bp_layer *L, uint64_t na, uint64_t nb, bp_layer *B = nullptr) {
        // B: the second run is the (sorted) tree of another layer, read where it is (deferred Layer::merge)
        const int c = L->cur, o = c ^ 1;
        MergeArgs<K, IdT> a;
        a.ka = keys(L, c);
        a.va = ids(L, c);
        a.na = (uint32_t)na;
        a.kb = B ? keys(B, B->cur) : keys(L, c) + na;
        a.vb = B ? ids(B, B->cur) : ids(L, c) + na;
        a.nb = (uint32_t)nb;
        a.kout = keys(L, o);
        a.vout = ids(L, o);
        a.id_mask = L->ids_flagged ? (IdT)((((IdT)1) << (8 * sizeof(IdT) - 3)) - 1) : (IdT) ~(IdT)0;
        if (B && B->stream != L->stream) { // after everything queued on the other layer's stream
            CU(L, cudaEventRecord(B->ev_sync, B->stream));
            CU(L, cudaStreamWaitEvent(L->stream, B->ev_sync, 0));
        }
        const uint32_t tiles = (uint32_t)((na + nb + MERGE_TILE - 1) / MERGE_TILE);
        TRY(ensure(L, L->scratch, (size_t)(tiles + 1) * sizeof(uint32_t)));
        a.partition = (uint32_t *)L->scratch.p;
        const double bytes = 2.0 * (double)(na + nb) * (sizeof(K) + sizeof(IdT));
        {
            LaunchScope ls(L, BP_K_MISC, 0);
            merge_partition_kernel<K, IdT><<<(tiles + 1 + 255) / 256, 256, 0, L->stream>>>(a, tiles);
        }
        TRY(check_launch(L, "merge_partition_kernel"));
        {
            LaunchScope ls(L, BP_K_MERGE, bytes);
            merge_tiles_kernel<K, IdT><<<tiles, MERGE_THREADS, 0, L->stream>>>(a);
        }
        TRY(check_launch(L, "merge_tiles_kernel"));
        if (B && B->stream != L->stream) { // and keep the other layer from touching its tree before the merge has read it
            CU(L, cudaEventRecord(L->ev_sync, L->stream));
            CU(L, cudaStreamWaitEvent(B->stream, L->ev_sync, 0));
        }
        L->cur = o;
        L->stats.merged = 1;
        return BP_OK;
    }

    static int sort(bp_layer *L) {
        if (!L->dirty) return BP_OK;
        const uint64_t R = L->n_records;
        L->stats.sort_passes = 0;
        L->stats.merged = 0;
        uint64_t prefix = std::min(L->prefix, R);
        const uint64_t tail = R - prefix;
        if (R >= 2 && tail > 0) {
            if (prefix == 0) {
                if (!L->tail_sorted) {
                    // the whole tree came from extend(): let the cell flags ride in the IDs' spare top bits.
                    // Normally the first radix pass folds them in while it loads the IDs; if that pass does not
                    // exist (all keys equal) or the ID digits come first, a small kernel does it instead.
                    const int id_bits = 64 - (L->id_or ? __builtin_clzll(L->id_or) : 64);
                    const bool want_flags = L->flags_valid && !L->ids_flagged && id_bits <= (int)(8 * sizeof(IdT)) - 3;
                    const bool id_passes_first = L->tail_nonmono && (L->id_or & ~L->id_and) != 0;
                    auto fold_in_place = [&]() {
                        LaunchScope ls(L, BP_K_MISC, (double)R * (2.0 * sizeof(IdT) + 1.0));
                        const int blocks = (int)std::min<uint64_t>((R + 1023) / 1024, 148 * 8);
                        flags_merge_kernel<IdT><<<blocks, 256, 0, L->stream>>>(ids(L, L->cur), (const uint8_t *)L->cell_flags.p, (uint32_t)R);
                        L->ids_flagged = true;
                    };
                    if (want_flags && id_passes_first) fold_in_place(); // before the ID-digit passes move the records
                    bool folded = false;
                    TRY(sort_range(L, 0, R, L->tail_nonmono,
                                   (want_flags && !id_passes_first) ? (const uint8_t *)L->cell_flags.p : nullptr, &folded));
                    if (want_flags && !id_passes_first) {
                        if (folded)
                            L->ids_flagged = true;
                        else
                            fold_in_place(); // no key pass ran: the records did not move
                    }
                    TRY(check_launch(L, "flags_merge_kernel"));
                }
            } else if (L->lazy_src) {
                // deferred Layer::merge of a sorted layer into this sorted layer: merge straight out of both trees
                bp_layer *B = L->lazy_src;
                TRY(merge_runs(L, prefix, tail, B));
                detach_lazy(L);
            } else if (L->tail_sorted || prefix * 8 >= R) {
                // two sorted runs: one linear merge.  An unsorted tail is radix-sorted on its own first,
                // unless the sorted prefix is so short that re-sorting everything is cheaper.
                if (!L->tail_sorted) TRY(sort_range(L, prefix, tail, L->tail_nonmono));
                TRY(merge_runs(L, prefix, tail));
            } else {
                TRY(sort_range(L, 0, R, true)); // short sorted prefix: cheaper to re-sort everything
            }
        }
        L->dirty = false;
        L->prefix = 0;
        L->tail_sorted = false;
        L->tail_nonmono = false;
        L->tail_has_last = false;
        return BP_OK;
    }

    // ---- scan ---------------------------------------------------------------------------------------------------
    // set_records + sort in one step for records that live in somebody else's buffer (a multi-GPU receive buffer): the first
    // radix pass reads them where they are (no staging copy), the digit plan comes from the caller's masks (no mask pass).
    static int sort_from(bp_layer *L, const void *src_k, const void *src_v, uint64_t n, bool ids_ascending) {
        L->stats.sort_passes = 0;
        L->stats.merged = 0;
        if (n) TRY(sort_range(L, 0, n, !ids_ascending, nullptr, nullptr, (const K *)src_k, (const IdT *)src_v));
        return BP_OK;
    }

    template <int FK, bool DEDUP> static int launch_emit(bp_layer *L, EmitArgs<IdT> &a, uint32_t chunks, double bytes) {
        auto kern = scan_emit_kernel<IdT, FK, T, DEDUP>;
        CU(L, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EmitSmem<IdT>::BYTES));
        LaunchScope ls(L, BP_K_SCAN_EMIT, bytes);
        kern<<<chunks, EMIT_THREADS, EmitSmem<IdT>::BYTES, L->stream>>>(a);
        return BP_OK;
    }
    template <bool DEDUP> static int emit_fk(bp_layer *L, EmitArgs<IdT> &a, int fk, uint32_t chunks, double bytes) {
        switch (fk) {
        case BP_FILTER_NONE: TRY((launch_emit<BP_FILTER_NONE, DEDUP>(L, a, chunks, bytes))); break;
        case BP_FILTER_ID_PARITY: TRY((launch_emit<BP_FILTER_ID_PARITY, DEDUP>(L, a, chunks, bytes))); break;
        case BP_FILTER_XOR_MASK: TRY((launch_emit<BP_FILTER_XOR_MASK, DEDUP>(L, a, chunks, bytes))); break;
        case BP_FILTER_CATEGORY: TRY((launch_emit<BP_FILTER_CATEGORY, DEDUP>(L, a, chunks, bytes))); break;
        case BP_FILTER_SPHERES: TRY((launch_emit<BP_FILTER_SPHERES, DEDUP>(L, a, chunks, bytes))); break;
        default: return fail(L, BP_ERR_INVALID_ARG, "unknown filter kind %d", fk);
        }
        return check_launch(L, "scan_emit_kernel");
    }
    static int emit(bp_layer *L, EmitArgs<IdT> &a, int fk, uint32_t chunks, double bytes, bool dedup) {
        return dedup ? emit_fk<true>(L, a, fk, chunks, bytes) : emit_fk<false>(L, a, fk, chunks, bytes);
    }

    template <int FK, bool DEDUP> static int launch_groups(bp_layer *L, GroupScanArgs<K, IdT> &a, uint32_t tiles, double bytes) {
        LaunchScope ls(L, BP_K_SCAN_EMIT, bytes);
        scan_groups_kernel<K, IdT, FK, DEDUP><<<tiles, GRP_THREADS, 0, L->stream>>>(a);
        return BP_OK;
    }
    template <bool DEDUP> static int groups_fk(bp_layer *L, GroupScanArgs<K, IdT> &a, int fk, uint32_t tiles, double bytes) {
        switch (fk) {
        case BP_FILTER_NONE: TRY((launch_groups<BP_FILTER_NONE, DEDUP>(L, a, tiles, bytes))); break;
        case BP_FILTER_ID_PARITY: TRY((launch_groups<BP_FILTER_ID_PARITY, DEDUP>(L, a, tiles, bytes))); break;
        case BP_FILTER_XOR_MASK: TRY((launch_groups<BP_FILTER_XOR_MASK, DEDUP>(L, a, tiles, bytes))); break;
        case BP_FILTER_CATEGORY: TRY((launch_groups<BP_FILTER_CATEGORY, DEDUP>(L, a, tiles, bytes))); break;
        case BP_FILTER_SPHERES: TRY((launch_groups<BP_FILTER_SPHERES, DEDUP>(L, a, tiles, bytes))); break;
        default: return fail(L, BP_ERR_INVALID_ARG, "unknown filter kind %d", fk);
        }
        return check_launch(L, "scan_groups_kernel");
    }

    // The scan of a tree whose records all sit at one depth (scan_groups_kernel): 0 = done (*out_raw pairs in praw[0]),
    // 1 = take the general path (a same-ID item or a crowded cell turned up), < 0 = -status.
    static int scan_uniform(bp_layer *L, int fk, const FilterArgs &fa, uint64_t *out_raw) {
        constexpr bool wide = sizeof(IdT) == 8;
        const uint64_t R = L->n_records;
        const uint32_t tiles = (uint32_t)((R + GRP_TILE - 1) / GRP_TILE);
        const bool dedup = L->ids_flagged && L->n_halo == 0 && L->scan_dedup;
        // the output is sized from the last scan of this layer (a frame loop: no second launch), else from the tree
        uint64_t cap = std::max<uint64_t>(L->last_raw_pairs + L->last_raw_pairs / 8 + 4096, R / 2 + R / 4 + 4096);
        cap = std::max<uint64_t>(cap, L->praw[0].cap / sizeof(uint64_t));
        for (int attempt = 0; attempt < 2; ++attempt) {
#define BP_TRYN(expr)                  \
    do {                               \
        const int s__ = (expr);        \
        if (s__ != BP_OK) return -s__; \
    } while (0)
            BP_TRYN(ensure(L, L->praw[0], cap * sizeof(uint64_t)));
            if (wide) BP_TRYN(ensure(L, L->praw_b[0], cap * sizeof(uint64_t)));
            BP_TRYN(ensure(L, L->scratch, 64));
            if (cudaMemsetAsync(L->d_tot, 0, sizeof(ScanTotals), L->stream) != cudaSuccess ||
                cudaMemsetAsync(L->scratch.p, 0, 64, L->stream) != cudaSuccess)
                return -fail(L, BP_ERR_CUDA, "cudaMemsetAsync failed");
            GroupScanArgs<K, IdT> ga;
            ga.keys = keys(L, L->cur);
            ga.ids = ids(L, L->cur);
            ga.n = (uint32_t)R;
            ga.id_mask = L->ids_flagged ? (IdT)((((IdT)1) << (8 * sizeof(IdT) - 3)) - 1) : (IdT) ~(IdT)0;
            ga.first_owned = (uint32_t)std::min<uint64_t>(L->n_halo, R);
            ga.out_packed = wide ? nullptr : (uint64_t *)L->praw[0].p;
            ga.out_a = wide ? (uint64_t *)L->praw[0].p : nullptr;
            ga.out_b = wide ? (uint64_t *)L->praw_b[0].p : nullptr;
            ga.capacity = cap;
            ga.pair_counter = (unsigned long long *)L->scratch.p;
            ga.work_counter = (unsigned long long *)L->scratch.p + 1;
            ga.later_count = nullptr;
            L->pair_cnt_n = 0;
            // (the counting sort of the pairs: same rule as the general path, on the capacity instead of the work-item count)
            if (L->want_pair_counts && !wide && L->id_or < (1ull << 22) && cap <= (1ull << 22)) {
                const int id_bits = 64 - (L->id_or ? __builtin_clzll(L->id_or) : 64);
                const uint64_t M = std::max<uint64_t>(4, 1ull << id_bits);
                BP_TRYN(ensure(L, L->pair_cnt, M * sizeof(uint32_t)));
                if (cudaMemsetAsync(L->pair_cnt.p, 0, M * sizeof(uint32_t), L->stream) != cudaSuccess)
                    return -fail(L, BP_ERR_CUDA, "cudaMemsetAsync failed");
                ga.later_count = (uint32_t *)L->pair_cnt.p;
                L->pair_cnt_n = M;
            }
            ga.totals = L->d_tot;
            ga.filter = fa;
            const double bytes = (double)R * (sizeof(K) + sizeof(IdT));
            BP_TRYN(dedup ? groups_fk<true>(L, ga, fk, tiles, bytes) : groups_fk<false>(L, ga, fk, tiles, bytes));
            if (cudaMemcpyAsync(&L->d_tot->n_raw_pairs, L->scratch.p, 8, cudaMemcpyDeviceToDevice, L->stream) != cudaSuccess ||
                cudaMemcpyAsync(&L->d_tot->n_work, (char *)L->scratch.p + 8, 8, cudaMemcpyDeviceToDevice, L->stream) != cudaSuccess)
                return -fail(L, BP_ERR_CUDA, "cudaMemcpyAsync failed");
            BP_TRYN(fetch_totals(L));
#undef BP_TRYN
            if (L->h_tot->pad || L->h_tot->any_same_id) {
                L->pair_cnt_n = 0;
                return 1;
            }
            const uint64_t P = L->h_tot->n_raw_pairs;
            if (P <= cap) {
                L->stats.n_work_items = L->h_tot->n_work;
                L->stats.n_raw_pairs = P;
                L->stats.algo_bytes[BP_K_SCAN_EMIT] += (double)P * 2.0 * sizeof(IdT);
                L->last_raw_pairs = P;
                *out_raw = P;
                return 0;
            }
            if (P > MAX_RECORDS) return -fail(L, BP_ERR_TOO_LARGE, "scan would emit %llu pairs (limit 2^30)", (unsigned long long)P);
            cap = P; // the kernel counted everything it could not write: once more with room for it
        }
        return -fail(L, BP_ERR_INTERNAL, "uniform scan did not converge");
    }

    static int fetch_totals(bp_layer *L) {
        unsigned int seq = 0;
        TRY(post_mail(L, 1, L->d_tot, (int)(sizeof(ScanTotals) / 8), &seq));
        return wait_mail(L, 1, seq, L->h_tot, sizeof(ScanTotals));
    }

    // Everything up to the raw (unsorted, duplicate-carrying) pairs, left in praw[0] (+ praw_b[0]).
    static int scan_raw(bp_layer *L, const bp_filter *f, uint64_t *out_raw) {
        *out_raw = 0;
        L->pair_cnt_n = 0;
        TRY(sort(L));
        L->n_pairs = 0;
        L->stats.n_work_items = L->stats.n_raw_pairs = L->stats.n_pairs = 0;
        L->stats.pair_sort_passes = 0;
        L->stats.rescans = 0;
        L->n_invalid = 0; // `self.invalid.clear()` -- src/layer.rs:468, :502
        const uint64_t R = L->n_records;
        if (R < 2) return BP_OK;
        const int fk = f ? f->kind : BP_FILTER_NONE;
        FilterArgs fa;
        fa.arg = f ? f->arg : 0;
        fa.table = nullptr;
        fa.n_table = 0;
        if (fk == BP_FILTER_CATEGORY || fk == BP_FILTER_SPHERES) {
            if (!f->table && f->n_table) return fail(L, BP_ERR_INVALID_ARG, "table filter without a table");
            const size_t row = fk == BP_FILTER_CATEGORY ? 8 : 16; // {cat, msk} or {x, y, z, r}
            fa.n_table = f->n_table;
            if (f->table_on_device) {
                if (fk == BP_FILTER_SPHERES && ((uintptr_t)f->table & 15u)) return fail(L, BP_ERR_INVALID_ARG, "the sphere table must be 16-byte aligned");
                fa.table = f->table;
            } else if (f->n_table) {
                TRY(ensure(L, L->filter_table, f->n_table * row));
                CU(L, cudaMemcpyAsync(L->filter_table.p, f->table, f->n_table * row, cudaMemcpyHostToDevice, L->stream));
                fa.table = (const uint32_t *)L->filter_table.p;
            }
        }

        // ---- every record at the same depth (no depth bit differs between two keys): one fused kernel ----
        {
            constexpr uint64_t DEPTH_FIELD = (1ull << T::DEPTH_BITS) - 1ull;
            static const bool off = getenv("BP_SCAN_UNIFORM") && atoi(getenv("BP_SCAN_UNIFORM")) == 0; // tuning aid
            if (!off && ((L->key_or & ~L->key_and) & DEPTH_FIELD) == 0) {
                const int r = scan_uniform(L, fk, fa, out_raw);
                if (r < 0) return -r;
                if (r == 0) return BP_OK;
            }
        }

        // ---- runs ----
        const uint32_t rtiles = (uint32_t)((R + RUNS_TILE - 1) / RUNS_TILE);
        TRY(ensure(L, L->src_idx, R * sizeof(uint32_t)));
        TRY(ensure(L, L->src_off, (R + 1) * sizeof(uint64_t)));
        static_assert(sizeof(ScanTotals) % 8 == 0, "the counters behind the totals must stay 8-byte aligned");
        char *runs_ctr = (char *)(L->d_tot + 1); // tile counter + packed (sources, work) counter, zeroed with the totals
        CU(L, cudaMemsetAsync(L->d_tot, 0, sizeof(ScanTotals) + 64, L->stream));
        RunsArgs<T> ra;
        ra.keys = keys(L, L->cur);
        ra.n = (uint32_t)R;
        ra.src_idx = (uint32_t *)L->src_idx.p;
        ra.src_off = (uint64_t *)L->src_off.p;
        ra.tile_counter = (uint32_t *)runs_ctr;
        ra.packed_counter = (unsigned long long *)(runs_ctr + 8);
        ra.totals = L->d_tot;
        ra.err = L->d_err;
        {
            auto kern = scan_runs_kernel<T>;
            CU(L, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RunsSmem<T>::BYTES));
            LaunchScope ls(L, BP_K_SCAN_RUNS, (double)R * sizeof(K));
            kern<<<rtiles, RUNS_THREADS, RunsSmem<T>::BYTES, L->stream>>>(ra);
        }
        TRY(check_launch(L, "scan_runs_kernel"));
        TRY(fetch_totals(L));
        if (L->h_tot->pad) return fail(L, BP_ERR_TOO_LARGE, "scan would visit more than 2^33 record pairs");
        const uint64_t C = L->h_tot->n_sources, W = L->h_tot->n_work;
        L->stats.n_work_items = W;
        L->stats.algo_bytes[BP_K_SCAN_RUNS] += (double)C * 12.0;
        if (W == 0) return BP_OK;
        if (W > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "scan would visit %llu record pairs (limit 2^30)", (unsigned long long)W);

        // ---- chunk partition + emit ----
        const uint32_t chunks = (uint32_t)((W + EMIT_CHUNK - 1) / EMIT_CHUNK);
        const bool wide = sizeof(IdT) == 8;
        TRY(ensure(L, L->chunk_src, (size_t)chunks * sizeof(uint32_t)));
        for (int i = 0; i < 2; ++i) {
            TRY(ensure(L, L->praw[i], W * sizeof(uint64_t)));
            if (wide) TRY(ensure(L, L->praw_b[i], W * sizeof(uint64_t)));
        }
        TRY(ensure(L, L->pout, W * 2 * sizeof(IdT)));
        {
            LaunchScope ls(L, BP_K_MISC, 0);
            scan_chunks_kernel<<<(uint32_t)((C + 255) / 256), 256, 0, L->stream>>>((const uint64_t *)L->src_off.p, (uint32_t)C,
                                                                                  (uint32_t *)L->chunk_src.p);
        }
        TRY(check_launch(L, "scan_chunks_kernel"));
        const size_t ebytes = 64;
        TRY(ensure(L, L->scratch, ebytes));
        EmitArgs<IdT> ea;
        ea.ids = ids(L, L->cur);
        ea.keys = keys(L, L->cur);
        ea.id_mask = L->ids_flagged ? (IdT)((((IdT)1) << (8 * sizeof(IdT) - 3)) - 1) : (IdT) ~(IdT)0;
        const bool dedup = L->ids_flagged && L->n_halo == 0 && L->scan_dedup;
        ea.src_idx = (const uint32_t *)L->src_idx.p;
        ea.src_off = (const uint64_t *)L->src_off.p;
        ea.chunk_src = (const uint32_t *)L->chunk_src.p;
        ea.n_work = W;
        ea.n_sources = (uint32_t)C;
        ea.first_owned = (uint32_t)std::min<uint64_t>(L->n_halo, R);
        ea.inactive = nullptr;
        ea.out_packed = wide ? nullptr : (uint64_t *)L->praw[0].p;
        ea.out_a = wide ? (uint64_t *)L->praw[0].p : nullptr;
        ea.out_b = wide ? (uint64_t *)L->praw_b[0].p : nullptr;
        ea.capacity = W;
        ea.pair_counter = (unsigned long long *)L->scratch.p;
        // dense 32-bit IDs: the emission also counts the pairs of every later ID, and finish_pairs replaces the
        // radix passes over the later ID by one scan over the ID range + one scatter (see pair_scatter_kernel)
        L->pair_cnt_n = 0;
        ea.later_count = nullptr;
        // Only while the pairs and the count array stay resident in the 126 MB L2: the scatter writes 8 bytes at
        // random places, which costs a DRAM read-modify-write per pair once they do not (measured at config 3,
        // 13.7M pairs: 0.54 ms against 0.30 ms for the three radix passes; at config 2, 2.1M pairs: 0.03 against 0.08).
        if (L->want_pair_counts && !wide && L->id_or < (1ull << 22) && W <= (1ull << 22)) {
            const int id_bits = 64 - (L->id_or ? __builtin_clzll(L->id_or) : 64);
            const uint64_t M = std::max<uint64_t>(4, 1ull << id_bits);
            {
                TRY(ensure(L, L->pair_cnt, M * sizeof(uint32_t)));
                CU(L, cudaMemsetAsync(L->pair_cnt.p, 0, M * sizeof(uint32_t), L->stream));
                ea.later_count = (uint32_t *)L->pair_cnt.p;
                L->pair_cnt_n = M;
            }
        }
        ea.totals = L->d_tot;
        ea.filter = fa;
        ea.err = L->d_err;
        // algorithmic bytes of one emission: both IDs of every work item + the pair it writes
        const double emit_bytes = (double)W * (2.0 * sizeof(IdT) + 2.0 * sizeof(IdT)) + (double)C * 12.0;
        CU(L, cudaMemsetAsync(L->scratch.p, 0, ebytes, L->stream));
        ea.mode = EMIT_MODE_FIRST;
        TRY(emit(L, ea, fk, chunks, emit_bytes + (dedup ? (double)W * 2.0 * sizeof(K) : 0.0), dedup));
        CU(L, cudaMemcpyAsync(&L->d_tot->n_raw_pairs, L->scratch.p, 8, cudaMemcpyDeviceToDevice, L->stream));
        TRY(fetch_totals(L));
        if (L->h_tot->any_same_id) {
            // some ID owns nested bounds: the reference skips such records entirely (src/layer.rs:562-564),
            // so mark them and emit again without them
            L->stats.rescans = 1;
            TRY(ensure(L, L->inactive, R));
            CU(L, cudaMemsetAsync(L->inactive.p, 0, R, L->stream));
            ea.inactive = (unsigned char *)L->inactive.p;
            CU(L, cudaMemsetAsync(L->scratch.p, 0, ebytes, L->stream));
            ea.mode = EMIT_MODE_FLAG;
            TRY(emit(L, ea, fk, chunks, (double)W * 2.0 * sizeof(IdT), false));
            CU(L, cudaMemsetAsync(L->scratch.p, 0, ebytes, L->stream));
            if (ea.later_count) CU(L, cudaMemsetAsync(L->pair_cnt.p, 0, L->pair_cnt_n * sizeof(uint32_t), L->stream));
            ea.mode = EMIT_MODE_ACTIVE;
            TRY(emit(L, ea, fk, chunks, emit_bytes, false));
            CU(L, cudaMemcpyAsync(&L->d_tot->n_raw_pairs, L->scratch.p, 8, cudaMemcpyDeviceToDevice, L->stream));
            TRY(fetch_totals(L));
        }
        L->stats.n_raw_pairs = L->h_tot->n_raw_pairs;
        *out_raw = L->h_tot->n_raw_pairs;
        return BP_OK;
    }

    // Sorts the P_raw raw pairs in praw[0] (+ praw_b[0]) and removes duplicates into pout.
    static int finish_pairs(bp_layer *L, uint64_t P_raw) {
        constexpr bool wide = sizeof(IdT) == 8;
        L->n_pairs = 0;
        L->stats.n_pairs = 0;
        L->stats.pair_sort_passes = 0;
        if (P_raw == 0) return BP_OK;
        TRY(ensure(L, L->pout, P_raw * 2 * sizeof(IdT)));
        TRY(ensure(L, L->praw[1], P_raw * sizeof(uint64_t)));
        if (wide) TRY(ensure(L, L->praw_b[1], P_raw * sizeof(uint64_t)));

        // ---- sort the raw pairs (src/layer.rs:473, :516) ----
        // Fast path: radix passes over the LATER ID's bits only, then pair_finish_kernel orders and
        // deduplicates the (tiny) groups of equal later IDs.  Fallback (a group larger than the finish
        // kernel's window): the remaining passes over all bits, then pair_unique_kernel.
        const uint64_t imask = L->id_or & ~L->id_and;
        uint64_t *a0 = (uint64_t *)(L->pairs_src ? L->pairs_src : L->praw[0].p), *a1 = (uint64_t *)L->praw[1].p;
        uint64_t *b0 = wide ? (uint64_t *)L->praw_b[0].p : nullptr, *b1 = wide ? (uint64_t *)L->praw_b[1].p : nullptr;
        int passes = 0, total_passes = 0;
        bool in_alt = false;
        uint32_t fin_gshift = 32; // pair_finish_kernel: one group per later ID
        if (L->pairs_grouped) {
            // (query, ID) pairs of a batched query: written query by query, only the order inside a group is open
        } else if (!wide && L->pair_cnt_n) {
            // counting sort by the later ID: offsets = exclusive scan of the per-ID counts, then one scatter
            const uint32_t M = (uint32_t)L->pair_cnt_n;
            const uint32_t ctiles = (M + CSCAN_TILE - 1) / CSCAN_TILE;
            const size_t cbytes = 64 + (size_t)ctiles * sizeof(uint64_t);
            TRY(ensure(L, L->scratch, cbytes));
            CU(L, cudaMemsetAsync(L->scratch.p, 0, cbytes, L->stream));
            {
                LaunchScope ls(L, BP_K_PAIR_HIST, 2.0 * (double)M * sizeof(uint32_t));
                count_scan_kernel<<<ctiles, CSCAN_THREADS, 0, L->stream>>>((uint32_t *)L->pair_cnt.p, M, (uint64_t *)((char *)L->scratch.p + 64),
                                                                         (uint32_t *)L->scratch.p, L->d_err);
            }
            TRY(check_launch(L, "count_scan_kernel"));
            {
                LaunchScope ls(L, BP_K_PAIR_PASS, 2.0 * (double)P_raw * sizeof(uint64_t));
                const int blocks = (int)std::min<uint64_t>((P_raw + 1023) / 1024, 148 * 8);
                pair_scatter_kernel<<<blocks, 256, 0, L->stream>>>(a0, (uint32_t)P_raw, (uint32_t *)L->pair_cnt.p, a1);
            }
            TRY(check_launch(L, "pair_scatter_kernel"));
            L->pair_cnt_n = 0; // consumed
            in_alt = true;
        } else if (!wide) {
            // The finish kernel orders whole packed pairs inside a group, so the lowest bits of the later ID need no radix
            // pass of their own when leaving them out saves one: a group then holds the pairs of 2^drop later IDs -- taken
            // while that is expected to be at most ~6 pairs (2^25 IDs, 49 M pairs: 3 passes over 24 bits instead of 4).
            uint64_t lmask = imask & 0xffffffffull & ~L->pair_later_fixed;
            {
                RadixPlan pl;
                memset(&pl, 0, sizeof pl);
                const int np_all = plan_passes(lmask << 32, pl, 0, 8);
                const int nb = __builtin_popcountll(lmask);
                uint64_t m = lmask;
                for (int drop = 1; drop <= 3 && drop < nb && np_all >= 2; ++drop) {
                    m &= m - 1; // without its lowest varying bit
                    static const double max_group = getenv("BP_PAIR_DROP_MAX_GROUP") ? atof(getenv("BP_PAIR_DROP_MAX_GROUP")) : 6.0; // tuning aid
                    if ((double)P_raw * (double)(1u << drop) > max_group * (double)(1ull << nb)) break;
                    memset(&pl, 0, sizeof pl);
                    if (plan_passes(m << 32, pl, 0, 8) < np_all) {
                        lmask = m;
                        fin_gshift = 32u + (uint32_t)__builtin_ctzll(m);
                        break;
                    }
                }
            }
            TRY((radix_sort<uint64_t, NoVal>(L, a0, (NoVal *)nullptr, a1, (NoVal *)nullptr, (uint32_t)P_raw, nullptr, lmask << 32,
                                             BP_K_PAIR_HIST, BP_K_PAIR_PASS, &passes, &in_alt, 8)));
        } else {
            TRY((radix_sort<uint64_t, uint64_t>(L, a0, b0, a1, b1, (uint32_t)P_raw, nullptr, imask, BP_K_PAIR_HIST, BP_K_PAIR_PASS,
                                                &passes, &in_alt, 16)));
        }
        total_passes = passes;
        if (in_alt) {
            std::swap(a0, a1);
            std::swap(b0, b1);
        }
        {
            const uint32_t ftiles = (uint32_t)((P_raw + FinCfg<IdT>::TILE - 1) / FinCfg<IdT>::TILE);
            const size_t fbytes = 64 + (size_t)ftiles * sizeof(uint64_t);
            TRY(ensure(L, L->scratch, fbytes));
            CU(L, cudaMemsetAsync(L->scratch.p, 0, fbytes, L->stream)); // (d_tot->pad is still zero: every caller cleared *d_tot)
            FinishArgs<IdT> fa;
            fa.in_packed = wide ? nullptr : a0;
            fa.in_a = wide ? a0 : nullptr;
            fa.in_b = b0;
            fa.n = (uint32_t)P_raw;
            fa.gshift = fin_gshift;
            fa.out = (IdT *)L->pout.p;
            fa.tile_counter = (uint32_t *)L->scratch.p;
            fa.status = (uint64_t *)((char *)L->scratch.p + 64);
            fa.totals = L->d_tot;
            fa.err = L->d_err;
            {
                LaunchScope ls(L, BP_K_PAIR_UNIQUE, (double)P_raw * 2.0 * sizeof(IdT));
                if constexpr (!wide) {
                    if (fin_gshift != 32)
                        pair_finish_kernel<IdT, true><<<ftiles, FIN_THREADS, 0, L->stream>>>(fa);
                    else
                        pair_finish_kernel<IdT, false><<<ftiles, FIN_THREADS, 0, L->stream>>>(fa);
                } else {
                    pair_finish_kernel<IdT, false><<<ftiles, FIN_THREADS, 0, L->stream>>>(fa);
                }
            }
            TRY(check_launch(L, "pair_finish_kernel"));
            TRY(fetch_totals(L));
        }
        if (L->h_tot->pad) { // some later ID has more partners than the finish window: full-width sort instead
            uint64_t *sorted_a = nullptr, *sorted_b = nullptr;
            if (!wide) {
                const uint64_t pmask = (imask << 32) | (imask & 0xffffffffull);
                TRY((radix_sort<uint64_t, NoVal>(L, a0, (NoVal *)nullptr, a1, (NoVal *)nullptr, (uint32_t)P_raw, nullptr, pmask,
                                                 BP_K_PAIR_HIST, BP_K_PAIR_PASS, &passes, &in_alt, 8)));
                total_passes += passes;
                sorted_a = in_alt ? a1 : a0;
            } else {
                TRY((radix_sort<uint64_t, uint64_t>(L, b0, a0, b1, a1, (uint32_t)P_raw, nullptr, imask, BP_K_PAIR_HIST,
                                                    BP_K_PAIR_PASS, &passes, &in_alt, 16)));
                total_passes += passes;
                if (in_alt) {
                    std::swap(a0, a1);
                    std::swap(b0, b1);
                }
                TRY((radix_sort<uint64_t, uint64_t>(L, a0, b0, a1, b1, (uint32_t)P_raw, nullptr, imask, BP_K_PAIR_HIST,
                                                    BP_K_PAIR_PASS, &passes, &in_alt, 16)));
                total_passes += passes;
                if (in_alt) {
                    std::swap(a0, a1);
                    std::swap(b0, b1);
                }
                sorted_a = a0;
                sorted_b = b0;
            }
            const uint32_t utiles = (uint32_t)((P_raw + UNIQ_TILE - 1) / UNIQ_TILE);
            const size_t ubytes = 64 + (size_t)utiles * sizeof(uint64_t);
            TRY(ensure(L, L->scratch, ubytes));
            CU(L, cudaMemsetAsync(L->scratch.p, 0, ubytes, L->stream));
            UniqueArgs<IdT> ua;
            ua.in_packed = wide ? nullptr : sorted_a;
            ua.in_a = wide ? sorted_a : nullptr;
            ua.in_b = sorted_b;
            ua.n_host = (uint32_t)P_raw;
            ua.n_dev = nullptr;
            ua.out = (IdT *)L->pout.p;
            ua.tile_counter = (uint32_t *)L->scratch.p;
            ua.status = (uint64_t *)((char *)L->scratch.p + 64);
            ua.totals = L->d_tot;
            ua.err = L->d_err;
            {
                LaunchScope ls(L, BP_K_PAIR_UNIQUE, (double)P_raw * 2.0 * sizeof(IdT));
                pair_unique_kernel<IdT><<<utiles, UNIQ_THREADS, 0, L->stream>>>(ua);
            }
            TRY(check_launch(L, "pair_unique_kernel"));
            TRY(fetch_totals(L));
        }
        L->stats.pair_sort_passes = (uint32_t)total_passes;
        L->n_pairs = L->h_tot->n_pairs;
        L->stats.n_pairs = L->n_pairs;
        L->stats.algo_bytes[BP_K_PAIR_UNIQUE] += (double)L->n_pairs * 2.0 * sizeof(IdT);
        return BP_OK;
    }

    static int scan(bp_layer *L, const bp_filter *f) {
        uint64_t P_raw = 0;
        L->want_pair_counts = true;
        const int st = scan_raw(L, f, &P_raw);
        L->want_pair_counts = false;
        TRY(st);
        return finish_pairs(L, P_raw);
    }

    // ---- batched queries (Layer::test_box / test_ray, src/layer.rs:244-351) --------------------------------
    // d_params: n_queries x Geom::PARAMS floats on the device.  Leaves the sorted, duplicate-free (query, id)
    // pairs in pout (n_pairs of them) and the CSR offsets of every query in query_offsets.
    template <class Geom> static int query(bp_layer *L, const float *sysb, const float *d_params, size_t nq, int max_depth) {
        TRY(sort(L)); // `self.sort()` -- src/layer.rs:262
        L->n_pairs = 0;
        TRY(ensure(L, L->query_offsets, (nq + 1) * sizeof(uint32_t)));
        CU(L, cudaMemsetAsync(L->query_offsets.p, 0, (nq + 1) * sizeof(uint32_t), L->stream));
        const uint64_t R = L->n_records;
        if (nq == 0 || R == 0) return BP_OK;
        const bool wide = sizeof(IdT) == 8;
        const size_t ncnt = (nq + 3) & ~(size_t)3; // count_scan_kernel works on whole uint4
        TRY(ensure(L, L->query_counts, ncnt * sizeof(uint32_t)));
        CU(L, cudaMemsetAsync(L->query_counts.p, 0, ncnt * sizeof(uint32_t), L->stream));
        const uint32_t ctiles = (uint32_t)((ncnt + CSCAN_TILE - 1) / CSCAN_TILE);
        const size_t sbytes = 64 + (size_t)ctiles * sizeof(uint64_t);
        TRY(ensure(L, L->scratch, sbytes));
        CU(L, cudaMemsetAsync(L->scratch.p, 0, sbytes, L->stream));
        QueryArgs<T, IdT> qa;
        qa.keys = keys(L, L->cur);
        qa.ids = ids(L, L->cur);
        qa.id_mask = L->ids_flagged ? (IdT)((((IdT)1) << (8 * sizeof(IdT) - 3)) - 1) : (IdT) ~(IdT)0;
        qa.n = (uint32_t)R;
        qa.params = d_params;
        qa.n_queries = (uint32_t)nq;
        for (int i = 0; i < 6; ++i) qa.sysb[i] = i < 2 * T::DIM ? sysb[i] : 0.f;
        qa.max_depth = max_depth;
        qa.counts = (uint32_t *)L->query_counts.p;
        qa.out_packed = qa.out_a = qa.out_b = nullptr;
        qa.total = (unsigned long long *)L->scratch.p;
        qa.err = L->d_err;
        const int blocks = (int)std::min<size_t>((nq + QUERY_WARPS - 1) / QUERY_WARPS, 148 * 16);
        {
            LaunchScope ls(L, BP_K_QUERY, (double)nq * Geom::PARAMS * 4);
            query_kernel<T, IdT, Geom, true><<<blocks, QUERY_THREADS, 0, L->stream>>>(qa);
        }
        TRY(check_launch(L, "query_kernel<count>"));
        CU(L, cudaMemcpyAsync(&L->d_tot->n_work, L->scratch.p, 8, cudaMemcpyDeviceToDevice, L->stream));
        {
            LaunchScope ls(L, BP_K_MISC, 2.0 * (double)ncnt * sizeof(uint32_t));
            count_scan_kernel<<<ctiles, CSCAN_THREADS, 0, L->stream>>>((uint32_t *)L->query_counts.p, (uint32_t)ncnt,
                                                                     (uint64_t *)((char *)L->scratch.p + 64), (uint32_t *)L->scratch.p + 2, L->d_err);
        }
        TRY(check_launch(L, "count_scan_kernel"));
        TRY(fetch_totals(L));
        const uint64_t P_raw = L->h_tot->n_work;
        if (P_raw > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "the queries report %llu records (limit 2^30)", (unsigned long long)P_raw);
        if (P_raw == 0) return BP_OK;
        TRY(ensure(L, L->praw[0], P_raw * sizeof(uint64_t)));
        if (wide) TRY(ensure(L, L->praw_b[0], P_raw * sizeof(uint64_t)));
        qa.out_packed = wide ? nullptr : (uint64_t *)L->praw[0].p;
        qa.out_a = wide ? (uint64_t *)L->praw[0].p : nullptr;
        qa.out_b = wide ? (uint64_t *)L->praw_b[0].p : nullptr;
        {
            LaunchScope ls(L, BP_K_QUERY, (double)P_raw * (sizeof(IdT) + 2.0 * sizeof(IdT)));
            query_kernel<T, IdT, Geom, false><<<blocks, QUERY_THREADS, 0, L->stream>>>(qa);
        }
        TRY(check_launch(L, "query_kernel<write>"));
        // `results.sort(); results.dedup()` (src/layer.rs:276-277) for every query: the pairs are grouped by query
        // already; the digit plan of the (rare) full-width fallback has to cover the query numbers as well
        const uint64_t keep_or = L->id_or, keep_and = L->id_and;
        uint64_t qbits = 0;
        while ((qbits + 1) < nq) qbits = (qbits << 1) | 1;
        L->id_or |= qbits;
        L->id_and = 0;
        L->pairs_grouped = true;
        CU(L, cudaMemsetAsync(L->d_tot, 0, sizeof(ScanTotals), L->stream));
        const int st = finish_pairs(L, P_raw);
        L->pairs_grouped = false;
        L->id_or = keep_or;
        L->id_and = keep_and;
        TRY(st);
        {
            LaunchScope ls(L, BP_K_MISC, (double)L->n_pairs * sizeof(IdT) + (double)nq * sizeof(uint32_t));
            const uint32_t np = (uint32_t)L->n_pairs;
            query_offsets_kernel<IdT><<<(np + 1 + 255) / 256, 256, 0, L->stream>>>((const IdT *)L->pout.p, np, (uint32_t)nq,
                                                                                 (uint32_t *)L->query_offsets.p);
        }
        return check_launch(L, "query_offsets_kernel");
    }
    // Layer::pick_ray -- src/layer.rs:424-446, batched; results in pick_out (one PickResult per ray)
    template <class Shape> static int pick_launch(bp_layer *L, PickArgs<T, IdT> &pa, int blocks) {
        LaunchScope ls(L, BP_K_QUERY, (double)pa.n_queries * (2.0 * T::DIM * 4 + sizeof(PickResult)));
        pick_kernel<T, IdT, Shape><<<blocks, QUERY_THREADS, 0, L->stream>>>(pa);
        return BP_OK;
    }
    static int pick(bp_layer *L, const float *sysb, const float *d_rays, size_t nq, float max_dist, int max_depth, int shape_kind,
                    const float *d_shapes, size_t n_shapes) {
        TRY(sort(L)); // `self.sort()` -- src/layer.rs:375
        TRY(ensure(L, L->pick_out, std::max<size_t>(nq, 1) * sizeof(PickResult)));
        if (nq == 0) return BP_OK;
        PickArgs<T, IdT> pa;
        pa.keys = keys(L, L->cur);
        pa.ids = ids(L, L->cur);
        pa.id_mask = L->ids_flagged ? (IdT)((((IdT)1) << (8 * sizeof(IdT) - 3)) - 1) : (IdT) ~(IdT)0;
        pa.n = (uint32_t)L->n_records;
        pa.rays = d_rays;
        pa.n_queries = (uint32_t)nq;
        for (int i = 0; i < 6; ++i) pa.sysb[i] = i < 2 * T::DIM ? sysb[i] : 0.f;
        pa.max_dist = max_dist;
        pa.max_depth = max_depth;
        pa.shapes = d_shapes;
        pa.n_shapes = n_shapes;
        pa.out = (PickResult *)L->pick_out.p;
        pa.err = L->d_err;
        const int blocks = (int)std::min<size_t>((nq + QUERY_WARPS - 1) / QUERY_WARPS, 148 * 16);
        if (shape_kind == BP_PICK_SPHERE)
            TRY(pick_launch<PickSphere<T::DIM>>(L, pa, blocks));
        else if (shape_kind == BP_PICK_AABB)
            TRY(pick_launch<PickAabb<T::DIM>>(L, pa, blocks));
        else
            return fail(L, BP_ERR_INVALID_ARG, "unknown pick shape kind %d", shape_kind);
        return check_launch(L, "pick_kernel");
    }

    static int query_kind(bp_layer *L, int ray, const float *sysb, const float *d_params, size_t nq, int max_depth) {
        return ray ? query<RayTestGeom<T::DIM>>(L, sysb, d_params, nq, max_depth) : query<BoxTestGeom<T::DIM>>(L, sysb, d_params, nq, max_depth);
    }

    // ---- multi-GPU building blocks ---------------------------------------------------------------------
    // Stable range partition of (key, payload) by `n_spl` splitters on (key >> shift): the same onesweep
    // pass as the sort, with a splitter-search digit.  counts_out[b] = elements in bucket b.
    template <class PK, class PV>
    static int partition(bp_layer *L, const PK *kin, const PV *vin, uint32_t n, const uint64_t *spl, int n_spl, uint32_t shift,
                         PK *kout, PV *vout, uint64_t *counts_out) {
        typedef PassTune<PK, PV> Tune;
        typedef RadixPassCfg<PK, PV, Tune::THREADS, Tune::ITEMS> Cfg;
        for (int b = 0; b <= n_spl; ++b) counts_out[b] = 0;
        if (n == 0) return BP_OK;
        SplitterDigit<PK> op;
        for (int i = 0; i < MAX_SPLITTERS; ++i) op.spl[i] = i < n_spl ? spl[i] : ~0ull;
        op.n = (uint32_t)n_spl;
        op.shift = shift;
        const uint32_t tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
        const size_t total = (size_t)(RADIX + 64 + (size_t)tiles * RADIX) * sizeof(uint32_t);
        TRY(ensure(L, L->scratch, total));
        uint32_t *hist = (uint32_t *)L->scratch.p, *counters = hist + RADIX, *status = counters + 64;
        CU(L, cudaMemsetAsync(L->scratch.p, 0, total, L->stream));
        {
            LaunchScope ls(L, BP_K_PARTITION, (double)n * sizeof(PK));
            const int blocks = (int)std::min<size_t>((n + 4095) / 4096, 148 * 8);
            partition_hist_kernel<PK, T, false><<<std::max(blocks, 1), 512, 0, L->stream>>>(kin, n, op, hist, nullptr);
        }
        TRY(check_launch(L, "partition_hist_kernel"));
        // bucket counts back to the host (they are the all-to-all split sizes) before the scan overwrites them
        uint32_t h_counts[RADIX];
        CU(L, cudaMemcpyAsync(h_counts, hist, (n_spl + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, L->stream));
        {
            LaunchScope ls(L, BP_K_MISC, 0);
            radix_scan_hist_kernel<<<1, RADIX, 0, L->stream>>>(hist);
        }
        TRY(check_launch(L, "radix_scan_hist_kernel"));
        auto kern = radix_pass_kernel<PK, PV, Tune::THREADS, Tune::ITEMS, Tune::MINB, SplitterDigit<PK>>;
        CU(L, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
        RadixPassArgs<PK, PV, SplitterDigit<PK>> a;
        a.kin = kin;
        a.kout = kout;
        a.vin = vin;
        a.vout = vout;
        a.n_host = n;
        a.n_dev = nullptr;
        a.ghist_excl = hist;
        a.status = status;
        a.tile_counter = counters;
        a.vflags = nullptr;
        a.op = op;
        a.err = L->d_err;
        {
            const double eb = std::is_same<PV, NoVal>::value ? sizeof(PK) : sizeof(PK) + sizeof(PV);
            LaunchScope ls(L, BP_K_PARTITION, 2.0 * (double)n * eb);
            kern<<<tiles, Tune::THREADS, Cfg::SMEM_BYTES, L->stream>>>(a);
        }
        TRY(check_launch(L, "radix_pass_kernel<splitters>"));
        CU(L, cudaStreamSynchronize(L->stream));
        for (int b = 0; b <= n_spl; ++b) counts_out[b] = h_counts[b];
        return BP_OK;
    }

    // Bucket sizes of a splitter partition (and, for records, the halo copies every bucket will receive).
    template <class PK, bool HALO>
    static int partition_count(bp_layer *L, const PK *kin, uint32_t n, const uint64_t *spl, int n_spl, uint32_t shift,
                               uint64_t *counts_out, uint64_t *halo_out, const uint64_t *d_rows = nullptr, int n_rows = 0,
                               const uint64_t *tags = nullptr, int n_tags = 0) {
        if (d_rows) { // counts stay on the device (a row of a peer-visible matrix): no host round trip
            SplitterDigit<PK> op;
            for (int i = 0; i < MAX_SPLITTERS; ++i) op.spl[i] = i < n_spl ? spl[i] : ~0ull;
            op.n = (uint32_t)n_spl;
            op.shift = shift;
            TRY(ensure(L, L->scratch, 256));
            uint32_t *hist = (uint32_t *)L->scratch.p, *halo = hist + 32;
            CU(L, cudaMemsetAsync(L->scratch.p, 0, 256, L->stream));
            if (n) {
                LaunchScope ls(L, BP_K_PARTITION, (double)n * sizeof(PK));
                const int blocks = (int)std::min<size_t>((n + 4095) / 4096, 148 * 8);
                partition_hist_kernel<PK, T, HALO><<<std::max(blocks, 1), 512, 0, L->stream>>>(kin, n, op, hist, halo);
            }
            TRY(check_launch(L, "partition_hist_kernel"));
            {
                LaunchScope ls(L, BP_K_MISC, 0);
                RowTags rt;
                rt.n = (uint32_t)std::min(n_tags, MAX_ROW_TAGS);
                for (uint32_t i = 0; i < (uint32_t)MAX_ROW_TAGS; ++i) rt.v[i] = i < rt.n ? tags[i] : 0;
                RowDst rd;
                rd.n = (uint32_t)std::min(n_rows, MAX_ROW_COPIES);
                for (uint32_t i = 0; i < (uint32_t)MAX_ROW_COPIES; ++i) rd.p[i] = i < rd.n ? (uint64_t *)d_rows[i] : nullptr;
                count_row_kernel<<<1, 32, 0, L->stream>>>(hist, HALO ? halo : nullptr, (uint32_t)n_spl + 1, rt, rd);
            }
            return check_launch(L, "count_row_kernel");
        }
        for (int b = 0; b <= n_spl; ++b) {
            counts_out[b] = 0;
            if (halo_out) halo_out[b] = 0;
        }
        if (n == 0) return BP_OK;
        SplitterDigit<PK> op;
        for (int i = 0; i < MAX_SPLITTERS; ++i) op.spl[i] = i < n_spl ? spl[i] : ~0ull;
        op.n = (uint32_t)n_spl;
        op.shift = shift;
        TRY(ensure(L, L->scratch, 256));
        uint32_t *hist = (uint32_t *)L->scratch.p, *halo = hist + 32;
        CU(L, cudaMemsetAsync(L->scratch.p, 0, 256, L->stream));
        {
            LaunchScope ls(L, BP_K_PARTITION, (double)n * sizeof(PK));
            const int blocks = (int)std::min<size_t>((n + 4095) / 4096, 148 * 8);
            partition_hist_kernel<PK, T, HALO><<<std::max(blocks, 1), 512, 0, L->stream>>>(kin, n, op, hist, halo);
        }
        TRY(check_launch(L, "partition_hist_kernel"));
        uint32_t h[64];
        CU(L, cudaMemcpyAsync(h, hist, 64 * sizeof(uint32_t), cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaStreamSynchronize(L->stream));
        for (int b = 0; b <= n_spl; ++b) {
            counts_out[b] = h[b];
            if (halo_out) halo_out[b] = h[32 + b];
        }
        return BP_OK;
    }

    // The partition pass with one destination array per bucket (device addresses; peers' symmetric
    // memory in the multi-GPU path), plus the halo copies.  Nothing is returned: the sizes are known
    // from partition_count, which is also how the caller computed the destinations.
    template <class PK, class PV>
    static int partition_scatter(bp_layer *L, const PK *kin, const PV *vin, uint32_t n, const uint64_t *spl, int n_spl,
                                 uint32_t shift, const uint64_t *kdst, const uint64_t *vdst, const uint64_t *hkdst,
                                 const uint64_t *hvdst, const uint8_t *vflags = nullptr) {
        // (256 x 12 x 5 and 256 x 16 x 4 CTAs/SM were measured too: 0.90 / 0.92 ms against 0.87 ms for 97.8 M records into 8 local buckets)
        return exchange_launch<PK, PV, 384, 12, 3>(L, kin, vin, n, spl, n_spl, shift, kdst, vdst, hkdst, hvdst, vflags);
    }

    template <class PK, class PV, int XT, int XI, int XMINB>
    static int exchange_launch(bp_layer *L, const PK *kin, const PV *vin, uint32_t n, const uint64_t *spl, int n_spl, uint32_t shift,
                               const uint64_t *kdst, const uint64_t *vdst, const uint64_t *hkdst, const uint64_t *hvdst,
                               const uint8_t *vflags) {
        typedef ExchangeCfg<PK, PV, XT, XI> Cfg;
        if (n == 0) return BP_OK;
        SplitterScatterDigit<PK> op; // (the halo copies below still go through the splitter functor)
        for (int i = 0; i < MAX_SPLITTERS; ++i) op.spl[i] = i < n_spl ? spl[i] : ~0ull;
        op.n = (uint32_t)n_spl;
        op.shift = shift;
        for (int b = 0; b <= MAX_SPLITTERS; ++b) {
            op.kdst[b] = b <= n_spl ? kdst[b] : 0;
            op.vdst[b] = (b <= n_spl && vdst) ? vdst[b] : 0;
        }
        const uint32_t tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
        const size_t total = (size_t)(64 + (size_t)tiles * XCH_BUCKETS) * sizeof(uint32_t);
        TRY(ensure(L, L->scratch, total));
        uint32_t *counters = (uint32_t *)L->scratch.p, *status = counters + 64;
        CU(L, cudaMemsetAsync(L->scratch.p, 0, total, L->stream));
        ExchangeArgs<PK, PV> a;
        a.kin = kin;
        a.vin = vin;
        a.vflags = vflags; // dedup at the source across the exchange: the cell flags leave inside the IDs
        a.n = n;
        a.n_spl = (uint32_t)n_spl;
        a.shift = shift;
        for (int i = 0; i < MAX_SPLITTERS; ++i) a.spl[i] = op.spl[i];
        for (int b = 0; b < XCH_BUCKETS; ++b) {
            a.kdst[b] = op.kdst[b];
            a.vdst[b] = op.vdst[b];
        }
        a.status = status;
        a.tile_counter = counters;
        a.err = L->d_err;
        {
            static const bool use_tma = !(getenv("BP_EXCHANGE_TMA") && atoi(getenv("BP_EXCHANGE_TMA")) == 0);
            auto kern_tma = exchange_pass_kernel<PK, PV, XT, XI, XMINB, true>;
            auto kern_st = exchange_pass_kernel<PK, PV, XT, XI, XMINB, false>;
            auto kern = use_tma ? kern_tma : kern_st;
            CU(L, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
            const double eb = std::is_same<PV, NoVal>::value ? sizeof(PK) : sizeof(PK) + sizeof(PV);
            LaunchScope ls(L, BP_K_PARTITION, 2.0 * (double)n * eb);
            kern<<<tiles, XT, Cfg::SMEM_BYTES, L->stream>>>(a);
        }
        TRY(check_launch(L, "exchange_pass_kernel"));
        if constexpr (!std::is_same<PV, NoVal>::value) {
            if (hkdst && hvdst) {
                SplitterScatterDigit<PK> hop = op;
                for (int b = 0; b <= MAX_SPLITTERS; ++b) {
                    hop.kdst[b] = b <= n_spl ? hkdst[b] : 0;
                    hop.vdst[b] = b <= n_spl ? hvdst[b] : 0;
                }
                uint32_t *cursor = counters + 32; // zeroed above
                LaunchScope ls(L, BP_K_PARTITION, (double)n * sizeof(PK));
                const int blocks = (int)std::min<size_t>((n + 1023) / 1024, 148 * 8);
                halo_scatter_kernel<PK, PV, T><<<std::max(blocks, 1), 256, 0, L->stream>>>(kin, vin, n, hop, cursor, vflags);
                TRY(check_launch(L, "halo_scatter_kernel"));
            }
        }
        return BP_OK;
    }

    static int count_records(bp_layer *L, const void *kin, size_t n, const uint64_t *spl, int n_spl, uint64_t *counts, uint64_t *halo) {
        return partition_count<K, true>(L, (const K *)kin, (uint32_t)n, spl, n_spl, 0, counts, halo);
    }
    static int count_records_row(bp_layer *L, const void *kin, size_t n, const uint64_t *spl, int n_spl, const uint64_t *tags, int n_tags,
                                 const uint64_t *d_rows, int n_rows) {
        return partition_count<K, true>(L, (const K *)kin, (uint32_t)n, spl, n_spl, 0, nullptr, nullptr, d_rows, n_rows, tags, n_tags);
    }
    static int scatter_records(bp_layer *L, const void *kin, const void *vin, size_t n, const uint64_t *spl, int n_spl,
                               const uint64_t *kdst, const uint64_t *vdst, const uint64_t *hkdst, const uint64_t *hvdst, bool fold) {
        const uint8_t *vflags = nullptr;
        if (fold) { // only the layer's own freshly encoded records have cell flags, and only small enough IDs leave room
            const int id_bits = 64 - (L->id_or ? __builtin_clzll(L->id_or) : 64);
            if (kin != (const void *)keys(L, L->cur) || n != L->n_records || !L->flags_valid || L->ids_flagged ||
                id_bits > (int)(8 * sizeof(IdT)) - 3)
                return fail(L, BP_ERR_INVALID_ARG, "scatter_records: cell flags can only be folded into the layer's own encoded records");
            vflags = (const uint8_t *)L->cell_flags.p;
        }
        return partition_scatter<K, IdT>(L, (const K *)kin, (const IdT *)vin, (uint32_t)n, spl, n_spl, 0, kdst, vdst, hkdst, hvdst, vflags);
    }

    static int partition_records(bp_layer *L, const void *kin, const void *vin, size_t n, const uint64_t *spl, int n_spl,
                                 void *kout, void *vout, uint64_t *counts) {
        return partition<K, IdT>(L, (const K *)kin, (const IdT *)vin, (uint32_t)n, spl, n_spl, 0, (K *)kout, (IdT *)vout, counts);
    }

    static int lookup_ranges(bp_layer *L, const void *sorted_keys, size_t n, const uint64_t *queries, int nq, uint64_t *lo,
                             uint64_t *hi) {
        if (nq <= 0) return BP_OK;
        const size_t bytes = (size_t)nq * (sizeof(K) + 2 * sizeof(uint32_t));
        TRY(ensure(L, L->scratch, bytes));
        std::vector<K> hq(nq);
        for (int i = 0; i < nq; ++i) hq[i] = (K)queries[i];
        K *dq = (K *)L->scratch.p;
        uint32_t *dlo = (uint32_t *)(dq + nq), *dhi = dlo + nq;
        CU(L, cudaMemcpyAsync(dq, hq.data(), nq * sizeof(K), cudaMemcpyHostToDevice, L->stream));
        {
            LaunchScope ls(L, BP_K_MISC, 0);
            lookup_ranges_kernel<K><<<(nq + 127) / 128, 128, 0, L->stream>>>((const K *)sorted_keys, (uint32_t)n, dq, (uint32_t)nq, dlo, dhi);
        }
        TRY(check_launch(L, "lookup_ranges_kernel"));
        std::vector<uint32_t> h(2 * nq);
        CU(L, cudaMemcpyAsync(h.data(), dlo, 2 * nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaStreamSynchronize(L->stream));
        for (int i = 0; i < nq; ++i) {
            lo[i] = h[i];
            hi[i] = h[nq + i];
        }
        return BP_OK;
    }

    static int masks_from_records(bp_layer *L, uint64_t n) {
        ExtendResult init;
        memset(&init, 0, sizeof init);
        init.key_and = ~0ull;
        init.id_and = ~0ull;
        *L->h_res = init;
        CU(L, cudaMemcpyAsync(L->d_res, L->h_res, sizeof init, cudaMemcpyHostToDevice, L->stream));
        if (n) {
            LaunchScope ls(L, BP_K_MISC, 0);
            const int blocks = (int)std::min<uint64_t>((n + 255) / 256, 148 * 8);
            const IdT id_mask = L->ids_flagged ? (IdT)((((IdT)1) << (8 * sizeof(IdT) - 3)) - 1) : (IdT) ~(IdT)0;
            record_masks_kernel<K, IdT><<<blocks, 256, 0, L->stream>>>(keys(L, L->cur), ids(L, L->cur), (uint32_t)n, id_mask, L->d_res);
        }
        TRY(check_launch(L, "record_masks_kernel"));
        CU(L, cudaMemcpyAsync(L->h_res, L->d_res, sizeof(ExtendResult), cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaStreamSynchronize(L->stream));
        return BP_OK;
    }
};

// ---- runtime dispatch over (kind, id_bytes) ----------------------------------------------------------------
#define DISPATCH(L, CALL)                                                              \
    do {                                                                               \
        if ((L)->id_bytes == 4) {                                                      \
            switch ((L)->kind) {                                                       \
            case BP_INDEX32_2D: return Impl<BP_INDEX32_2D, uint32_t>::CALL;            \
            case BP_INDEX64_2D: return Impl<BP_INDEX64_2D, uint32_t>::CALL;            \
            default: return Impl<BP_INDEX64_3D, uint32_t>::CALL;                       \
            }                                                                          \
        } else {                                                                       \
            switch ((L)->kind) {                                                       \
            case BP_INDEX32_2D: return Impl<BP_INDEX32_2D, uint64_t>::CALL;            \
            case BP_INDEX64_2D: return Impl<BP_INDEX64_2D, uint64_t>::CALL;            \
            default: return Impl<BP_INDEX64_3D, uint64_t>::CALL;                       \
            }                                                                          \
        }                                                                              \
    } while (0)

int do_ensure_tree(bp_layer *L, uint64_t records) { DISPATCH(L, ensure_tree(L, records)); }
int do_extend_device(bp_layer *L, const float *sysb, const float *b, const void *ids, size_t n) {
    DISPATCH(L, extend_device(L, sysb, b, ids, n));
}
int do_sort(bp_layer *L) { DISPATCH(L, sort(L)); }
int do_scan(bp_layer *L, const bp_filter *f) { DISPATCH(L, scan(L, f)); }
int do_scan_raw(bp_layer *L, const bp_filter *f, uint64_t *out_raw) { DISPATCH(L, scan_raw(L, f, out_raw)); }
int do_finish_pairs(bp_layer *L, uint64_t n) { DISPATCH(L, finish_pairs(L, n)); }
int do_pick(bp_layer *L, const float *sysb, const float *d_rays, size_t nq, float max_dist, int max_depth, int shape_kind,
            const float *d_shapes, size_t n_shapes) {
    DISPATCH(L, pick(L, sysb, d_rays, nq, max_dist, max_depth, shape_kind, d_shapes, n_shapes));
}
int do_query(bp_layer *L, int ray, const float *sysb, const float *d_params, size_t nq, int max_depth) {
    DISPATCH(L, query_kind(L, ray, sysb, d_params, nq, max_depth));
}
int do_partition_records(bp_layer *L, const void *kin, const void *vin, size_t n, const uint64_t *spl, int n_spl, void *kout,
                         void *vout, uint64_t *counts) {
    DISPATCH(L, partition_records(L, kin, vin, n, spl, n_spl, kout, vout, counts));
}
int do_partition_pairs(bp_layer *L, const uint64_t *kin, size_t n, const uint64_t *spl, int n_spl, uint64_t *kout, uint64_t *counts) {
    return Impl<BP_INDEX64_3D, uint32_t>::partition<uint64_t, NoVal>(L, kin, (const NoVal *)nullptr, (uint32_t)n, spl, n_spl, 32, kout,
                                                                     (NoVal *)nullptr, counts);
}
int do_count_records(bp_layer *L, const void *kin, size_t n, const uint64_t *spl, int n_spl, uint64_t *counts, uint64_t *halo) {
    DISPATCH(L, count_records(L, kin, n, spl, n_spl, counts, halo));
}
int do_scatter_records(bp_layer *L, const void *kin, const void *vin, size_t n, const uint64_t *spl, int n_spl, const uint64_t *kdst,
                       const uint64_t *vdst, const uint64_t *hkdst, const uint64_t *hvdst, bool fold) {
    DISPATCH(L, scatter_records(L, kin, vin, n, spl, n_spl, kdst, vdst, hkdst, hvdst, fold));
}
int do_count_records_row(bp_layer *L, const void *kin, size_t n, const uint64_t *spl, int n_spl, const uint64_t *tags, int n_tags,
                         const uint64_t *d_rows, int n_rows) {
    DISPATCH(L, count_records_row(L, kin, n, spl, n_spl, tags, n_tags, d_rows, n_rows));
}
int do_count_pairs_row(bp_layer *L, const uint64_t *kin, size_t n, const uint64_t *spl, int n_spl, const uint64_t *tags, int n_tags,
                       const uint64_t *d_rows, int n_rows) {
    return Impl<BP_INDEX64_3D, uint32_t>::partition_count<uint64_t, false>(L, kin, (uint32_t)n, spl, n_spl, 32, nullptr, nullptr, d_rows,
                                                                           n_rows, tags, n_tags);
}
int do_count_pairs(bp_layer *L, const uint64_t *kin, size_t n, const uint64_t *spl, int n_spl, uint64_t *counts) {
    return Impl<BP_INDEX64_3D, uint32_t>::partition_count<uint64_t, false>(L, kin, (uint32_t)n, spl, n_spl, 32, counts, nullptr);
}
int do_scatter_pairs(bp_layer *L, const uint64_t *kin, size_t n, const uint64_t *spl, int n_spl, const uint64_t *kdst) {
    return Impl<BP_INDEX64_3D, uint32_t>::partition_scatter<uint64_t, NoVal>(L, kin, (const NoVal *)nullptr, (uint32_t)n, spl, n_spl, 32,
                                                                             kdst, nullptr, nullptr, nullptr);
}
int do_lookup_ranges(bp_layer *L, const void *keys, size_t n, const uint64_t *q, int nq, uint64_t *lo, uint64_t *hi) {
    DISPATCH(L, lookup_ranges(L, keys, n, q, nq, lo, hi));
}
int do_extend_count_rows(bp_layer *L, const float *sysb, const float *d_bounds, const void *d_ids, size_t n, const uint64_t *spl,
                         int n_spl, bool allow_fold, const uint64_t *d_rows, int n_rows) {
    DISPATCH(L, extend_count_rows(L, sysb, d_bounds, d_ids, n, spl, n_spl, allow_fold, d_rows, n_rows));
}
int do_sort_from(bp_layer *L, const void *k, const void *v, uint64_t n, bool asc) { DISPATCH(L, sort_from(L, k, v, n, asc)); }
int do_masks(bp_layer *L, uint64_t n) { DISPATCH(L, masks_from_records(L, n)); }
int do_strip_flags(bp_layer *L) { DISPATCH(L, strip_flags(L)); }

void detach_lazy(bp_layer *L) {
    if (!L->lazy_src) return;
    auto &rd = L->lazy_src->lazy_readers;
    rd.erase(std::remove(rd.begin(), rd.end(), L), rd.end());
    L->lazy_src = nullptr;
    L->lazy_n = 0;
}

// Performs a deferred Layer::merge as the plain append it stands for: the other layer's records are copied behind ours.
int materialize_lazy(bp_layer *L) {
    bp_layer *O = L->lazy_src;
    if (!O) return BP_OK;
    const uint64_t add = L->lazy_n, base = L->n_records - add;
    DeviceGuard g(L->device);
    if (O->stream != L->stream) {
        CU(L, cudaEventRecord(O->ev_sync, O->stream));
        CU(L, cudaStreamWaitEvent(L->stream, O->ev_sync, 0));
    }
    CU(L, cudaMemcpyAsync((char *)L->keys[L->cur].p + base * L->key_bytes, O->keys[O->cur].p, add * L->key_bytes,
                          cudaMemcpyDeviceToDevice, L->stream));
    CU(L, cudaMemcpyAsync((char *)L->ids[L->cur].p + base * L->id_bytes, O->ids[O->cur].p, add * L->id_bytes,
                          cudaMemcpyDeviceToDevice, L->stream));
    if (O->stream != L->stream) { // and keep the other layer from overwriting its tree before the copy ran
        CU(L, cudaEventRecord(L->ev_sync, L->stream));
        CU(L, cudaStreamWaitEvent(O->stream, L->ev_sync, 0));
    }
    detach_lazy(L);
    return BP_OK;
}

// Folds the result of the last (still asynchronous) extend into the host-side state.  Every entry point calls this
// first; unless keep_lazy (the sort / scan entry points, which merge a deferred Layer::merge out of both trees), a
// deferred merge INTO this layer is materialised, and so is every deferred merge FROM this layer (the call may change it).
int resolve_pending(bp_layer *L, bool keep_lazy) {
    if (!keep_lazy && L->lazy_src) TRY(materialize_lazy(L));
    while (!L->lazy_readers.empty()) TRY(materialize_lazy(L->lazy_readers.back()));
    if (!L->pending) return BP_OK;
    TRY(wait_mail(L, 0, L->pending_seq, L->h_res, sizeof(ExtendResult)));
    L->pending = false;
    const ExtendResult &r = *L->h_res;
    // An object that wants more cells than the encoder enumerates (only a min_depth far above its natural depth does that;
    // the reference warn!s and heap-allocates, src/geom.rs:299-301): the extend call is rejected as a whole -- nothing it
    // wrote lies inside the tree, whose length, flags and masks stay what they were before the call.
    if (r.too_many)
        return fail(L, BP_ERR_TOO_LARGE, "an object of the last extend wanted more than 2^20 cells (min_depth too high for its size): "
                                          "the call was rejected, the tree is unchanged");
    L->n_invalid += r.n_invalid;
    L->stats.n_invalid = L->n_invalid;
    L->tail_has_last = true;
    L->last_slot ^= 1;
    const uint64_t room = L->cap_records - L->pending_base;
    const uint64_t added = std::min<uint64_t>(r.total_records, room);
    if (added > 0) {
        L->stats.algo_bytes[BP_K_ENCODE] += (double)added * (L->key_bytes + L->id_bytes);
        if (!L->dirty) { // first append onto a sorted tree: everything before it stays a sorted prefix
            L->dirty = true;
            L->prefix = L->pending_base;
            L->tail_nonmono = false;
        }
        L->tail_sorted = false;
        if (r.nonmono) L->tail_nonmono = true;
        if (L->pending_base == 0) L->id_first = r.id_first;
        L->id_last = r.id_last;
        L->n_records = L->pending_base + added;
        L->key_or |= r.key_or;
        L->key_and &= r.key_and;
        L->id_or |= r.id_or;
        L->id_and &= r.id_and;
    }
    L->stats.n_records = L->n_records;
    if (r.total_records > room) return fail(L, BP_ERR_INTERNAL, "record capacity exceeded");
    return BP_OK;
}

int pinned_ensure(bp_layer *L, void **p, size_t *cap, size_t bytes) {
    if (bytes <= *cap) return BP_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    size_t want = std::max(bytes, (size_t)4096);
    want += want / 4;
    cudaError_t e = cudaMallocHost(p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(L, BP_ERR_OOM, "cudaMallocHost of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    *cap = want;
    return BP_OK;
}

} // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int bp_version(void) { return BP_VERSION; }

const char *bp_status_string(int s) {
    switch (s) {
    case BP_OK: return "ok";
    case BP_ERR_INVALID_ARG: return "invalid argument";
    case BP_ERR_CUDA: return "CUDA error";
    case BP_ERR_OOM: return "out of memory";
    case BP_ERR_TOO_LARGE: return "too large";
    case BP_ERR_INTERNAL: return "internal error";
    case BP_ERR_MISMATCH: return "layer type mismatch";
    default: return "unknown status";
    }
}

int bp_device_count(int *out) {
    if (!out) return BP_ERR_INVALID_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = 0;
        return BP_ERR_CUDA;
    }
    *out = n;
    return BP_OK;
}

int bp_plan_sort_finish(uint64_t varying_mask, uint64_t n_records, uint64_t *out_top_mask, uint32_t *out_group_shift) {
    uint64_t top = 0;
    uint32_t gshift = 0;
    const bool yes = plan_top_bits(varying_mask, n_records, &top, &gshift);
    if (out_top_mask) *out_top_mask = yes ? top : varying_mask;
    if (out_group_shift) *out_group_shift = yes ? gshift : 0u;
    return yes ? 1 : 0;
}

int bp_plan_radix_passes(uint64_t varying_mask, uint32_t *out_shift, uint32_t *out_bits, int max_passes) {
    RadixPlan plan;
    memset(&plan, 0, sizeof plan);
    const int np = plan_passes(varying_mask, plan);
    for (int i = 0; i < np && i < max_passes; ++i) { // second field in the high half-words
        if (out_shift) out_shift[i] = plan.shift[i] | ((uint32_t)plan.shift2[i] << 16);
        if (out_bits) out_bits[i] = plan.bits[i] | ((uint32_t)plan.bits2[i] << 16);
    }
    return np;
}

const char *bp_layer_last_error(const bp_layer *L) { return L ? L->last_error.c_str() : "null layer"; }

int bp_layer_create(const bp_layer_config *cfg, bp_layer **out) {
    if (!cfg || !out) return BP_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->index_kind < BP_INDEX32_2D || cfg->index_kind > BP_INDEX64_3D) return BP_ERR_INVALID_ARG;
    if (cfg->id_bytes != 4 && cfg->id_bytes != 8) return BP_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return BP_ERR_CUDA; // no CPU fallback by design
    }
    int dev = cfg->device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) return BP_ERR_CUDA;
    }
    if (dev >= ndev) return BP_ERR_INVALID_ARG;
    bp_layer *L = new bp_layer();
    L->cfg = *cfg;
    L->kind = cfg->index_kind;
    L->id_bytes = cfg->id_bytes;
    L->key_bytes = cfg->index_kind == BP_INDEX32_2D ? 4 : 8;
    L->dim = cfg->index_kind == BP_INDEX64_3D ? 3 : 2;
    L->device = dev;
    L->min_depth = cfg->min_depth;
    memset(&L->stats, 0, sizeof L->stats);
    if (const char *e = getenv("BP_RADIX_BITS")) L->radix_bits_cap = atoi(e); // tuning aid: 8 = the 8-bit pass everywhere
    if (const char *e = getenv("BP_SORT_FINISH_MIN")) L->sort_finish_min = strtoull(e, nullptr, 10);
    DeviceGuard g(dev);
    auto bail = [&](int st) {
        bp_layer_destroy(L);
        return st;
    };
    if (cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(BP_ERR_CUDA);
    L->own_stream = true;
    if (cudaEventCreateWithFlags(&L->ev_sync, cudaEventDisableTiming) != cudaSuccess) return bail(BP_ERR_CUDA);
    if (cudaMallocHost((void **)&L->h_res, sizeof(ExtendResult)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMallocHost((void **)&L->h_tot, sizeof(ScanTotals)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMallocHost((void **)&L->h_err, sizeof(int)) != cudaSuccess) return bail(BP_ERR_OOM);
    for (int i = 0; i < 2; ++i) {
        if (cudaHostAlloc((void **)&L->h_mail[i], sizeof(Mailbox), cudaHostAllocMapped) != cudaSuccess) return bail(BP_ERR_OOM);
        memset(L->h_mail[i], 0, sizeof(Mailbox));
        if (cudaHostGetDevicePointer((void **)&L->d_mail[i], L->h_mail[i], 0) != cudaSuccess) return bail(BP_ERR_CUDA);
    }
    if (cudaMalloc((void **)&L->d_res, sizeof(ExtendResult)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMalloc((void **)&L->d_cnt, 32 * sizeof(uint32_t)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMalloc((void **)&L->d_tot, sizeof(ScanTotals) + 64) != cudaSuccess) return bail(BP_ERR_OOM); // + scan_runs_kernel's counters
    if (cudaMalloc((void **)&L->d_err, sizeof(int)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMalloc((void **)&L->d_last, 16) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMemsetAsync(L->d_err, 0, sizeof(int), L->stream) != cudaSuccess) return bail(BP_ERR_CUDA);
    {
        std::vector<uint32_t> lut((size_t)1 << ENCODE_LUT_BITS);
        for (uint32_t v = 0; v < lut.size(); ++v) lut[v] = (uint32_t)(L->dim == 2 ? spread2(v) : spread3(v));
        if (ensure(L, L->spread_lut, lut.size() * sizeof(uint32_t)) != BP_OK) return bail(BP_ERR_OOM);
        if (cudaMemcpyAsync(L->spread_lut.p, lut.data(), lut.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, L->stream) != cudaSuccess)
            return bail(BP_ERR_CUDA); // (pageable source: the copy has been staged when the call returns)
    }
    if (cudaStreamSynchronize(L->stream) != cudaSuccess) return bail(BP_ERR_CUDA);
    *L->h_err = 0;
    if (cfg->index_capacity) {
        int st = do_ensure_tree(L, std::min<uint64_t>(cfg->index_capacity, MAX_RECORDS));
        if (st != BP_OK) return bail(st);
    }
    if (cfg->collision_capacity) {
        const size_t c = std::min<uint64_t>(cfg->collision_capacity, MAX_RECORDS);
        int st = ensure(L, L->pout, c * 2 * L->id_bytes);
        for (int i = 0; i < 2 && st == BP_OK; ++i) st = ensure(L, L->praw[i], c * 8);
        if (st != BP_OK) return bail(st);
    }
    *out = L;
    return BP_OK;
}

int bp_layer_destroy(bp_layer *L) {
    if (!L) return BP_OK;
    while (!L->lazy_readers.empty()) materialize_lazy(L->lazy_readers.back()); // deferred merges from this layer: copy now
    detach_lazy(L);
    DeviceGuard g(L->device);
    if (L->stream) cudaStreamSynchronize(L->stream);
    for (int i = 0; i < 2; ++i) {
        release(L->keys[i]);
        release(L->ids[i]);
        release(L->praw[i]);
        release(L->praw_b[i]);
    }
    release(L->scratch);
    release(L->cell_flags);
    release(L->spread_lut);
    release(L->stage_bounds);
    release(L->stage_ids);
    release(L->src_idx);
    release(L->src_off);
    release(L->chunk_src);
    release(L->inactive);
    release(L->pout);
    release(L->pair_cnt);
    release(L->query_params);
    release(L->pick_shapes);
    release(L->pick_out);
    release(L->query_counts);
    release(L->query_offsets);
    release(L->filter_table);
    if (L->h_pairs) cudaFreeHost(L->h_pairs);
    if (L->h_offsets) cudaFreeHost(L->h_offsets);
    if (L->h_pick) cudaFreeHost(L->h_pick);
    if (L->h_keys) cudaFreeHost(L->h_keys);
    if (L->h_ids) cudaFreeHost(L->h_ids);
    if (L->h_res) cudaFreeHost(L->h_res);
    if (L->h_tot) cudaFreeHost(L->h_tot);
    if (L->h_err) cudaFreeHost(L->h_err);
    for (int i = 0; i < 2; ++i)
        if (L->h_mail[i]) cudaFreeHost(L->h_mail[i]);
    if (L->d_res) cudaFree(L->d_res);
    if (L->d_cnt) cudaFree(L->d_cnt);
    if (L->d_tot) cudaFree(L->d_tot);
    if (L->d_err) cudaFree(L->d_err);
    if (L->d_last) cudaFree(L->d_last);
    for (ProfEvent &pe : L->prof_events) {
        cudaEventDestroy(pe.start);
        cudaEventDestroy(pe.stop);
    }
    for (ProfEvent &pe : L->prof_pool) {
        cudaEventDestroy(pe.start);
        cudaEventDestroy(pe.stop);
    }
    if (L->ev_sync) cudaEventDestroy(L->ev_sync);
    for (cudaEvent_t e : L->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (L->copy_stream) cudaStreamDestroy(L->copy_stream);
    if (L->own_stream && L->stream) cudaStreamDestroy(L->stream);
    cudaGetLastError();
    delete L;
    return BP_OK;
}

int bp_layer_set_stream(bp_layer *L, void *stream) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    CU(L, cudaStreamSynchronize(L->stream));
    if (L->own_stream) cudaStreamDestroy(L->stream);
    L->stream = (cudaStream_t)stream;
    L->own_stream = false;
    return BP_OK;
}

int bp_layer_clear(bp_layer *L) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    if (L->pending) { // the result of an extend nobody looked at is discarded with the tree
        TRY(wait_mail(L, 0, L->pending_seq, L->h_res, sizeof(ExtendResult)));
        L->pending = false;
        L->n_invalid += L->h_res->n_invalid;
    }
    L->n_records = 0;
    L->n_halo = 0;
    L->flags_valid = true;
    L->ids_flagged = false;
    L->dirty = false;
    L->prefix = 0;
    L->tail_sorted = false;
    L->tail_nonmono = false;
    L->tail_has_last = false;
    L->key_or = L->id_or = 0;
    L->key_and = L->id_and = ~0ull;
    L->stats.n_records = 0;
    return BP_OK;
}

int bp_layer_extend_device(bp_layer *L, const float *sysb, const float *d_bounds, const void *d_ids, size_t n) {
    if (!L || !sysb || (n && (!d_bounds || !d_ids))) return fail(L, BP_ERR_INVALID_ARG, "null argument to extend");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_extend_device(L, sysb, d_bounds, d_ids, n);
}

int bp_layer_extend_host(bp_layer *L, const float *sysb, const float *bounds, const void *ids, size_t n) {
    if (!L || !sysb || (n && (!bounds || !ids))) return fail(L, BP_ERR_INVALID_ARG, "null argument to extend");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    if (n == 0) return BP_OK;
    const size_t bbytes = n * 2 * L->dim * sizeof(float), ibytes = n * L->id_bytes;
    // Pinned (page-locked, device-mapped) host buffers are read by the encode kernel where they are: every AABB and ID is
    // read exactly once, so the PCIe transfer IS the kernel's input stream and the encode hides under it, instead of a
    // staging copy followed by the encode.  Pageable memory still goes through the staging buffers.
    const void *d_b = nullptr, *d_i = nullptr;
    {
        static const bool zero_copy = !(getenv("BP_EXTEND_ZERO_COPY") && atoi(getenv("BP_EXTEND_ZERO_COPY")) == 0);
        cudaPointerAttributes ab, ai;
        if (zero_copy && cudaPointerGetAttributes(&ab, bounds) == cudaSuccess && cudaPointerGetAttributes(&ai, ids) == cudaSuccess &&
            ab.type == cudaMemoryTypeHost && ai.type == cudaMemoryTypeHost && ab.devicePointer && ai.devicePointer &&
            ((uintptr_t)ab.devicePointer & 15u) == 0 && ((uintptr_t)ai.devicePointer & 15u) == 0) {
            d_b = ab.devicePointer;
            d_i = ai.devicePointer;
        }
        cudaGetLastError(); // (a failed attribute query on a plain malloc pointer leaves an error behind on old drivers)
    }
    const size_t obj_bytes = 2 * L->dim * sizeof(float) + L->id_bytes;
    constexpr size_t ZERO_COPY_MAX = 64u << 20; // beyond this the copy engine's DMA beats the SMs' PCIe reads (measured: 2^25 objects)
    if (d_b && n * obj_bytes <= ZERO_COPY_MAX) {
        TRY(do_extend_device(L, sysb, (const float *)d_b, d_i, n));
    } else {
        TRY(ensure(L, L->stage_bounds, bbytes));
        TRY(ensure(L, L->stage_ids, ibytes));
        // Chunked: every chunk's copy is queued on the copy stream up front, the encode of chunk c starts when its copy
        // has landed -- extend + extend + ... of consecutive object ranges, i.e. the same records in the same order, with
        // the encode hidden under the next chunk's transfer.
        const size_t min_chunk = (size_t)1 << 20; // objects
        int chunks = (int)std::min<size_t>(16, std::max<size_t>(1, n / min_chunk));
        if (!d_b) chunks = 1;                     // pageable memory: cudaMemcpyAsync is synchronous with the host anyway
        if (chunks > 1 && !L->copy_stream) {
            if (cudaStreamCreateWithFlags(&L->copy_stream, cudaStreamNonBlocking) != cudaSuccess) chunks = 1;
            for (int c = 0; c < 16 && chunks > 1; ++c)
                if (cudaEventCreateWithFlags(&L->ev_chunk[c], cudaEventDisableTiming) != cudaSuccess) chunks = 1;
        }
        if (chunks == 1) {
            CU(L, cudaMemcpyAsync(L->stage_bounds.p, bounds, bbytes, cudaMemcpyHostToDevice, L->stream));
            CU(L, cudaMemcpyAsync(L->stage_ids.p, ids, ibytes, cudaMemcpyHostToDevice, L->stream));
            TRY(do_extend_device(L, sysb, (const float *)L->stage_bounds.p, L->stage_ids.p, n));
        } else {
            const size_t per = ((n + chunks - 1) / chunks + 3) & ~(size_t)3; // 16-byte aligned ID chunks
            const size_t bstride = 2 * L->dim * sizeof(float);
            CU(L, cudaEventRecord(L->ev_sync, L->stream)); // the staging buffers may still be read by an earlier extend
            CU(L, cudaStreamWaitEvent(L->copy_stream, L->ev_sync, 0));
            for (int c = 0; c < chunks; ++c) {
                const size_t o = std::min(n, per * c), m = std::min(n, per * (c + 1)) - o;
                if (m) {
                    CU(L, cudaMemcpyAsync((char *)L->stage_bounds.p + o * bstride, (const char *)bounds + o * bstride, m * bstride,
                                          cudaMemcpyHostToDevice, L->copy_stream));
                    CU(L, cudaMemcpyAsync((char *)L->stage_ids.p + o * L->id_bytes, (const char *)ids + o * L->id_bytes, m * L->id_bytes,
                                          cudaMemcpyHostToDevice, L->copy_stream));
                }
                CU(L, cudaEventRecord(L->ev_chunk[c], L->copy_stream));
            }
            for (int c = 0; c < chunks; ++c) {
                const size_t o = std::min(n, per * c), m = std::min(n, per * (c + 1)) - o;
                CU(L, cudaStreamWaitEvent(L->stream, L->ev_chunk[c], 0));
                if (!m) continue;
                TRY(do_extend_device(L, sysb, (const float *)((char *)L->stage_bounds.p + o * bstride),
                                      (char *)L->stage_ids.p + o * L->id_bytes, m));
                if (c + 1 < chunks) TRY(resolve_pending(L)); // the next chunk appends behind this one's records
            }
        }
    }
    // the caller may reuse its buffers as soon as we return
    return resolve_pending(L);
}

int bp_layer_merge(bp_layer *L, const bp_layer *O_) {
    bp_layer *O = const_cast<bp_layer *>(O_);
    if (!L || !O) return BP_ERR_INVALID_ARG;
    if (L == O) return fail(L, BP_ERR_INVALID_ARG, "cannot merge a layer into itself");
    if (L->kind != O->kind || L->id_bytes != O->id_bytes || L->device != O->device)
        return fail(L, BP_ERR_MISMATCH, "layers differ in index kind, ID width or device");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    TRY(resolve_pending(O));
    // Two sorted layers whose cell flags ride in their IDs (dedup at the source) keep them through the merge: the merged
    // tree is flagged too, and the scan emits every ID pair from its canonical cell only.  Anything else: strip, as before.
    const uint64_t id_all = L->id_or | O->id_or;
    const int id_bits = 64 - (id_all ? __builtin_clzll(id_all) : 64);
    const bool keep_flags = !L->dirty && !O->dirty && L->n_records && O->n_records && L->ids_flagged && O->ids_flagged &&
                            id_bits <= 8 * L->id_bytes - 3;
    if (!keep_flags) {
        TRY(do_strip_flags(L));
        TRY(do_strip_flags(O));
    }
    L->flags_valid = false; // the appended records carry no separate cell flags
    if (O->min_depth < L->min_depth) L->min_depth = O->min_depth; // src/layer.rs:131-134
    const uint64_t base = L->n_records, add = O->n_records;
    if (add) {
        TRY(do_ensure_tree(L, base + add));
        if (!L->dirty && !O->dirty && base > 0 && O->lazy_src == nullptr) {
            // sorted into sorted: defer the append, the next sort merges straight out of both trees
            L->lazy_src = O;
            L->lazy_n = add;
            O->lazy_readers.push_back(L);
        } else {
            // order the copy after everything queued on the other layer's stream
            if (O->stream != L->stream) {
                CU(L, cudaEventRecord(O->ev_sync, O->stream));
                CU(L, cudaStreamWaitEvent(L->stream, O->ev_sync, 0));
            }
            CU(L, cudaMemcpyAsync((char *)L->keys[L->cur].p + base * L->key_bytes, O->keys[O->cur].p, add * L->key_bytes,
                                  cudaMemcpyDeviceToDevice, L->stream));
            CU(L, cudaMemcpyAsync((char *)L->ids[L->cur].p + base * L->id_bytes, O->ids[O->cur].p, add * L->id_bytes,
                                  cudaMemcpyDeviceToDevice, L->stream));
            if (O->stream != L->stream) { // and keep the other layer from overwriting its tree before the copy ran
                CU(L, cudaEventRecord(L->ev_sync, L->stream));
                CU(L, cudaStreamWaitEvent(O->stream, L->ev_sync, 0));
            }
        }
    }
    if (!L->dirty) {
        L->prefix = base;
        L->tail_sorted = !O->dirty; // a sorted tree appended to a sorted tree: two runs -> merge path
        L->tail_nonmono = true;
    } else {
        L->tail_sorted = false;
        L->tail_nonmono = true;
    }
    L->tail_has_last = false;
    L->dirty = true; // `*sorted = false` even when `other` is empty -- src/layer.rs:137
    L->n_records = base + add;
    L->key_or |= O->key_or;
    L->key_and &= O->key_and;
    L->id_or |= O->id_or;
    L->id_and &= O->id_and;
    L->stats.n_records = L->n_records;
    return BP_OK;
}

int bp_layer_sort(bp_layer *L) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L, true));
    return do_sort(L);
}

int bp_layer_scan_device(bp_layer *L, const bp_filter *f, const void **out_pairs, size_t *out_count) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L, true));
    TRY(do_scan(L, f));
    if (out_pairs) *out_pairs = L->n_pairs ? L->pout.p : nullptr;
    if (out_count) *out_count = (size_t)L->n_pairs;
    return BP_OK;
}

int bp_layer_scan(bp_layer *L, const bp_filter *f, const void **out_pairs, size_t *out_count) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L, true));
    TRY(do_scan(L, f));
    const size_t bytes = (size_t)L->n_pairs * 2 * L->id_bytes;
    if (bytes) {
        TRY(pinned_ensure(L, &L->h_pairs, &L->h_pairs_cap, bytes));
        CU(L, cudaMemcpyAsync(L->h_pairs, L->pout.p, bytes, cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaStreamSynchronize(L->stream));
    }
    if (out_pairs) *out_pairs = bytes ? L->h_pairs : nullptr;
    if (out_count) *out_count = (size_t)L->n_pairs;
    return BP_OK;
}

static int query_batch(bp_layer *L, int ray, const float *sysb, const float *params, size_t nq, int32_t max_depth, int on_device,
                       const void **out_pairs, const uint32_t **out_offsets, size_t *out_count) {
    if (!L || !sysb || (nq && !params)) return fail(L, BP_ERR_INVALID_ARG, "bad arguments to a batched query");
    if (nq >= (1ull << 30)) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 queries in one batch");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    const size_t per = (size_t)(ray ? 2 * L->dim + 2 : 2 * L->dim) * sizeof(float);
    const float *d_params = params;
    if (!on_device && nq) {
        TRY(ensure(L, L->query_params, nq * per));
        CU(L, cudaMemcpyAsync(L->query_params.p, params, nq * per, cudaMemcpyHostToDevice, L->stream));
        d_params = (const float *)L->query_params.p;
    }
    TRY(do_query(L, ray, sysb, d_params, nq, max_depth));
    const void *pairs = L->n_pairs ? L->pout.p : nullptr;
    const uint32_t *offsets = (const uint32_t *)L->query_offsets.p;
    if (!on_device) {
        const size_t pbytes = (size_t)L->n_pairs * 2 * L->id_bytes, obytes = (nq + 1) * sizeof(uint32_t);
        TRY(pinned_ensure(L, &L->h_offsets, &L->h_offsets_cap, obytes));
        CU(L, cudaMemcpyAsync(L->h_offsets, L->query_offsets.p, obytes, cudaMemcpyDeviceToHost, L->stream));
        if (pbytes) {
            TRY(pinned_ensure(L, &L->h_pairs, &L->h_pairs_cap, pbytes));
            CU(L, cudaMemcpyAsync(L->h_pairs, L->pout.p, pbytes, cudaMemcpyDeviceToHost, L->stream));
        }
        CU(L, cudaStreamSynchronize(L->stream));
        pairs = pbytes ? L->h_pairs : nullptr;
        offsets = (const uint32_t *)L->h_offsets;
    }
    if (out_pairs) *out_pairs = pairs;
    if (out_offsets) *out_offsets = offsets;
    if (out_count) *out_count = (size_t)L->n_pairs;
    return BP_OK;
}

int bp_layer_test_box_batch(bp_layer *L, const float *sysb, const float *boxes, size_t nq, int32_t max_depth, int on_device,
                            const void **out_pairs, const uint32_t **out_offsets, size_t *out_count) {
    return query_batch(L, 0, sysb, boxes, nq, max_depth, on_device, out_pairs, out_offsets, out_count);
}

int bp_layer_test_ray_batch(bp_layer *L, const float *sysb, const float *rays, size_t nq, int32_t max_depth, int on_device,
                            const void **out_pairs, const uint32_t **out_offsets, size_t *out_count) {
    return query_batch(L, 1, sysb, rays, nq, max_depth, on_device, out_pairs, out_offsets, out_count);
}

int bp_layer_pick_ray_batch(bp_layer *L, const float *sysb, const float *rays, size_t nq, float max_dist, int32_t max_depth,
                            int32_t shape_kind, const float *shapes, size_t n_shapes, int on_device, const bp_pick_result **out_results) {
    static_assert(sizeof(bp_pick_result) == sizeof(PickResult) && sizeof(PickResult) == 32, "bp_pick_result layout");
    if (!L || !sysb || (nq && !rays) || (n_shapes && !shapes)) return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_layer_pick_ray_batch");
    if (nq >= (1ull << 30)) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 rays in one batch");
    if (shape_kind != BP_PICK_SPHERE && shape_kind != BP_PICK_AABB) return fail(L, BP_ERR_INVALID_ARG, "unknown pick shape kind %d", shape_kind);
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    const float *d_rays = rays, *d_shapes = shapes;
    if (!on_device) {
        const size_t rbytes = nq * 2 * L->dim * sizeof(float);
        const size_t sbytes = n_shapes * (shape_kind == BP_PICK_SPHERE ? L->dim + 1 : 2 * L->dim) * sizeof(float);
        if (rbytes) {
            TRY(ensure(L, L->query_params, rbytes));
            CU(L, cudaMemcpyAsync(L->query_params.p, rays, rbytes, cudaMemcpyHostToDevice, L->stream));
            d_rays = (const float *)L->query_params.p;
        }
        if (sbytes) {
            TRY(ensure(L, L->pick_shapes, sbytes));
            CU(L, cudaMemcpyAsync(L->pick_shapes.p, shapes, sbytes, cudaMemcpyHostToDevice, L->stream));
            d_shapes = (const float *)L->pick_shapes.p;
        }
    }
    TRY(do_pick(L, sysb, d_rays, nq, max_dist, max_depth, shape_kind, d_shapes, n_shapes));
    const bp_pick_result *res = (const bp_pick_result *)L->pick_out.p;
    if (!on_device) {
        const size_t obytes = std::max<size_t>(nq, 1) * sizeof(bp_pick_result);
        TRY(pinned_ensure(L, &L->h_pick, &L->h_pick_cap, obytes));
        if (nq) CU(L, cudaMemcpyAsync(L->h_pick, L->pick_out.p, nq * sizeof(bp_pick_result), cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaStreamSynchronize(L->stream));
        res = (const bp_pick_result *)L->h_pick;
    }
    if (out_results) *out_results = res;
    return BP_OK;
}

int bp_layer_set_halo(bp_layer *L, size_t n_halo) {
    if (!L) return BP_ERR_INVALID_ARG;
    L->n_halo = n_halo;
    return BP_OK;
}

int bp_layer_set_scan_dedup(bp_layer *L, int enabled) {
    if (!L) return BP_ERR_INVALID_ARG;
    L->scan_dedup = enabled != 0;
    return BP_OK;
}

int bp_layer_set_pair_later_fixed(bp_layer *L, uint64_t fixed_bits) {
    if (!L) return BP_ERR_INVALID_ARG;
    L->pair_later_fixed = fixed_bits;
    return BP_OK;
}

int bp_layer_scan_raw_device(bp_layer *L, const bp_filter *f, const void **out_raw, size_t *out_count) {
    if (!L) return BP_ERR_INVALID_ARG;
    if (L->id_bytes != 4) return fail(L, BP_ERR_INVALID_ARG, "raw pairs are exposed for 32-bit IDs only");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L, true));
    uint64_t n = 0;
    TRY(do_scan_raw(L, f, &n));
    if (out_raw) *out_raw = n ? L->praw[0].p : nullptr;
    if (out_count) *out_count = (size_t)n;
    return BP_OK;
}

static int unique_pairs_impl(bp_layer *L, const void *d_raw, size_t n, uint64_t id_mask, const void **out_pairs, size_t *out_count,
                             bool in_place);

int bp_layer_unique_pairs_device(bp_layer *L, const void *d_raw, size_t n, uint64_t id_mask, const void **out_pairs, size_t *out_count) {
    return unique_pairs_impl(L, d_raw, n, id_mask, out_pairs, out_count, false);
}

int bp_layer_unique_pairs_inplace_device(bp_layer *L, void *d_raw, size_t n, uint64_t id_mask, const void **out_pairs, size_t *out_count) {
    return unique_pairs_impl(L, d_raw, n, id_mask, out_pairs, out_count, true);
}

static int unique_pairs_impl(bp_layer *L, const void *d_raw, size_t n, uint64_t id_mask, const void **out_pairs, size_t *out_count,
                             bool in_place) {
    if (!L || (n && !d_raw)) return BP_ERR_INVALID_ARG;
    if (L->id_bytes != 4) return fail(L, BP_ERR_INVALID_ARG, "raw pairs are exposed for 32-bit IDs only");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 pairs");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    if (n && in_place && d_raw != L->praw[0].p) {
        L->pairs_src = (void *)d_raw; // sorted where they are: the caller's array is one side of the ping-pong
    } else if (n) {
        TRY(ensure(L, L->praw[0], n * sizeof(uint64_t)));
        if (d_raw != L->praw[0].p)
            CU(L, cudaMemcpyAsync(L->praw[0].p, d_raw, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, L->stream));
    }
    // the digit plan comes from the ID bits that can differ: the caller passes the mask for foreign pairs
    const uint64_t keep_or = L->id_or, keep_and = L->id_and;
    if (id_mask) {
        L->id_or = id_mask;
        L->id_and = 0;
    }
    CU(L, cudaMemsetAsync(L->d_tot, 0, sizeof(ScanTotals), L->stream));
    const int st = do_finish_pairs(L, n);
    L->pairs_src = nullptr;
    L->pair_later_fixed = 0; // (one call only)
    L->id_or = keep_or;
    L->id_and = keep_and;
    TRY(st);
    if (out_pairs) *out_pairs = L->n_pairs ? L->pout.p : nullptr;
    if (out_count) *out_count = (size_t)L->n_pairs;
    return BP_OK;
}

int bp_dist_partition_records(bp_layer *L, const void *d_keys, const void *d_ids, size_t n, const uint64_t *splitters,
                              int n_splitters, void *d_out_keys, void *d_out_ids, uint64_t *out_counts) {
    if (!L || !out_counts || n_splitters < 0 || n_splitters > MAX_SPLITTERS || (n_splitters && !splitters))
        return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_partition_records");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 records");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_partition_records(L, d_keys, d_ids, n, splitters, n_splitters, d_out_keys, d_out_ids, out_counts);
}

int bp_dist_partition_pairs(bp_layer *L, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                            void *d_out_pairs, uint64_t *out_counts) {
    if (!L || !out_counts || n_splitters < 0 || n_splitters > MAX_SPLITTERS || (n_splitters && !splitters))
        return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_partition_pairs");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 pairs");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_partition_pairs(L, (const uint64_t *)d_pairs, n, splitters, n_splitters, (uint64_t *)d_out_pairs, out_counts);
}

static bool bad_splitters(const uint64_t *splitters, int n) { return n < 0 || n > MAX_SPLITTERS || (n && !splitters); }

int bp_dist_count_records(bp_layer *L, const void *d_keys, size_t n, const uint64_t *splitters, int n_splitters,
                          uint64_t *out_counts, uint64_t *out_halo_counts) {
    if (!L || !out_counts || bad_splitters(splitters, n_splitters)) return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_count_records");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 records");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_count_records(L, d_keys, n, splitters, n_splitters, out_counts, out_halo_counts);
}

int bp_dist_scatter_records(bp_layer *L, const void *d_keys, const void *d_ids, size_t n, const uint64_t *splitters,
                            int n_splitters, const uint64_t *dst_keys, const uint64_t *dst_ids, const uint64_t *halo_dst_keys,
                            const uint64_t *halo_dst_ids) {
    return bp_dist_scatter_records_flagged(L, d_keys, d_ids, n, splitters, n_splitters, dst_keys, dst_ids, halo_dst_keys, halo_dst_ids, 0);
}

int bp_dist_scatter_records_flagged(bp_layer *L, const void *d_keys, const void *d_ids, size_t n, const uint64_t *splitters,
                                    int n_splitters, const uint64_t *dst_keys, const uint64_t *dst_ids, const uint64_t *halo_dst_keys,
                                    const uint64_t *halo_dst_ids, int fold_cell_flags) {
    if (!L || !dst_keys || !dst_ids || bad_splitters(splitters, n_splitters))
        return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_scatter_records");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 records");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_scatter_records(L, d_keys, d_ids, n, splitters, n_splitters, dst_keys, dst_ids, halo_dst_keys, halo_dst_ids,
                              fold_cell_flags != 0 && n > 0);
}

int bp_dist_count_pairs(bp_layer *L, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters, uint64_t *out_counts) {
    if (!L || !out_counts || bad_splitters(splitters, n_splitters)) return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_count_pairs");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 pairs");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_count_pairs(L, (const uint64_t *)d_pairs, n, splitters, n_splitters, out_counts);
}

static bool bad_rows(const uint64_t *tags, int n_tags, const uint64_t *rows, int n_rows) {
    return n_tags < 0 || n_tags > MAX_ROW_TAGS || (n_tags && !tags) || n_rows < 1 || n_rows > MAX_ROW_COPIES || !rows;
}

int bp_dist_count_records_rows(bp_layer *L, const void *d_keys, size_t n, const uint64_t *splitters, int n_splitters,
                               const uint64_t *tags, int n_tags, const uint64_t *d_out_rows, int n_out_rows) {
    if (!L || bad_splitters(splitters, n_splitters) || bad_rows(tags, n_tags, d_out_rows, n_out_rows))
        return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_count_records_rows");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 records");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_count_records_row(L, d_keys, n, splitters, n_splitters, tags, n_tags, d_out_rows, n_out_rows);
}

int bp_dist_extend_count_rows(bp_layer *L, const float *sysb, const float *d_bounds, const void *d_ids, size_t n,
                              const uint64_t *splitters, int n_splitters, int allow_fold, const uint64_t *d_out_rows, int n_out_rows) {
    if (!L || !sysb || (n && (!d_bounds || !d_ids)) || bad_splitters(splitters, n_splitters) || n_out_rows < 1 ||
        n_out_rows > MAX_ROW_COPIES || !d_out_rows)
        return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_extend_count_rows");
    if (L->min_depth != 0) return fail(L, BP_ERR_INVALID_ARG, "bp_dist_extend_count_rows needs a layer with min_depth 0");
    DeviceGuard g(L->device);
    TRY(bp_layer_clear(L));
    return do_extend_count_rows(L, sysb, d_bounds, d_ids, n, splitters, n_splitters, allow_fold != 0, d_out_rows, n_out_rows);
}

int bp_dist_count_pairs_rows(bp_layer *L, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                             const uint64_t *tags, int n_tags, const uint64_t *d_out_rows, int n_out_rows) {
    if (!L || bad_splitters(splitters, n_splitters) || bad_rows(tags, n_tags, d_out_rows, n_out_rows))
        return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_count_pairs_rows");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 pairs");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_count_pairs_row(L, (const uint64_t *)d_pairs, n, splitters, n_splitters, tags, n_tags, d_out_rows, n_out_rows);
}

int bp_dist_scatter_pairs(bp_layer *L, const void *d_pairs, size_t n, const uint64_t *splitters, int n_splitters,
                          const uint64_t *dst_pairs) {
    if (!L || !dst_pairs || bad_splitters(splitters, n_splitters)) return fail(L, BP_ERR_INVALID_ARG, "bad arguments to bp_dist_scatter_pairs");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 pairs");
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_scatter_pairs(L, (const uint64_t *)d_pairs, n, splitters, n_splitters, dst_pairs);
}

int bp_dist_lookup_ranges(bp_layer *L, const void *d_sorted_keys, size_t n, const uint64_t *queries, int n_queries,
                          uint64_t *out_lo, uint64_t *out_hi) {
    if (!L || n_queries < 0 || (n_queries && (!queries || !out_lo || !out_hi))) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    return do_lookup_ranges(L, d_sorted_keys, n, queries, n_queries, out_lo, out_hi);
}

int bp_layer_records_device(bp_layer *L, const void **out_keys, const void **out_ids, size_t *out_n, int *out_sorted) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    TRY(do_strip_flags(L));
    if (out_keys) *out_keys = L->keys[L->cur].p;
    if (out_ids) *out_ids = L->ids[L->cur].p;
    if (out_n) *out_n = (size_t)L->n_records;
    if (out_sorted) *out_sorted = L->dirty ? 0 : 1;
    return BP_OK;
}

int bp_layer_records(bp_layer *L, const void **out_keys, const void **out_ids, size_t *out_n, int *out_sorted) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    TRY(do_strip_flags(L));
    const size_t n = (size_t)L->n_records;
    if (n) {
        TRY(pinned_ensure(L, &L->h_keys, &L->h_keys_cap, n * L->key_bytes));
        TRY(pinned_ensure(L, &L->h_ids, &L->h_ids_cap, n * L->id_bytes));
        CU(L, cudaMemcpyAsync(L->h_keys, L->keys[L->cur].p, n * L->key_bytes, cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaMemcpyAsync(L->h_ids, L->ids[L->cur].p, n * L->id_bytes, cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaMemcpyAsync(L->h_err, L->d_err, sizeof(int), cudaMemcpyDeviceToHost, L->stream));
        CU(L, cudaStreamSynchronize(L->stream));
        if (*L->h_err) {
            cudaMemsetAsync(L->d_err, 0, sizeof(int), L->stream);
            return fail(L, BP_ERR_INTERNAL, "a kernel reported a look-back time-out");
        }
    }
    if (out_keys) *out_keys = n ? L->h_keys : nullptr;
    if (out_ids) *out_ids = n ? L->h_ids : nullptr;
    if (out_n) *out_n = n;
    if (out_sorted) *out_sorted = L->dirty ? 0 : 1;
    return BP_OK;
}

int bp_layer_set_records(bp_layer *L, const void *keys, const void *ids, size_t n, int sorted, int on_device) {
    return bp_layer_set_records_flagged(L, keys, ids, n, sorted, on_device, 0);
}

int bp_layer_set_records_flagged(bp_layer *L, const void *keys, const void *ids, size_t n, int sorted, int on_device, int flagged) {
    if (!L || (n && (!keys || !ids))) return fail(L, BP_ERR_INVALID_ARG, "null argument to set_records");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 records");
    DeviceGuard g(L->device);
    TRY(bp_layer_clear(L));
    if (n) {
        TRY(do_ensure_tree(L, n));
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CU(L, cudaMemcpyAsync(L->keys[L->cur].p, keys, n * L->key_bytes, kind, L->stream));
        CU(L, cudaMemcpyAsync(L->ids[L->cur].p, ids, n * L->id_bytes, kind, L->stream));
    }
    L->n_records = n;
    L->flags_valid = false;        // foreign records: no separate cell flags ...
    L->ids_flagged = flagged != 0; // ... but they may ride in the IDs' top 3 bits (bp_dist_scatter_records_flagged on the sender)
    TRY(do_masks(L, n));
    L->key_or = L->h_res->key_or;
    L->key_and = L->h_res->key_and;
    L->id_or = L->h_res->id_or;
    L->id_and = L->h_res->id_and;
    L->dirty = !sorted;
    L->prefix = 0;
    L->tail_sorted = false;
    L->tail_nonmono = L->h_res->nonmono != 0;
    L->stats.n_records = n;
    return BP_OK;
}

int bp_layer_sort_from_device(bp_layer *L, const void *d_keys, const void *d_ids, size_t n, int flagged, uint64_t key_or,
                              uint64_t key_and, uint64_t id_or, uint64_t id_and, int ids_ascending) {
    if (!L || (n && (!d_keys || !d_ids))) return fail(L, BP_ERR_INVALID_ARG, "null argument to sort_from_device");
    if (n > MAX_RECORDS) return fail(L, BP_ERR_TOO_LARGE, "more than 2^30 records");
    DeviceGuard g(L->device);
    TRY(bp_layer_clear(L));
    if (n) TRY(do_ensure_tree(L, n));
    L->n_records = n;
    L->flags_valid = false;
    L->ids_flagged = flagged != 0;
    L->key_or = key_or;
    L->key_and = key_and;
    L->id_or = id_or;
    L->id_and = id_and;
    L->stats.n_records = n;
    TRY(do_sort_from(L, d_keys, d_ids, n, ids_ascending != 0));
    L->dirty = false;
    return BP_OK;
}

int bp_layer_id_order(bp_layer *L, uint64_t *out_first, uint64_t *out_last, int *out_ascending) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    const bool from_extends = L->n_records > 0 && L->dirty && L->prefix == 0 && !L->tail_sorted;
    if (out_first) *out_first = from_extends ? L->id_first : ~0ull;
    if (out_last) *out_last = from_extends ? L->id_last : 0;
    if (out_ascending) *out_ascending = (L->n_records == 0 || (from_extends && !L->tail_nonmono)) ? 1 : 0;
    return BP_OK;
}

int bp_layer_len(bp_layer *L, size_t *out_n) {
    if (!L || !out_n) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    *out_n = (size_t)L->n_records;
    return BP_OK;
}

int bp_layer_is_sorted(bp_layer *L, int *out_sorted) {
    if (!L || !out_sorted) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    *out_sorted = L->dirty ? 0 : 1;
    return BP_OK;
}

int bp_layer_min_depth(const bp_layer *L, uint32_t *out) {
    if (!L || !out) return BP_ERR_INVALID_ARG;
    *out = L->min_depth;
    return BP_OK;
}

int bp_layer_masks(bp_layer *L, uint64_t *key_or, uint64_t *key_and, uint64_t *id_or, uint64_t *id_and) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    if (key_or) *key_or = L->key_or;
    if (key_and) *key_and = L->key_and;
    if (id_or) *id_or = L->id_or;
    if (id_and) *id_and = L->id_and;
    return BP_OK;
}

int bp_layer_set_profiling(bp_layer *L, int enabled) {
    if (!L) return BP_ERR_INVALID_ARG;
    L->profiling = enabled != 0;
    return BP_OK;
}

int bp_layer_reset_stats(bp_layer *L) {
    if (!L) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    if (!L->prof_events.empty()) {
        CU(L, cudaStreamSynchronize(L->stream));
        collect_profile(L);
    }
    const uint64_t total = L->stats.launches_total;
    memset(&L->stats, 0, sizeof L->stats);
    L->stats.launches_total = total;
    L->stats.n_records = L->n_records;
    return BP_OK;
}

int bp_layer_stats(bp_layer *L, bp_stats *out) {
    if (!L || !out) return BP_ERR_INVALID_ARG;
    DeviceGuard g(L->device);
    TRY(resolve_pending(L));
    if (!L->prof_events.empty()) {
        CU(L, cudaStreamSynchronize(L->stream));
        collect_profile(L);
    }
    L->stats.n_records = L->n_records;
    L->stats.n_invalid = L->n_invalid;
    *out = L->stats;
    return BP_OK;
}

} // extern "C"
