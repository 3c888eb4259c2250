// bp_dist.cu -- the sharded frame behind the C ABI (include/bp.h, "bp_dist_*" context).
//
// No reference counterpart: the reference is one process (Rayon on one host).  This is the B200-native way to run its
// hot path -- extend -> par_sort -> [merge] -> par_scan(_filtered) -- on g GPUs of one NVLink domain, one process per
// GPU: Morton-prefix range sharding.  A frame, on every rank:
//
//   1. encode      the rank's objects (K1), counting the records per destination shard while they are generated when the
//                  splitters are cached from the last frame
//   2. splitters   g-1 key splitters from a sample of every rank's keys (sample sort), kept while the shards stay balanced
//   3. counts      this rank's row of the g x (2g+7) count matrix is stored into EVERY rank's copy by the kernel that
//                  finishes the counts (peer stores); one device barrier later everybody knows where its records go
//   4. exchange    exchange_pass_kernel (bp_exchange.cuh) writes every (tile, shard) run straight into the owner's receive
//                  buffer through its peer mapping -- the partition pass IS the all-to-all -- halo copies alongside
//   5. sort        straight out of the receive buffer, planned from the tag words that travelled with the counts
//   6. scan        over [halo | owned], only pairs whose later record is owned
//   7. dedup       raw pairs range-partitioned on the later ID and exchanged the same way, then sorted + deduplicated
//
// Everything here is assembled from the public entry points of include/bp.h (bp_layer_* / bp_dist_* building blocks) plus
// three tiny kernels (device barrier, sampling, nothing else): the peer mappings are CUDA IPC handles of ONE arena per
// rank -- [barrier flags | sample matrix | count matrices | receive buffers] -- which the caller all-gathers with whatever
// transport it has (MPI, torch.distributed, a file) and hands back: no torch type, no NCCL call, no Python in the path.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bp.h"

namespace {

constexpr int MAX_WORLD = 16;
constexpr int SAMPLES = 8192;        // keys every rank contributes to the splitter sample (8 ranks: shard sizes within ~1 % of the mean)
constexpr int N_TAGS = 7;            // id_or | fold bit, key_or, key_and, id_and, first ID, last ID, IDs ascending
constexpr double REBALANCE_AT = 1.15; // cached splitters are dropped when the fullest shard exceeds the mean by this factor
constexpr uint64_t FOLD_BIT = 1ull << 63;
constexpr uint32_t SPIN_LIMIT = 1u << 26;

struct Blob { // what bp_dist_export writes and bp_dist_connect reads, one per rank
    cudaIpcMemHandle_t mem;
    uint64_t arena_bytes;
    uint64_t layout_check; // every rank must have been created with the same capacities
};

// Device barrier over peer memory: rank `me` raises flag [me] in every rank's flag array to `epoch`, then waits until all g
// flags of its own array have reached it.  The stores of earlier kernels of this stream into peer memory are complete when
// this kernel starts; the fence orders them before the flag as seen from the peers.
__global__ void dist_barrier_kernel(uint64_t *const *flag_arrays, int me, int g, uint64_t epoch, int *err) {
    const int t = threadIdx.x;
    if (t < g) {
        __threadfence_system();
        *((volatile uint64_t *)(flag_arrays[t] + me)) = epoch;
        const volatile uint64_t *mine = (const volatile uint64_t *)(flag_arrays[me] + t);
        uint32_t spins = 0;
        while (*mine < epoch) {
            if (++spins > SPIN_LIMIT) { // about a minute: a peer never arrived
                *err = 2;
                break;
            }
            if (spins > 4096) __nanosleep(1000);
        }
        __threadfence_system();
    }
}

// The barrier and, behind it, this rank's copy of a small matrix handed to the host through pinned, device-mapped memory:
// the words, a system-scope fence, then a sequence number the host spins on.  (A cudaMemcpyAsync + cudaStreamSynchronize
// per matrix cost ~15-20 us of idle GPU each, twice per frame -- what bp_layer's mailbox already avoids for its counts.)
constexpr int MAIL_WORDS = 1024;
struct DistMail {
    uint64_t words[MAIL_WORDS];
    uint32_t err;
    uint32_t seq;
};
__global__ void dist_gather_kernel(uint64_t *const *flag_arrays, int me, int g, uint64_t epoch, int *err, const uint64_t *matrix,
                                   int words, DistMail *mail, uint32_t seq) {
    const int t = threadIdx.x;
    if (t < g) {
        __threadfence_system();
        *((volatile uint64_t *)(flag_arrays[t] + me)) = epoch;
        const volatile uint64_t *mine = (const volatile uint64_t *)(flag_arrays[me] + t);
        uint32_t spins = 0;
        while (*mine < epoch) {
            if (++spins > SPIN_LIMIT) {
                *err = 2;
                break;
            }
            if (spins > 4096) __nanosleep(1000);
        }
        __threadfence_system();
    }
    __syncthreads();
    for (int i = t; i < words; i += blockDim.x) mail->words[i] = ((const volatile uint64_t *)matrix)[i];
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        mail->err = (uint32_t)*err;
        __threadfence_system();
        *((volatile uint32_t *)&mail->seq) = seq;
    }
}

// A regular sample of `n` 64-bit words (every stride-th, at most SAMPLES; the rest of the row is ~0 = "no sample"),
// shifted right by `shift`, plus one trailing word, stored as this rank's row of every rank's sample matrix.
struct RowPtrs {
    uint64_t *p[MAX_WORLD];
};
__global__ void sample_rows_kernel(const uint64_t *src, uint64_t n, uint64_t stride, uint32_t shift, uint64_t trailer, RowPtrs rows, int g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > SAMPLES) return;
    uint64_t v = ~0ull;
    if (i == SAMPLES)
        v = trailer;
    else if ((uint64_t)i * stride < n)
        v = src[(uint64_t)i * stride] >> shift;
    for (int r = 0; r < g; ++r) rows.p[r][i] = v;
}

bool snap_splitters() { // BP_DIST_SNAP_SPLITTERS=0: plain quantiles (measurement aid)
    static const bool on = !(getenv("BP_DIST_SNAP_SPLITTERS") && atoi(getenv("BP_DIST_SNAP_SPLITTERS")) == 0);
    return on;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct bp_dist {
    bp_dist_config cfg;
    int me = 0, g = 1, dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bp_layer *enc = nullptr, *shard = nullptr, *stat = nullptr;
    // the arena and its layout (identical on every rank)
    char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t off_flags = 0, off_sample = 0, off_cm_rec = 0, off_cm_pair = 0, off_rk = 0, off_ri = 0, off_rp = 0;
    int row_rec = 0, row_pair = 0;
    char *peer[MAX_WORLD] = {};
    bool connected = false;
    uint64_t **d_flag_arrays = nullptr; // device array of g pointers
    int *d_err = nullptr;
    uint64_t epoch = 0;
    DistMail *h_mail = nullptr, *d_mail = nullptr; // pinned + device-mapped: the count matrices come back through it
    uint32_t mail_seq = 0;
    uint64_t *h_mat = nullptr; // pinned landing buffer for the matrices
    size_t h_mat_words = 0;
    // protocol state
    bool have_splitters = false, have_pair_splitters = false, have_static = false;
    uint64_t splitters[MAX_WORLD] = {}, pair_splitters[MAX_WORLD] = {};
    uint64_t id_mask = 0, static_id_bits = 0;
    uint64_t static_halo = 0;
    bool reuse_splitters = true, fuse_counts = true, global_dedup_decision = true, trace = false;
    bp_dist_info info;
    std::string error;
};

namespace {

int fail(bp_dist *D, int st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (D) D->error = buf;
    return st;
}

#define DTRY(x)                         \
    do {                                \
        int st__ = (x);                 \
        if (st__ != BP_OK) return st__; \
    } while (0)
// a failing layer call: carry its message along
#define LTRY(D, L, x)                                                          \
    do {                                                                       \
        int st__ = (x);                                                        \
        if (st__ != BP_OK) return fail(D, st__, "%s", bp_layer_last_error(L)); \
    } while (0)
#define DCU(D, x)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (x);                                                                             \
        if (e__ != cudaSuccess) return fail(D, BP_ERR_CUDA, "%s failed: %s", #x, cudaGetErrorString(e__)); \
    } while (0)

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int d) {
        cudaGetDevice(&prev);
        if (prev != d) cudaSetDevice(d);
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void layout(bp_dist *D) {
    const int g = D->g;
    D->row_rec = 2 * g + N_TAGS;
    D->row_pair = g + 1;
    size_t o = 0;
    D->off_flags = o;
    o = align_up(o + MAX_WORLD * sizeof(uint64_t), 256);
    D->off_sample = o;
    o = align_up(o + (size_t)g * (SAMPLES + 1) * sizeof(uint64_t), 256);
    D->off_cm_rec = o;
    o = align_up(o + (size_t)g * D->row_rec * sizeof(uint64_t), 256);
    D->off_cm_pair = o;
    o = align_up(o + (size_t)g * D->row_pair * sizeof(uint64_t), 256);
    D->off_rk = o;
    o = align_up(o + D->cfg.record_capacity * sizeof(uint64_t), 256);
    D->off_ri = o;
    o = align_up(o + D->cfg.record_capacity * sizeof(uint32_t), 256);
    D->off_rp = o;
    o = align_up(o + D->cfg.pair_capacity * sizeof(uint64_t), 256);
    D->arena_bytes = o;
}

int barrier(bp_dist *D) {
    ++D->epoch;
    dist_barrier_kernel<<<1, 32, 0, D->stream>>>(D->d_flag_arrays, D->me, D->g, D->epoch, D->d_err);
    DCU(D, cudaGetLastError());
    return BP_OK;
}

// Barrier, then this rank's copy of a matrix (rows x cols 64-bit words at arena offset `off`) on the host.
int gather(bp_dist *D, size_t off, int rows, int cols, const uint64_t **out) {
    const size_t words = (size_t)rows * cols;
    if (words <= (size_t)MAIL_WORDS) { // the count matrices: barrier + hand-over in one kernel, the host spins on the mailbox
        ++D->epoch;
        const uint32_t seq = ++D->mail_seq;
        dist_gather_kernel<<<1, 256, 0, D->stream>>>(D->d_flag_arrays, D->me, D->g, D->epoch, D->d_err, (const uint64_t *)(D->arena + off),
                                                      (int)words, D->d_mail, seq);
        DCU(D, cudaGetLastError());
        volatile DistMail *mb = D->h_mail;
        for (unsigned spins = 1; mb->seq != seq; ++spins) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xffffu) == 0) { // the stream must still be busy; anything else is an error
                const cudaError_t e = cudaStreamQuery(D->stream);
                if (e == cudaSuccess) {
                    if (mb->seq == seq) break;
                    return fail(D, BP_ERR_INTERNAL, "the stream drained without posting the count matrix");
                }
                if (e != cudaErrorNotReady) return fail(D, BP_ERR_CUDA, "waiting for the count matrix: %s", cudaGetErrorString(e));
            }
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if (D->h_mail->err) return fail(D, BP_ERR_INTERNAL, "device barrier timed out (a peer never arrived)");
        memcpy(D->h_mat, (const void *)D->h_mail->words, words * sizeof(uint64_t));
        *out = D->h_mat;
        return BP_OK;
    }
    DTRY(barrier(D));
    DCU(D, cudaMemcpyAsync(D->h_mat, D->arena + off, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, D->stream));
    int herr = 0;
    DCU(D, cudaMemcpyAsync(&herr, D->d_err, sizeof(int), cudaMemcpyDeviceToHost, D->stream));
    DCU(D, cudaStreamSynchronize(D->stream));
    if (herr) return fail(D, BP_ERR_INTERNAL, "device barrier timed out (a peer never arrived)");
    *out = D->h_mat;
    return BP_OK;
}

void row_addresses(const bp_dist *D, size_t off, int row_words, uint64_t *out) {
    for (int r = 0; r < D->g; ++r) out[r] = (uint64_t)(uintptr_t)(D->peer[r] + off + (size_t)D->me * row_words * sizeof(uint64_t));
}

int bit_length(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }

// parts-1 ascending splitters at the quantiles of a sample (identical on every rank because the gathered sample is).
// A splitter may be ANY value -- it only decides the balance -- so each one is moved to the roundest value (most trailing
// zero bits) whose sample rank stays within 1/32 of a shard's size of its quantile: the keys of a shard [lower, upper) share
// every bit above the highest one in which `lower` and `upper - 1` differ, and with round splitters those are the top
// log2(parts) bits or so, which the shard's sort then never looks at (shard_fixed_bits below; 8 shards of a 30-bit scene:
// 27 varying bits = four 7-bit passes instead of four 8-bit ones).
void choose_splitters(uint64_t *s, size_t m, int parts, uint64_t *out) {
    std::sort(s, s + m);
    const size_t slack = snap_splitters() ? m / (size_t)parts / 32 : 0;
    for (int i = 1; i < parts; ++i) {
        if (m == 0) {
            out[i - 1] = ~0ull;
            continue;
        }
        const size_t t = std::min(m - 1, (size_t)i * m / parts);
        uint64_t v = s[t];
        if (slack) { // (slack < the distance between two quantiles: the windows of neighbouring splitters never overlap)
            const uint64_t lo = s[t - std::min(t, slack)], hi = s[std::min(m - 1, t + slack)];
            if (lo < hi) v = hi & (~0ull << (bit_length(lo ^ hi) - 1)); // in (lo, hi]: hi without the bits below the first difference
        }
        out[i - 1] = v;
    }
}
void choose_splitters(std::vector<uint64_t> &s, int parts, uint64_t *out) { choose_splitters(s.data(), s.size(), parts, out); }

// The bits every key of shard `me` is known to share, from its two splitters alone: *fixed = their positions, *value = the
// bits themselves (both inside `top`'s width and above).  A key v belongs to shard d iff splitters[d-1] <= v < splitters[d]
// (splitter_rank15, bp_common.cuh).
void shard_fixed_bits(const uint64_t *splitters, int parts, int me, uint64_t top, uint64_t *fixed, uint64_t *value) {
    *fixed = 0;
    *value = 0;
    const uint64_t lo = me > 0 ? splitters[me - 1] : 0;
    if ((me < parts - 1 && splitters[me] <= lo) || lo > top) return; // an empty shard (or splitters of an empty sample)
    const uint64_t hi = me < parts - 1 ? std::min(splitters[me] - 1, top) : top; // top = the largest value a key can have
    const int b = bit_length(lo ^ hi); // bits [b, 64) agree between the two ends, hence in everything between them
    if (b >= 64) return;
    *fixed = ~0ull << b;
    *value = lo & *fixed;
}

double imbalance(const uint64_t *col_sums, int g) {
    double sum = 0, mx = 0;
    for (int i = 0; i < g; ++i) {
        sum += (double)col_sums[i];
        mx = std::max(mx, (double)col_sums[i]);
    }
    return sum > 0 ? mx / (sum / g) : 1.0;
}

// All-gathers a regular sample of `n` device words (+ one trailer word per rank) and returns every rank's row.
int sample_gather(bp_dist *D, const uint64_t *d_src, uint64_t n, uint32_t shift, uint64_t trailer, const uint64_t **rows) {
    RowPtrs rp;
    for (int r = 0; r < D->g; ++r) rp.p[r] = (uint64_t *)(D->peer[r] + D->off_sample + (size_t)D->me * (SAMPLES + 1) * sizeof(uint64_t));
    const uint64_t stride = std::max<uint64_t>(1, n / SAMPLES);
    sample_rows_kernel<<<(SAMPLES + 1 + 255) / 256, 256, 0, D->stream>>>(d_src, n, stride, shift, trailer, rp, D->g);
    DCU(D, cudaGetLastError());
    return gather(D, D->off_sample, D->g, SAMPLES + 1, rows);
}

struct Mark {
    bp_dist *D;
    double t0;
    int i = 0;
    explicit Mark(bp_dist *d) : D(d), t0(0) {
        if (D->trace) {
            cudaStreamSynchronize(D->stream);
            t0 = now_ms();
        }
    }
    void operator()() {
        if (!D->trace) return;
        cudaStreamSynchronize(D->stream);
        const double t = now_ms();
        if (i < BP_DIST_PHASES) D->info.phase_ms[i] = t - t0;
        ++i;
        t0 = t;
    }
};

// Steps 3-4 for one set of freshly encoded records: count matrix -> destinations -> exchange.  `mat` = the gathered matrix.
int exchange_records(bp_dist *D, const void *d_keys, const void *d_ids, uint64_t r, const uint64_t *mat, bool fold, uint64_t *n_recv,
                     uint64_t *n_halo) {
    const int g = D->g, me = D->me, row = D->row_rec;
    uint64_t recv[MAX_WORLD] = {}, own_off[MAX_WORLD] = {};
    bool any_halo = false;
    for (int s = 0; s < g; ++s)
        for (int d = 0; d < g; ++d) {
            const uint64_t own = mat[s * row + d], halo = mat[s * row + g + d];
            recv[d] += own + halo;
            if (s < me) own_off[d] += own + halo;
            if (s == me && halo) any_halo = true;
        }
    uint64_t need = 0;
    for (int d = 0; d < g; ++d) need = std::max(need, recv[d]);
    D->info.records_needed = need;
    if (need > D->cfg.record_capacity)
        return fail(D, BP_ERR_TOO_LARGE, "a shard would receive %llu records (record_capacity %llu)", (unsigned long long)need,
                    (unsigned long long)D->cfg.record_capacity);
    uint64_t dk[MAX_WORLD], di[MAX_WORLD], hk[MAX_WORLD], hi[MAX_WORLD];
    for (int d = 0; d < g; ++d) {
        dk[d] = (uint64_t)(uintptr_t)(D->peer[d] + D->off_rk) + 8 * own_off[d];
        di[d] = (uint64_t)(uintptr_t)(D->peer[d] + D->off_ri) + 4 * own_off[d];
        const uint64_t ho = own_off[d] + mat[me * row + d];
        hk[d] = (uint64_t)(uintptr_t)(D->peer[d] + D->off_rk) + 8 * ho;
        hi[d] = (uint64_t)(uintptr_t)(D->peer[d] + D->off_ri) + 4 * ho;
    }
    // (no halo arrays = no halo copy leaves this rank, the usual case: no second pass over the keys)
    LTRY(D, D->enc,
         bp_dist_scatter_records_flagged(D->enc, d_keys, d_ids, r, D->splitters, g - 1, dk, di, any_halo ? hk : nullptr,
                                         any_halo ? hi : nullptr, fold ? 1 : 0));
    DTRY(barrier(D)); // every rank's stores have landed before anybody reads its receive buffer
    *n_recv = recv[me];
    uint64_t h = 0;
    for (int s = 0; s < g; ++s) h += mat[s * row + g + me];
    *n_halo = h;
    return BP_OK;
}

} // namespace

extern "C" {

size_t bp_dist_handle_bytes(void) { return sizeof(Blob); }

int bp_dist_plan_splitters(uint64_t *sample, size_t n, int parts, uint64_t *out_splitters) {
    if ((n && !sample) || parts < 1 || parts > MAX_WORLD || (parts > 1 && !out_splitters)) return BP_ERR_INVALID_ARG;
    choose_splitters(sample, n, parts, out_splitters);
    return BP_OK;
}

int bp_dist_plan_shard_bits(const uint64_t *splitters, int parts, int shard, uint64_t top, uint64_t *out_fixed, uint64_t *out_value) {
    if (parts < 1 || parts > MAX_WORLD || shard < 0 || shard >= parts || (parts > 1 && !splitters) || !out_fixed || !out_value)
        return BP_ERR_INVALID_ARG;
    shard_fixed_bits(splitters, parts, shard, top, out_fixed, out_value);
    return BP_OK;
}

int bp_dist_create(const bp_dist_config *cfg, bp_dist **out) {
    if (!cfg || !out) return BP_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->world < 1 || cfg->world > MAX_WORLD || cfg->rank < 0 || cfg->rank >= cfg->world) return BP_ERR_INVALID_ARG;
    if (cfg->index_kind != BP_INDEX64_2D && cfg->index_kind != BP_INDEX64_3D) return BP_ERR_INVALID_ARG; // 64-bit keys only
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return BP_ERR_CUDA; // no CPU fallback
    }
    int dev = cfg->device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return BP_ERR_CUDA;
    if (dev >= ndev) return BP_ERR_INVALID_ARG;
    bp_dist *D = new bp_dist();
    D->cfg = *cfg;
    D->me = cfg->rank;
    D->g = cfg->world;
    D->dev = dev;
    memset(&D->info, 0, sizeof D->info);
    DevGuard guard(dev);
    auto bail = [&](int st) {
        bp_dist_destroy(D);
        return st;
    };
    if (cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(BP_ERR_CUDA);
    D->own_stream = true;
    bp_layer_config lc;
    memset(&lc, 0, sizeof lc);
    lc.index_kind = cfg->index_kind;
    lc.id_bytes = 4;
    lc.min_depth = cfg->min_depth;
    lc.device = dev;
    for (bp_layer **l : {&D->enc, &D->shard, &D->stat}) {
        const int st = bp_layer_create(&lc, l);
        if (st != BP_OK) return bail(st);
        bp_layer_set_stream(*l, D->stream);
    }
    layout(D);
    if (cudaMalloc((void **)&D->arena, D->arena_bytes) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMemset(D->arena, 0, D->off_rk) != cudaSuccess) return bail(BP_ERR_CUDA); // flags + matrices
    if (cudaMalloc((void **)&D->d_flag_arrays, MAX_WORLD * sizeof(uint64_t *)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMalloc((void **)&D->d_err, sizeof(int)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaMemset(D->d_err, 0, sizeof(int)) != cudaSuccess) return bail(BP_ERR_CUDA);
    D->h_mat_words = (size_t)D->g * std::max<size_t>(SAMPLES + 1, D->row_rec);
    if (cudaMallocHost((void **)&D->h_mat, D->h_mat_words * sizeof(uint64_t)) != cudaSuccess) return bail(BP_ERR_OOM);
    if (cudaHostAlloc((void **)&D->h_mail, sizeof(DistMail), cudaHostAllocMapped) != cudaSuccess) return bail(BP_ERR_OOM);
    memset(D->h_mail, 0, sizeof(DistMail));
    if (cudaHostGetDevicePointer((void **)&D->d_mail, D->h_mail, 0) != cudaSuccess) return bail(BP_ERR_CUDA);
    if (D->g == 1) { // a single rank is its own (only) peer
        D->peer[0] = D->arena;
        uint64_t *fa = (uint64_t *)(D->arena + D->off_flags);
        if (cudaMemcpy(D->d_flag_arrays, &fa, sizeof fa, cudaMemcpyHostToDevice) != cudaSuccess) return bail(BP_ERR_CUDA);
        D->connected = true;
    }
    *out = D;
    return BP_OK;
}

int bp_dist_destroy(bp_dist *D) {
    if (!D) return BP_OK;
    DevGuard guard(D->dev);
    if (D->stream) cudaStreamSynchronize(D->stream);
    for (int r = 0; r < D->g; ++r)
        if (D->peer[r] && r != D->me && D->peer[r] != D->arena) cudaIpcCloseMemHandle(D->peer[r]);
    if (D->enc) bp_layer_destroy(D->enc);
    if (D->shard) bp_layer_destroy(D->shard);
    if (D->stat) bp_layer_destroy(D->stat);
    if (D->arena) cudaFree(D->arena);
    if (D->d_flag_arrays) cudaFree(D->d_flag_arrays);
    if (D->d_err) cudaFree(D->d_err);
    if (D->h_mat) cudaFreeHost(D->h_mat);
    if (D->h_mail) cudaFreeHost(D->h_mail);
    if (D->own_stream && D->stream) cudaStreamDestroy(D->stream);
    delete D;
    return BP_OK;
}

int bp_dist_export(bp_dist *D, void *out_blob) {
    if (!D || !out_blob) return BP_ERR_INVALID_ARG;
    DevGuard guard(D->dev);
    Blob b;
    memset(&b, 0, sizeof b);
    DCU(D, cudaIpcGetMemHandle(&b.mem, D->arena));
    b.arena_bytes = D->arena_bytes;
    b.layout_check = (uint64_t)D->cfg.record_capacity * 1000003ull + D->cfg.pair_capacity * 7ull + (uint64_t)D->g;
    memcpy(out_blob, &b, sizeof b);
    return BP_OK;
}

int bp_dist_connect(bp_dist *D, const void *all_blobs) {
    if (!D || !all_blobs) return BP_ERR_INVALID_ARG;
    if (D->connected) return fail(D, BP_ERR_INVALID_ARG, "already connected");
    DevGuard guard(D->dev);
    const Blob *blobs = (const Blob *)all_blobs;
    uint64_t *fa[MAX_WORLD] = {};
    for (int r = 0; r < D->g; ++r) {
        Blob b;
        memcpy(&b, blobs + r, sizeof b);
        if (b.arena_bytes != D->arena_bytes) return fail(D, BP_ERR_MISMATCH, "rank %d was created with different capacities", r);
        if (r == D->me) {
            D->peer[r] = D->arena;
        } else {
            void *p = nullptr;
            DCU(D, cudaIpcOpenMemHandle(&p, b.mem, cudaIpcMemLazyEnablePeerAccess));
            D->peer[r] = (char *)p;
        }
        fa[r] = (uint64_t *)(D->peer[r] + D->off_flags);
    }
    DCU(D, cudaMemcpy(D->d_flag_arrays, fa, sizeof fa, cudaMemcpyHostToDevice));
    D->connected = true;
    return BP_OK;
}

int bp_dist_set_stream(bp_dist *D, void *cuda_stream) {
    if (!D) return BP_ERR_INVALID_ARG;
    DevGuard guard(D->dev);
    if (D->stream) cudaStreamSynchronize(D->stream);
    for (bp_layer *l : {D->enc, D->shard, D->stat}) bp_layer_set_stream(l, cuda_stream); // (they synchronise their old stream first)
    if (D->own_stream && D->stream) cudaStreamDestroy(D->stream);
    D->own_stream = false;
    D->stream = (cudaStream_t)cuda_stream;
    return BP_OK;
}

int bp_dist_set_option(bp_dist *D, int option, int value) {
    if (!D) return BP_ERR_INVALID_ARG;
    switch (option) {
    case BP_DIST_OPT_REUSE_SPLITTERS: D->reuse_splitters = value != 0; break;
    case BP_DIST_OPT_FUSE_COUNTS: D->fuse_counts = value != 0; break;
    case BP_DIST_OPT_GLOBAL_DEDUP_DECISION: D->global_dedup_decision = value != 0; break;
    case BP_DIST_OPT_TRACE: D->trace = value != 0; break;
    default: return fail(D, BP_ERR_INVALID_ARG, "unknown option %d", option);
    }
    return BP_OK;
}

bp_layer *bp_dist_layer(bp_dist *D, int which) {
    if (!D) return nullptr;
    return which == 0 ? D->enc : which == 1 ? D->shard : which == 2 ? D->stat : nullptr;
}

const char *bp_dist_last_error(const bp_dist *D) { return D ? D->error.c_str() : "null context"; }

int bp_dist_last_info(const bp_dist *D, bp_dist_info *out) {
    if (!D || !out) return BP_ERR_INVALID_ARG;
    *out = D->info;
    return BP_OK;
}

// Shards a static scene once (BASELINE config 4 at N > 1): its records are range-partitioned with splitters sampled from the
// static keys -- which stay FIXED from then on, so every frame's dynamic records are routed to the same owners -- sorted,
// and kept resident; bp_dist_frame then merges them in (Layer::merge, src/layer.rs:127-138) before the scan.
int bp_dist_set_static(bp_dist *D, const float *sysb, const float *d_bounds, const void *d_ids, size_t n) {
    if (!D || !sysb || (n && (!d_bounds || !d_ids))) return fail(D, BP_ERR_INVALID_ARG, "null argument");
    if (!D->connected) return fail(D, BP_ERR_INVALID_ARG, "bp_dist_connect has not been called");
    DevGuard guard(D->dev);
    const int g = D->g;
    LTRY(D, D->enc, bp_layer_clear(D->enc));
    LTRY(D, D->enc, bp_layer_extend_device(D->enc, sysb, d_bounds, d_ids, n));
    const void *dk = nullptr, *di = nullptr;
    size_t r = 0;
    int srt = 0;
    LTRY(D, D->enc, bp_layer_records_device(D->enc, &dk, &di, &r, &srt));
    uint64_t key_or, key_and, id_or, id_and;
    LTRY(D, D->enc, bp_layer_masks(D->enc, &key_or, &key_and, &id_or, &id_and));
    const uint64_t *rows = nullptr;
    DTRY(sample_gather(D, (const uint64_t *)dk, r, 0, id_or, &rows));
    std::vector<uint64_t> sample;
    for (int s = 0; s < g; ++s) {
        for (int i = 0; i < SAMPLES; ++i)
            if (rows[s * (SAMPLES + 1) + i] != ~0ull) sample.push_back(rows[s * (SAMPLES + 1) + i]);
        D->static_id_bits |= rows[s * (SAMPLES + 1) + SAMPLES];
    }
    choose_splitters(sample, g, D->splitters);
    D->have_splitters = true;
    uint64_t tags[N_TAGS] = {};
    uint64_t row_addr[MAX_WORLD];
    row_addresses(D, D->off_cm_rec, D->row_rec, row_addr);
    LTRY(D, D->enc, bp_dist_count_records_rows(D->enc, dk, r, D->splitters, g - 1, tags, N_TAGS, row_addr, g));
    const uint64_t *mat = nullptr;
    DTRY(gather(D, D->off_cm_rec, g, D->row_rec, &mat));
    std::vector<uint64_t> m(mat, mat + (size_t)g * D->row_rec); // (exchange_records' barrier does not touch h_mat, but keep a copy)
    uint64_t n_recv = 0, n_halo = 0;
    DTRY(exchange_records(D, dk, di, r, m.data(), false, &n_recv, &n_halo));
    LTRY(D, D->stat, bp_layer_set_records(D->stat, D->arena + D->off_rk, D->arena + D->off_ri, n_recv, 0, 1));
    LTRY(D, D->stat, bp_layer_sort(D->stat));
    D->static_halo = n_halo;
    D->have_static = true;
    D->info.records_owned = n_recv - n_halo;
    D->info.n_halo = n_halo;
    return BP_OK;
}

// One frame on this rank's objects: clear -> extend -> par_sort -> [merge static] -> par_scan_filtered of the whole
// distributed scene.  *out_d_pairs: this rank's slice of the globally sorted, deduplicated (later, earlier) pair list
// (u32 IDs), valid until the next call; the slices of ranks 0..g-1 concatenated are the reference's scan() vector.
int bp_dist_frame(bp_dist *D, const float *sysb, const float *d_bounds, const void *d_ids, size_t n, const bp_filter *flt,
                  const void **out_d_pairs, size_t *out_count) {
    if (!D || !sysb || (n && (!d_bounds || !d_ids))) return fail(D, BP_ERR_INVALID_ARG, "null argument");
    if (!D->connected) return fail(D, BP_ERR_INVALID_ARG, "bp_dist_connect has not been called");
    DevGuard guard(D->dev);
    const int g = D->g, me = D->me;
    bp_dist_info &I = D->info;
    memset(&I, 0, sizeof I);
    Mark mark(D);

    // 1. encode -- together with step 3's counts when the splitters are already known (cached from the last frame)
    const bool need_splitters = !D->have_splitters || (!D->reuse_splitters && !D->have_static);
    const bool fused = !need_splitters && D->fuse_counts && D->cfg.min_depth == 0;
    I.fused = fused;
    uint64_t row_addr[MAX_WORLD];
    row_addresses(D, D->off_cm_rec, D->row_rec, row_addr);
    std::vector<uint64_t> mat_copy;
    const uint64_t *mat = nullptr;
    const void *dk = nullptr, *di = nullptr;
    size_t r = 0;
    int srt = 0;
    uint64_t id_or = 0;
    if (fused) {
        LTRY(D, D->enc, bp_dist_extend_count_rows(D->enc, sysb, d_bounds, d_ids, n, D->splitters, g - 1, D->have_static ? 0 : 1, row_addr, g));
        DTRY(gather(D, D->off_cm_rec, g, D->row_rec, &mat));
        LTRY(D, D->enc, bp_layer_records_device(D->enc, &dk, &di, &r, &srt));
    } else {
        LTRY(D, D->enc, bp_layer_clear(D->enc));
        LTRY(D, D->enc, bp_layer_extend_device(D->enc, sysb, d_bounds, d_ids, n));
        LTRY(D, D->enc, bp_layer_records_device(D->enc, &dk, &di, &r, &srt));
        uint64_t key_or, key_and, id_and;
        LTRY(D, D->enc, bp_layer_masks(D->enc, &key_or, &key_and, &id_or, &id_and));
    }
    I.records_local = r;
    mark(); // encode

    // 2. key splitters (+ the ID bits, piggybacked) from an all-gathered sample
    if (need_splitters) {
        const uint64_t *rows = nullptr;
        DTRY(sample_gather(D, (const uint64_t *)dk, r, 0, id_or, &rows));
        std::vector<uint64_t> sample;
        uint64_t id_bits = 0;
        for (int s = 0; s < g; ++s) {
            for (int i = 0; i < SAMPLES; ++i)
                if (rows[s * (SAMPLES + 1) + i] != ~0ull) sample.push_back(rows[s * (SAMPLES + 1) + i]);
            id_bits |= rows[s * (SAMPLES + 1) + SAMPLES];
        }
        choose_splitters(sample, g, D->splitters);
        D->have_splitters = true;
        D->id_mask = (1ull << std::max(1, bit_length(id_bits))) - 1;
    }
    mark(); // splitters

    // 3. count, complete the count matrix over NVLink
    if (!fused) {
        const bool can_fold = !D->have_static && id_or < (1ull << 29);
        uint64_t key_or, key_and, ido, id_and, first, last;
        int asc = 0;
        LTRY(D, D->enc, bp_layer_masks(D->enc, &key_or, &key_and, &ido, &id_and));
        LTRY(D, D->enc, bp_layer_id_order(D->enc, &first, &last, &asc));
        const uint64_t tags[N_TAGS] = {id_or | (can_fold ? FOLD_BIT : 0), key_or, key_and, id_and, first, last, (uint64_t)asc};
        LTRY(D, D->enc, bp_dist_count_records_rows(D->enc, dk, r, D->splitters, g - 1, tags, N_TAGS, row_addr, g));
        DTRY(gather(D, D->off_cm_rec, g, D->row_rec, &mat));
    }
    mat_copy.assign(mat, mat + (size_t)g * D->row_rec);
    mat = mat_copy.data();
    const int row = D->row_rec;
    // the sort plan of the receive buffer, from the tag words every source sent with its counts: the buffer holds the sources'
    // chunks in rank order, each a stable partition of the source's records -- its IDs ascend iff every source's do, the
    // sources' ID ranges follow each other in rank order, and no (unordered) halo copies came
    bool flagged = true;
    uint64_t id_bits = 0, p_key_or = 0, p_key_and = ~0ull, p_id_and = ~0ull, halo_in = 0;
    for (int s = 0; s < g; ++s) halo_in += mat[s * row + g + me];
    bool ascending = halo_in == 0, have_prev = false;
    uint64_t prev_last = 0;
    for (int s = 0; s < g; ++s) {
        const uint64_t *t = mat + s * row + 2 * g;
        flagged = flagged && (t[0] & FOLD_BIT) != 0;
        id_bits |= t[0] & ~FOLD_BIT;
        p_key_or |= t[1];
        p_key_and &= t[2];
        p_id_and &= t[3];
        const uint64_t first = t[4], last = t[5];
        const bool asc = t[6] != 0;
        if (first > last && asc) continue; // an empty source
        ascending = ascending && asc && (!have_prev || first >= prev_last);
        prev_last = have_prev ? std::max(prev_last, last) : last;
        have_prev = true;
    }
    if (g > 1 && halo_in == 0 && snap_splitters()) { // (halo copies lie below my lower splitter)
        uint64_t fixed, value;
        shard_fixed_bits(D->splitters, g, me, p_key_or ? ~0ull >> (64 - bit_length(p_key_or)) : 0, &fixed, &value);
        p_key_or &= ~fixed | value;
        p_key_and |= value;
    }
    const uint64_t plan_id_or = id_bits;
    id_bits |= D->static_id_bits;
    D->id_mask |= (1ull << std::max(1, bit_length(id_bits))) - 1; // IDs seen since the splitters were cached
    mark(); // counts

    // 4. exchange (+ halo copies)
    uint64_t n_recv = 0, n_halo = 0;
    DTRY(exchange_records(D, dk, di, r, mat, flagged, &n_recv, &n_halo));
    I.records_owned = n_recv - n_halo;
    mark(); // exchange

    // 5. local sort straight out of the receive buffer; the halo records (all < my lower splitter) end up in front
    LTRY(D, D->shard,
         bp_layer_sort_from_device(D->shard, D->arena + D->off_rk, D->arena + D->off_ri, n_recv, flagged ? 1 : 0, p_key_or, p_key_and,
                                   plan_id_or, p_id_and, ascending ? 1 : 0));
    if (D->have_static) { // Layer::merge of the resident static shard (two sorted runs: one merge-path merge)
        LTRY(D, D->shard, bp_layer_merge(D->shard, D->stat));
        n_halo += D->static_halo;
    }
    I.n_halo = n_halo;
    mark(); // sort

    // 6. shard-local scan; pairs whose later record is a halo record belong to an earlier shard
    auto scan_raw = [&](bool dedup, const void **raw, size_t *p_raw, bool *same) -> int {
        LTRY(D, D->shard, bp_layer_set_halo(D->shard, n_halo));
        LTRY(D, D->shard, bp_layer_set_scan_dedup(D->shard, dedup ? 1 : 0));
        const int st = bp_layer_scan_raw_device(D->shard, flt, raw, p_raw);
        bp_layer_set_scan_dedup(D->shard, 1);
        bp_layer_set_halo(D->shard, 0);
        if (st != BP_OK) return fail(D, st, "%s", bp_layer_last_error(D->shard));
        bp_stats stt;
        LTRY(D, D->shard, bp_layer_stats(D->shard, &stt));
        *same = stt.rescans != 0; // an ID owns nested bounds here: some record is inactive
        return BP_OK;
    };
    const void *raw = nullptr;
    size_t p_raw = 0;
    bool same = false;
    DTRY(scan_raw(true, &raw, &p_raw, &same));
    mark(); // scan

    // 7. global dedup: range-partition the raw pairs on the later ID, exchange, sort + unique
    if (!D->have_pair_splitters || !D->reuse_splitters) {
        const uint64_t *rows = nullptr;
        DTRY(sample_gather(D, (const uint64_t *)raw, p_raw, 32, 0, &rows));
        std::vector<uint64_t> sample;
        for (int s = 0; s < g; ++s)
            for (int i = 0; i < SAMPLES; ++i)
                if (rows[s * (SAMPLES + 1) + i] != ~0ull) sample.push_back(rows[s * (SAMPLES + 1) + i] & 0xffffffffull);
        choose_splitters(sample, g, D->pair_splitters);
        D->have_pair_splitters = true;
    }
    uint64_t prow_addr[MAX_WORLD];
    row_addresses(D, D->off_cm_pair, D->row_pair, prow_addr);
    const int prow = D->row_pair;
    std::vector<uint64_t> pm;
    auto count_pairs = [&]() -> int {
        const uint64_t tag = same ? 1 : 0;
        LTRY(D, D->shard, bp_dist_count_pairs_rows(D->shard, raw, p_raw, D->pair_splitters, g - 1, &tag, 1, prow_addr, g));
        const uint64_t *m = nullptr;
        DTRY(gather(D, D->off_cm_pair, g, prow, &m));
        pm.assign(m, m + (size_t)g * prow);
        return BP_OK;
    };
    DTRY(count_pairs());
    // Dedup at the source is valid only while NO record of the whole scene is inactive: the shard holding a pair's canonical
    // cell skips it there if that record's ID owns an enclosing bound (src/layer.rs:562-564), and the reference then reports
    // the pair from another shared cell -- possibly in a neighbouring shard, which must not have suppressed its copy.  The
    // flag travels with the pair counts; when ANY shard saw an inactive record, every shard whose scan ran with the dedup
    // scans again without it.
    bool any_same = false;
    for (int s = 0; s < g; ++s) any_same = any_same || pm[s * prow + g] != 0;
    if (flagged && D->global_dedup_decision && any_same) {
        if (!same && n_halo == 0) {
            bool s2 = false;
            DTRY(scan_raw(false, &raw, &p_raw, &s2));
            I.rescanned = 1;
        }
        DTRY(count_pairs());
    }
    I.raw_pairs = p_raw;
    mark(); // pair_counts
    uint64_t precv[MAX_WORLD] = {}, poff[MAX_WORLD] = {}, pdst[MAX_WORLD];
    for (int s = 0; s < g; ++s)
        for (int d = 0; d < g; ++d) {
            precv[d] += pm[s * prow + d];
            if (s < me) poff[d] += pm[s * prow + d];
        }
    uint64_t pneed = 0;
    for (int d = 0; d < g; ++d) pneed = std::max(pneed, precv[d]);
    I.pairs_needed = pneed;
    if (pneed > D->cfg.pair_capacity)
        return fail(D, BP_ERR_TOO_LARGE, "a shard would receive %llu raw pairs (pair_capacity %llu)", (unsigned long long)pneed,
                    (unsigned long long)D->cfg.pair_capacity);
    for (int d = 0; d < g; ++d) pdst[d] = (uint64_t)(uintptr_t)(D->peer[d] + D->off_rp) + 8 * poff[d];
    LTRY(D, D->shard, bp_dist_scatter_pairs(D->shard, raw, p_raw, D->pair_splitters, g - 1, pdst));
    DTRY(barrier(D));
    mark(); // pair_exchange
    const void *pairs = nullptr;
    size_t n_pairs = 0;
    if (g > 1 && snap_splitters()) { // the later IDs of my slice lie between two pair splitters: their top bits need no radix pass
        uint64_t fixed, value;
        shard_fixed_bits(D->pair_splitters, g, me, D->id_mask & 0xffffffffull, &fixed, &value);
        LTRY(D, D->shard, bp_layer_set_pair_later_fixed(D->shard, fixed & 0xffffffffull));
    }
    LTRY(D, D->shard, bp_layer_unique_pairs_inplace_device(D->shard, D->arena + D->off_rp, precv[me], D->id_mask, &pairs, &n_pairs));
    I.pairs = n_pairs;
    mark(); // unique

    // cached splitters are recomputed next frame when a shard has drifted too far from the mean
    if (D->reuse_splitters) {
        uint64_t col[MAX_WORLD] = {};
        for (int s = 0; s < g; ++s)
            for (int d = 0; d < g; ++d) col[d] += mat[s * row + d] + mat[s * row + g + d];
        if (!D->have_static && imbalance(col, g) > REBALANCE_AT) { // (with a static layer the record splitters are fixed)
            D->have_splitters = false;
            I.rebalance_records = 1;
        }
        if (imbalance(precv, g) > REBALANCE_AT) {
            D->have_pair_splitters = false;
            I.rebalance_pairs = 1;
        }
    }
    if (out_d_pairs) *out_d_pairs = n_pairs ? pairs : nullptr;
    if (out_count) *out_count = n_pairs;
    return BP_OK;
}

} // extern "C"
