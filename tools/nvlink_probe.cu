// nvlink_probe.cu -- standalone probe (not part of the product library): what do SM-issued stores into a peer GPU's memory
// reach over NVLink, by access width, by the number of CTAs, mixed with local stores the way the fused partition +
// all-to-all pass of the multi-GPU frame issues them (csrc/bp_radix.cuh, SplitterScatterDigit), and against the copy
// engine and against pulling (loads from the peer)?  One process, devices 0 and 1 with peer access enabled.
// Build: tools/build_tools.sh.  Run: gpurun --gpus 2 -- tools/nvlink_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

// every thread moves one W-byte word per step, consecutive threads consecutive words
template <class W> __global__ void copy_kernel(const W *__restrict__ src, W *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// the scatter pass's store pattern: a tile of TILE records; the first `remote_run` records of every `run` go to the peer,
// the rest stay local; keys (8 bytes) and ids (4 bytes) are separate streams (SPLIT) or one 16-byte record (packed)
constexpr int TILE = 4608, THREADS = 384;
template <bool SPLIT>
__global__ void __launch_bounds__(THREADS) scatter_like_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ ids,
                                                              const ulonglong2 *__restrict__ recs, uint64_t *lk, uint32_t *li,
                                                              ulonglong2 *lr, uint64_t *rk, uint32_t *ri, ulonglong2 *rr,
                                                              uint32_t tiles, uint32_t run, uint32_t remote_run) {
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const size_t base = (size_t)t * TILE;
#pragma unroll
        for (int k = 0; k < TILE / THREADS; ++k) {
            const uint32_t i = k * THREADS + threadIdx.x;
            const bool remote = (i % run) < remote_run;
            if (SPLIT) {
                const uint64_t key = keys[base + i];
                const uint32_t id = ids[base + i];
                (remote ? rk : lk)[base + i] = key;
                (remote ? ri : li)[base + i] = id;
            } else {
                (remote ? rr : lr)[base + i] = recs[base + i];
            }
        }
    }
}

struct Timer {
    cudaEvent_t a, b;
    Timer() {
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
    }
    template <class F> float ms(F f, int reps = 5) {
        f(); // warm-up
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        for (int r = 0; r < reps; ++r) f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float t;
        CK(cudaEventElapsedTime(&t, a, b));
        return t / reps;
    }
};

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        printf("needs 2 GPUs\n");
        return 0;
    }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    printf("peer access 0 -> 1: %d\n", can);
    CK(cudaSetDevice(1));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    const size_t BYTES = (size_t)1 << 30; // per buffer
    void *src0, *dst0, *src1, *dst1;
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&src1, 2 * BYTES));
    CK(cudaMalloc(&dst1, 2 * BYTES));
    CK(cudaMemset(src1, 1, 2 * BYTES));
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&src0, 2 * BYTES));
    CK(cudaMalloc(&dst0, 2 * BYTES));
    CK(cudaMemset(src0, 2, 2 * BYTES));
    CK(cudaDeviceSynchronize());
    Timer T;
    auto gbs = [&](double bytes, float ms) { return bytes / (ms * 1e-3) / 1e9; };

    printf("copy engine  0 -> 1 (cudaMemcpyPeerAsync, 1 GiB): %.0f GB/s\n",
           gbs(BYTES, T.ms([&] { CK(cudaMemcpyPeerAsync(dst1, 1, src0, 0, BYTES, 0)); })));
    printf("local copy kernel (16 B/thread, 1 GiB): %.0f GB/s moved\n",
           gbs(BYTES, T.ms([&] { copy_kernel<uint4><<<148 * 8, 512>>>((const uint4 *)src0, (uint4 *)dst0, BYTES / 16); })));
    for (int blocks : {148, 148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        const float t4 = T.ms([&] { copy_kernel<uint32_t><<<blocks, 512>>>((const uint32_t *)src0, (uint32_t *)dst1, BYTES / 4); });
        const float t8 = T.ms([&] { copy_kernel<uint64_t><<<blocks, 512>>>((const uint64_t *)src0, (uint64_t *)dst1, BYTES / 8); });
        const float t16 = T.ms([&] { copy_kernel<uint4><<<blocks, 512>>>((const uint4 *)src0, (uint4 *)dst1, BYTES / 16); });
        printf("push (SM stores into the peer), %5d CTAs x 512:  4 B/thread %.0f  8 B %.0f  16 B %.0f GB/s\n", blocks, gbs(BYTES, t4),
               gbs(BYTES, t8), gbs(BYTES, t16));
    }
    for (int blocks : {148 * 2, 148 * 8}) {
        const float t8 = T.ms([&] { copy_kernel<uint64_t><<<blocks, 512>>>((const uint64_t *)src1, (uint64_t *)dst0, BYTES / 8); });
        const float t16 = T.ms([&] { copy_kernel<uint4><<<blocks, 512>>>((const uint4 *)src1, (uint4 *)dst0, BYTES / 16); });
        printf("pull (SM loads from the peer), %5d CTAs x 512:  8 B %.0f  16 B %.0f GB/s\n", blocks, gbs(BYTES, t8), gbs(BYTES, t16));
    }
    { // both directions at once: device 1 pushes into device 0 while device 0 pushes into device 1
        cudaStream_t s1;
        CK(cudaSetDevice(1));
        CK(cudaStreamCreate(&s1));
        CK(cudaSetDevice(0));
        const float t = T.ms([&] {
            CK(cudaSetDevice(1));
            copy_kernel<uint4><<<148 * 8, 512, 0, s1>>>((const uint4 *)src1, (uint4 *)dst0 + BYTES / 16, BYTES / 16);
            CK(cudaSetDevice(0));
            copy_kernel<uint4><<<148 * 8, 512>>>((const uint4 *)src0, (uint4 *)dst1, BYTES / 16);
        });
        CK(cudaSetDevice(1));
        CK(cudaStreamSynchronize(s1));
        CK(cudaSetDevice(0));
        printf("push both directions at once (16 B/thread): %.0f GB/s per direction (device 0's clock)\n", gbs(BYTES, t));
    }
    // the scatter pass's pattern: n records of 12 bytes (8 + 4, split streams) or 16 bytes (packed)
    const uint32_t tiles = 20000; // 92 M records
    const size_t n = (size_t)tiles * TILE;
    const uint64_t *keys = (const uint64_t *)src0;
    const uint32_t *ids = (const uint32_t *)((char *)src0 + BYTES);
    uint64_t *lk = (uint64_t *)dst0, *rk = (uint64_t *)dst1;
    uint32_t *li = (uint32_t *)((char *)dst0 + BYTES), *ri = (uint32_t *)((char *)dst1 + BYTES);
    for (int blocks : {148 * 3, 148 * 6}) {
        for (uint32_t frac8 : {0u, 4u, 7u, 8u}) { // eighths of every run that leave the GPU
            for (uint32_t run : {576u, 2304u}) {
                const uint32_t rr = run * frac8 / 8;
                const float ts = T.ms([&] {
                    scatter_like_kernel<true><<<blocks, THREADS>>>(keys, ids, nullptr, lk, li, nullptr, rk, ri, nullptr, tiles, run, rr);
                });
                const float tp = T.ms([&] {
                    scatter_like_kernel<false><<<blocks, THREADS>>>(nullptr, nullptr, (const ulonglong2 *)src0, nullptr, nullptr,
                                                                    (ulonglong2 *)dst0, nullptr, nullptr, (ulonglong2 *)dst1, tiles, run, rr);
                });
                printf("scatter-like %4d CTAs, %u/8 remote, runs of %4u: split 8+4 B %.3f ms (%.0f GB/s out, %.0f GB/s moved) | packed 16 B "
                       "%.3f ms (%.0f GB/s out)\n",
                       blocks, frac8, run, ts, gbs((double)n * 12 * frac8 / 8, ts), gbs((double)n * 12, ts), tp,
                       gbs((double)n * 16 * frac8 / 8, tp));
            }
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
