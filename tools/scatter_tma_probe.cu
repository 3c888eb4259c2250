// scatter_tma_probe.cu -- standalone probe for round 2 (not part of the product library; written at the end of round 1,
// compiled but NOT yet run on a GPU).  Question: the fused partition + all-to-all pass (csrc/bp_radix.cuh,
// SplitterScatterDigit) takes local-pass time PLUS NVLink time (profiles/r1_dist_phases.txt, r1_nvlink_probe.log) --
// backed-up peer stores seem to block the SM's load/store pipe for every warp, so no CTA ranks while another drains.
// Does draining the peer-bound run of a tile with ONE cp.async.bulk (shared -> global, the TMA engine) instead of
// per-thread stores let the "ranking" of the other resident CTAs proceed?
//
// Model of one tile (4608 u64 keys, 384 threads, one tile per CTA, 3 CTAs/SM like the real pass): load the keys
// (coalesced) into shared memory, do `work` rounds of shared-memory + ballot work (stand-in for the ranking: it uses
// the same LSU/MIO path the real ranking uses), then write the first `remote` keys of the tile into the peer's memory
// and the rest into local memory.  mode 0: per-thread stores for both (today's pass).  mode 1: the remote run by one
// bulk copy issued by thread 0, the local run by per-thread stores.  mode 2: both by bulk copies.
// Expectation if the hypothesis holds: mode 0 time ~ t(work) + t(NVLink), mode 1 time ~ max of the two.
// Build: tools/build_tools.sh.  Run: gpurun --gpus 2 -- tools/scatter_tma_probe
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

constexpr int THREADS = 384, ITEMS = 12, TILE = THREADS * ITEMS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared -> global bulk copy (both addresses 16-byte aligned, bytes a multiple of 16), tracked by the thread's bulk group
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 3) tile_kernel(const uint64_t *__restrict__ src, uint64_t *__restrict__ local_dst,
                                                         uint64_t *__restrict__ remote_dst, uint32_t remote, int work,
                                                         unsigned *__restrict__ sink) {
    __shared__ __align__(128) uint64_t stage[TILE];
    const unsigned tid = threadIdx.x;
    const size_t base = (size_t)blockIdx.x * TILE;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) stage[k * THREADS + tid] = src[base + k * THREADS + tid];
    __syncthreads();
    // stand-in for the ranking: dependent shared-memory reads + votes + a shared atomic now and then
    unsigned acc = tid;
    for (int it = 0; it < work; ++it) {
        const uint64_t v = stage[(acc * 33u + (unsigned)it) % TILE];
        acc += (unsigned)__popc(__ballot_sync(0xffffffffu, (v >> (it & 31)) & 1u)) + (unsigned)v;
        if ((it & 15) == 15) atomicAdd((unsigned *)&stage[TILE - 1 - (tid & 7)] + 1, acc & 1u); // (high word of a key nobody ships twice)
    }
    if (acc == 0x12345678u) sink[0] = acc; // keeps the loop alive
    __syncthreads();
    if (MODE == 0) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = k * THREADS + tid;
            (i < remote ? remote_dst : local_dst)[base + i] = stage[i];
        }
    } else {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the generic-proxy writes above, made visible to the TMA engine
        __syncthreads();
        if (MODE == 1) {
            for (uint32_t i = remote + tid; i < (uint32_t)TILE; i += THREADS) local_dst[base + i] = stage[i];
        }
        if (tid == 0) {
            if (remote) bulk_store(remote_dst + base, stage, remote * 8u);
            if (MODE == 2 && remote < (uint32_t)TILE) bulk_store(local_dst + base + remote, stage + remote, ((uint32_t)TILE - remote) * 8u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory may be released once it has been read
        }
        __syncthreads();
    }
}

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        printf("needs 2 GPUs\n");
        return 0;
    }
    CK(cudaSetDevice(1));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    const uint32_t tiles = 20000; // 92 M keys, 737 MB
    const size_t bytes = (size_t)tiles * TILE * 8;
    uint64_t *src, *ldst, *rdst;
    unsigned *sink;
    CK(cudaMalloc(&rdst, bytes));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&src, bytes));
    CK(cudaMalloc(&ldst, bytes));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(src, 0x5a, bytes));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    auto time_ms = [&](auto launch) {
        launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        for (int r = 0; r < 5; ++r) launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float t;
        CK(cudaEventElapsedTime(&t, a, b));
        return t / 5;
    };
    printf("%8s %6s | %10s %10s %10s   (ms per pass over %u tiles; GB/s leaving the GPU in mode 0 / 1)\n", "remote", "work", "stores", "bulk remote",
           "bulk both", tiles);
    for (uint32_t remote : {0u, 2304u, 4032u, 4608u}) { // 0, 1/2, 7/8, all of every tile leaves the GPU
        for (int work : {0, 8, 16, 32, 64}) {
            const float t0 = time_ms([&] { tile_kernel<0><<<tiles, THREADS>>>(src, ldst, rdst, remote, work, sink); });
            const float t1 = time_ms([&] { tile_kernel<1><<<tiles, THREADS>>>(src, ldst, rdst, remote, work, sink); });
            const float t2 = time_ms([&] { tile_kernel<2><<<tiles, THREADS>>>(src, ldst, rdst, remote, work, sink); });
            const double out = (double)tiles * remote * 8;
            printf("%8u %6d | %10.3f %10.3f %10.3f   %6.0f / %6.0f\n", remote, work, t0, t1, t2, out / (t0 * 1e-3) / 1e9, out / (t1 * 1e-3) / 1e9);
        }
    }
    CK(cudaDeviceSynchronize());
    // the bulk path must have moved the same bytes: compare the two destinations of the last configuration with the source
    CK(cudaGetLastError());
    printf("done\n");
    return 0;
}
