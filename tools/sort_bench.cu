// sort_bench.cu -- standalone tuning harness for radix_pass_kernel (not part of the product
// library): sorts n pseudo-random (u64 key, u32 value) records with several (THREADS, ITEMS, MINB)
// shapes, checks the result is sorted and a permutation checksum holds, and prints the per-pass
// time and algorithmic GB/s.  Build: see tools/build_tools.sh.  Run under gpurun.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../broadphase-rs_b200/csrc/bp_radix.cuh"

using namespace bp;

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __host__ inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <class K> __global__ void gen_kernel(K *k, uint32_t *v, uint32_t n, uint64_t mask) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        k[i] = (K)(mix64(i) & mask);
        v[i] = i;
    }
}
template <class K> __global__ void check_kernel(const K *k, const uint32_t *v, uint32_t n, unsigned long long *bad,
                                                unsigned long long *sum, uint64_t mask) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i > 0 && (k[i - 1] > k[i] || (k[i - 1] == k[i] && v[i - 1] > v[i]))) atomicAdd(bad, 1ull);
    if (k[i] != (K)(mix64(v[i]) & mask)) atomicAdd(bad, 1ull); // the payload still belongs to its key
    atomicAdd(sum, (unsigned long long)v[i]);
}

template <class K, class V, int THREADS, int ITEMS, int MINB, int RB = 8>
void run(const char *name, K *k0, V *v0, K *k1, V *v1, uint32_t n, uint64_t mask, uint32_t *scratch, size_t scratch_bytes,
         int *d_err, unsigned long long *d_chk) {
    typedef RadixPassCfg<K, V, THREADS, ITEMS, RB> Cfg;
    constexpr int NB = Cfg::NB;
    RadixPlan plan;
    int np = 0;
    for (uint64_t m = mask; m;) { // same greedy plan as the library
        int sh = __builtin_ctzll(m);
        uint64_t w = (m >> sh) & (uint64_t)(NB - 1);
        plan.shift[np] = sh;
        plan.bits[np] = 64 - __builtin_clzll(w);
        plan.shift2[np] = 0;
        plan.bits2[np] = 0;
        np++;
        if (sh + RB >= 64) break;
        m &= ~((uint64_t)(NB - 1) << sh);
    }
    plan.npasses = np;
    const uint32_t tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
    uint32_t *hist = scratch, *counters = hist + RADIX_MAX_PASSES * NB, *status = counters + 64;
    size_t need = (size_t)(RADIX_MAX_PASSES * NB + 64 + (size_t)np * tiles * NB) * 4;
    if (need > scratch_bytes) {
        printf("%s: scratch too small\n", name);
        return;
    }
    auto kern = radix_pass_kernel<K, V, THREADS, ITEMS, MINB, OneFieldDigit<K>, RB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, Cfg::SMEM_BYTES));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    cudaEvent_t e0, e1, h0, h1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&h0);
    cudaEventCreate(&h1);
    float best = 1e9f, best_hist = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        gen_kernel<K><<<(n + 255) / 256, 256>>>(k0, (uint32_t *)v0, n, mask);
        CK(cudaMemsetAsync(scratch, 0, need));
        cudaEventRecord(h0);
        radix_hist_kernel<K><<<148 * 8, 512, (size_t)np * NB * 4>>>(k0, n, nullptr, plan, hist, NB);
        radix_scan_hist_kernel<NB><<<np, NB>>>(hist);
        cudaEventRecord(h1);
        K *kin = k0, *kout = k1;
        V *vin = v0, *vout = v1;
        cudaEventRecord(e0);
        for (int p = 0; p < np; ++p) {
            RadixPassArgs<K, V, OneFieldDigit<K>> a;
            a.kin = kin; a.kout = kout; a.vin = vin; a.vout = vout;
            a.n_host = n; a.n_dev = nullptr;
            a.ghist_excl = hist + p * NB;
            a.status = status + (size_t)p * tiles * NB;
            a.tile_counter = counters + p;
            a.op.shift = plan.shift[p]; a.op.mask = (1u << plan.bits[p]) - 1u;
            a.vflags = nullptr;
            a.err = d_err;
            kern<<<tiles, THREADS, Cfg::SMEM_BYTES>>>(a);
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms, hms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventElapsedTime(&hms, h0, h1);
        best = ms < best ? ms : best;
        best_hist = hms < best_hist ? hms : best_hist;
        if (rep == 0) {
            CK(cudaMemset(d_chk, 0, 16));
            check_kernel<K><<<(n + 255) / 256, 256>>>(kin, (const uint32_t *)vin, n, d_chk, d_chk + 1, mask);
            unsigned long long h[2];
            int err;
            CK(cudaMemcpy(h, d_chk, 16, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
            unsigned long long want = (unsigned long long)n * (n - 1) / 2;
            if (h[0] || h[1] != want || err) printf("%s: WRONG (bad=%llu sum_ok=%d err=%d)\n", name, h[0], h[1] == want, err);
        }
    }
    const double bytes = 2.0 * n * (sizeof(K) + sizeof(V)) * np;
    printf("%-28s regs=%3d smem=%6zu occ=%d blocks/SM  passes=%d  %.3f ms/pass  %7.1f GB/s  whole sort %.3f ms (+ hist+scan %.3f ms)\n", name,
           fa.numRegs, Cfg::SMEM_BYTES, occ, np, best / np, bytes / (best * 1e-3) / 1e9, best, best_hist);
}

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : (64u << 20);
    uint64_t *k0, *k1;
    uint32_t *v0, *v1, *scratch;
    int *d_err;
    unsigned long long *d_chk;
    const size_t scratch_bytes = 512u << 20;
    CK(cudaMalloc(&k0, (size_t)n * 8));
    CK(cudaMalloc(&k1, (size_t)n * 8));
    CK(cudaMalloc(&v0, (size_t)n * 4));
    CK(cudaMalloc(&v1, (size_t)n * 4));
    CK(cudaMalloc(&scratch, scratch_bytes));
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMalloc(&d_chk, 16));
    CK(cudaMemset(d_err, 0, 4));
    const uint64_t mask = argc > 2 ? strtoull(argv[2], 0, 0) : ((1ull << 62) - 1); // full Index64_3D key
    printf("n = %u records (u64 key + u32 value), key mask 0x%llx\n", n, (unsigned long long)mask);
    const char *only = argc > 3 ? argv[3] : nullptr;
#define RUN(T, I, B)                                                                                     \
    if (!only || strstr("kv " #T "x" #I " minb" #B, only))                                               \
    run<uint64_t, uint32_t, T, I, B>("kv " #T "x" #I " minb" #B, k0, v0, k1, v1, n, mask, scratch, scratch_bytes, d_err, d_chk)
    {   // yardstick only: CUB's DeviceRadixSort (library code, not used by the product)
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, (int)n, 0, 64 - __builtin_clzll(mask));
        void *tmp;
        CK(cudaMalloc(&tmp, tb));
        cudaEvent_t c0, c1;
        cudaEventCreate(&c0);
        cudaEventCreate(&c1);
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            gen_kernel<uint64_t><<<(n + 255) / 256, 256>>>(k0, v0, n, mask);
            cudaEventRecord(c0);
            cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, (int)n, 0, 64 - __builtin_clzll(mask));
            cudaEventRecord(c1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, c0, c1);
            best = ms < best ? ms : best;
        }
        const int np = (64 - __builtin_clzll(mask) + 7) / 8;
        printf("%-28s whole sort %.3f ms = %.3f ms per 8-bit digit incl. histogram; %7.1f GB/s on the (2p+1)*n*12 B model\n",
               "CUB DeviceRadixSort", best, best / np, (2.0 * np + 1.0) * n * 12.0 / (best * 1e-3) / 1e9);
        cudaFree(tmp);
    }
#define RUNB(T, I, B, RBITS)                                                                             \
    if (!only || strstr("kv " #T "x" #I " minb" #B " rb" #RBITS, only))                                  \
    run<uint64_t, uint32_t, T, I, B, RBITS>("kv " #T "x" #I " minb" #B " rb" #RBITS, k0, v0, k1, v1, n, mask, scratch, scratch_bytes, d_err, d_chk)
    RUNB(384, 12, 3, 8);
    RUNB(384, 12, 3, 7);
    RUNB(384, 12, 3, 6);
    RUNB(384, 14, 3, 7);
    RUNB(384, 16, 3, 7);
    RUNB(512, 12, 2, 7);
    RUNB(256, 16, 4, 7);
    RUNB(384, 12, 4, 7);
    if (argc > 4) return 0;
    RUNB(384, 12, 3, 9);
    RUNB(384, 12, 3, 10);
    RUNB(384, 12, 2, 9);
    RUNB(384, 12, 2, 10);
    RUNB(384, 16, 2, 9);
    RUNB(384, 16, 2, 10);
    RUNB(512, 12, 2, 9);
    RUNB(512, 12, 2, 10);
    RUNB(256, 16, 3, 9);
    RUNB(256, 16, 4, 9);
    RUNB(256, 12, 4, 9);
    RUNB(320, 14, 3, 9);
    RUNB(384, 10, 3, 9);
    RUNB(384, 14, 3, 9);
    return 0;
}
