// tools/microbench2.cu -- what limits an insert-heavy shared-memory hash table (B200)?
// Each CTA repeatedly fills a 4096-slot table with KEYS random keys (load factor KEYS/4096) using the
// dependent probe loop of bucket_count_kernel, then clears it.  Variants:
//   0: 64-bit keys, atom.shared.cas.b64        1: 32-bit keys, atom.shared.cas.b32
//   2: 64-bit keys, 2 independent probe chains per thread (ILP 2)
//   3: 64-bit keys, store + verify (no atomics; counts would be approximate: timing only)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int SLOTS = 4096;
__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ uint32_t hash64(unsigned long long k) { uint32_t h = ((uint32_t)k * 0x9E3779B1u) ^ ((uint32_t)(k >> 32) * 0x85EBCA6Bu); h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13; return h; }

template <int MODE>
__global__ void __launch_bounds__(256) fill_kernel(int rounds, int keys, unsigned long long* sink) {
    __shared__ unsigned long long t64[SLOTS];
    __shared__ uint32_t t32[SLOTS];
    const int t = threadIdx.x;
    unsigned long long acc = 0;
    for (int r = 0; r < rounds; r++) {
        for (int i = t; i < SLOTS; i += 256) { t64[i] = ~0ull; t32[i] = ~0u; }
        __syncthreads();
        const uint32_t seed = mix32(blockIdx.x * 7919u + r * 104729u);
        if (MODE == 0 || MODE == 3) {
            for (int i = t; i < keys; i += 256) {
                unsigned long long key = ((unsigned long long)mix32(seed + i) << 20) ^ mix32(seed ^ (i * 2654435761u));
                uint32_t hf = hash64(key), h = hf & (SLOTS - 1), step = ((hf >> 12) | 1u) & (SLOTS - 1);
                for (;;) {
                    unsigned long long old;
                    if (MODE == 0) old = atomicCAS(&t64[h], ~0ull, key);
                    else { old = t64[h]; if (old == ~0ull) { t64[h] = key; __threadfence_block(); old = t64[h] == key ? ~0ull : t64[h]; } }
                    if (old == ~0ull) { acc++; break; }
                    if (old == key) { acc += 2; break; }
                    h = (h + step) & (SLOTS - 1);
                }
            }
        } else if (MODE == 1) {
            for (int i = t; i < keys; i += 256) {
                uint32_t key = mix32(seed + i) & 0x7fffffffu;
                uint32_t hf = mix32(key ^ 0x5bd1e995u), h = hf & (SLOTS - 1), step = ((hf >> 12) | 1u) & (SLOTS - 1);
                for (;;) {
                    uint32_t old = atomicCAS(&t32[h], ~0u, key);
                    if (old == ~0u) { acc++; break; }
                    if (old == key) { acc += 2; break; }
                    h = (h + step) & (SLOTS - 1);
                }
            }
        } else {
            for (int i = t; i < keys; i += 512) {
                unsigned long long key[2]; uint32_t h[2], step[2]; uint32_t pend = 0;
                for (int q = 0; q < 2; q++) {
                    int ii = i + q * 256;
                    key[q] = ((unsigned long long)mix32(seed + ii) << 20) ^ mix32(seed ^ (ii * 2654435761u));
                    uint32_t hf = hash64(key[q]); h[q] = hf & (SLOTS - 1); step[q] = ((hf >> 12) | 1u) & (SLOTS - 1);
                    if (ii < keys) pend |= 1u << q;
                }
                while (pend) {
                    unsigned long long old[2];
                    for (int q = 0; q < 2; q++) old[q] = (pend >> q) & 1 ? atomicCAS(&t64[h[q]], ~0ull, key[q]) : 0ull;
                    for (int q = 0; q < 2; q++) {
                        if (!((pend >> q) & 1)) continue;
                        if (old[q] == ~0ull || old[q] == key[q]) { acc++; pend &= ~(1u << q); }
                        else h[q] = (h[q] + step[q]) & (SLOTS - 1);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (acc == 12345) sink[0] = acc;
}

template <int MODE>
void run(const char* name, int sms, unsigned long long* sink) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int keys : {1000, 2000, 3000})
        for (int ctas : {1, 2, 4}) {
            const int rounds = 200;
            fill_kernel<MODE><<<sms * ctas, 256>>>(rounds, keys, sink);
            cudaEventRecord(e0);
            fill_kernel<MODE><<<sms * ctas, 256>>>(rounds, keys, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double k = (double)sms * ctas * rounds * keys;
            printf("%-34s keys=%d ctas/SM=%d : %.3f ms  %.1f Gkeys/s  %.3f keys/clk/SM  (%.0f cycles per table fill per CTA)\n", name, keys, ctas, ms,
                   k / ms / 1e6, k / (ms * 1e-3) / sms / 1.9e9, ms * 1e-3 * 1.9e9 / rounds);
        }
}

int main() {
    unsigned long long* sink; cudaMalloc(&sink, 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    run<0>("cas64 dependent loop", sms, sink);
    run<1>("cas32 dependent loop", sms, sink);
    run<2>("cas64 two chains per thread", sms, sink);
    run<3>("store+verify 64 (no atomics)", sms, sink);
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
