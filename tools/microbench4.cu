// tools/microbench4.cu -- how fast can records be APPENDED to D regions of HBM, by store width?
// N 8-byte records go to region hash(i) % D at the region's current fill level (slot = i / D: no atomics, a pure store
// test).  Widths: one 8-byte store per thread; one whole 32-byte sector per thread (st.v4.u64); a 64- / 128- / 256-byte
// chunk written by 2 / 4 / 8 neighbouring lanes (32 bytes each).  Decides the write-combining granularity of the
// partition pass.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// LANES lanes cooperate on one chunk of LANES*WORDS 8-byte records
template <int WORDS, int LANES>
__global__ void append(unsigned long long* __restrict__ dst, uint64_t n_chunks, uint32_t D, uint64_t cap_chunks) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x / LANES;
    const int sub = threadIdx.x % LANES;
    for (uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; c < n_chunks; c += stride) {
        const uint32_t r = mix32((uint32_t)c) % D;
        const uint64_t slot = c / D;                      // fill level of the region when chunk c arrives
        if (slot >= cap_chunks) continue;
        unsigned long long* p = dst + ((uint64_t)r * cap_chunks + slot) * (WORDS * LANES) + sub * WORDS;
        const unsigned long long v = c * 0x9E3779B97F4A7C15ull;
        if (WORDS == 1) asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
        else if (WORDS == 2) asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v), "l"(v + 1) : "memory");
        else asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v), "l"(v + 1), "l"(v + 2), "l"(v + 3) : "memory");
    }
}

// the same with a slot atomicAdd (with return) on fill[region] per chunk
template <int WORDS, int LANES>
__global__ void append_atomic(unsigned long long* __restrict__ dst, unsigned long long* __restrict__ fill, uint64_t n_chunks,
                              uint32_t D, uint64_t cap_chunks) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x / LANES;
    const int sub = threadIdx.x % LANES;
    for (uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; c < n_chunks; c += stride) {
        const uint32_t r = mix32((uint32_t)c) % D;
        unsigned long long slot = 0;
        if (sub == 0) slot = atomicAdd(&fill[r], 1ull);
        slot = __shfl_sync(0xffffffffu, slot, (threadIdx.x & 31) - sub);
        if (slot >= cap_chunks) continue;
        unsigned long long* p = dst + ((uint64_t)r * cap_chunks + slot) * (WORDS * LANES) + sub * WORDS;
        const unsigned long long v = c * 0x9E3779B97F4A7C15ull;
        if (WORDS == 1) asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
        else if (WORDS == 2) asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v), "l"(v + 1) : "memory");
        else asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v), "l"(v + 1), "l"(v + 2), "l"(v + 3) : "memory");
    }
}

__global__ void copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    const uint64_t N = 216ull << 20;                      // records (1.8 GB)
    unsigned long long *dst, *fill;
    cudaMalloc(&dst, (N + (64ull << 20)) * 8);
    cudaMalloc(&fill, (1u << 20) * 8);
    {
        uint4 *a, *b; const uint64_t n = 1ull << 26;      // 1 GiB each way
        cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
        float ms = timeit([&] { copy_kernel<<<148 * 8, 256>>>(a, b, n); });
        printf("copy 1 GiB -> 1 GiB: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * n * 16 / ms / 1e6);
        cudaFree(a); cudaFree(b);
    }
    const int grid = 148 * 8;
    for (uint32_t D : {798u, 3192u, 6384u, 25536u, 817152u}) {
#define RUN(W, L)                                                                                                              \
    {                                                                                                                          \
        const uint64_t nc = N / (W * L), capc = nc / D + 1;                                                                    \
        float ms = timeit([&] { append<W, L><<<grid, 256>>>(dst, nc, D, capc); });                                             \
        printf("D=%7u  %3d-byte chunks (%d lanes x %2d B): %.3f ms  %6.1f G records/s  %6.1f G stores/s  %5.0f GB/s\n", D,     \
               W * L * 8, L, W * 8, ms, N / ms / 1e6, nc * (double)L / ms / 1e6, N * 8.0 / ms / 1e6);                          \
    }
        RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(4, 2) RUN(4, 4) RUN(4, 8) RUN(1, 4) RUN(1, 16)
#undef RUN
#define RUNA(W, L)                                                                                                             \
    {                                                                                                                          \
        const uint64_t nc = N / (W * L), capc = nc / D + 1 + nc / D / 4;                                                       \
        cudaMemset(fill, 0, (1u << 20) * 8);                                                                                   \
        float ms = timeit([&] { cudaMemsetAsync(fill, 0, (1u << 20) * 8); append_atomic<W, L><<<grid, 256>>>(dst, fill, nc, D, capc); }); \
        printf("D=%7u  %3d-byte chunks + slot atomic: %.3f ms  %6.1f G records/s\n", D, W * L * 8, ms, N / ms / 1e6);          \
    }
        RUNA(1, 1) RUNA(4, 1) RUNA(4, 4)
#undef RUNA
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
