// tools/microbench.cu -- throughput of the primitives the counting kernels lean on (B200, sm_100a).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

constexpr int SLOTS = 4096;
constexpr int ITERS = 4096;

// mode 0: smem atomicCAS64 random slot ; 1: smem atomicAdd32 random ; 2: smem STS64+LDS64 random ; 3: smem atomicCAS32
__global__ void smem_kernel(int mode, unsigned long long* sink) {
    __shared__ unsigned long long t64[SLOTS];
    __shared__ uint32_t t32[SLOTS];
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) { t64[i] = ~0ull; t32[i] = 0; }
    __syncthreads();
    uint32_t x = mix32(blockIdx.x * 1315423911u + threadIdx.x);
    unsigned long long acc = 0;
    for (int it = 0; it < ITERS; it++) {
        x = x * 1664525u + 1013904223u;
        uint32_t slot = (x >> 12) & (SLOTS - 1);
        if (mode == 0) acc += atomicCAS(&t64[slot], ~0ull, (unsigned long long)x);
        else if (mode == 1) acc += atomicAdd(&t32[slot], 1u);
        else if (mode == 2) { t64[slot] = x; acc += t64[(slot + 1) & (SLOTS - 1)]; }
        else acc += atomicCAS(&t32[slot], 0u, x);
        if (mode == 0 && (it & 1023) == 1023) { __syncthreads(); for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) t64[i] = ~0ull; __syncthreads(); }
    }
    if (acc == 12345) sink[0] = acc;
}

// mode 0: global atomicAdd u64 with return, random over n ; 1: same without using the return (RED)
__global__ void gmem_kernel(int mode, unsigned long long* tbl, uint32_t n_mask, unsigned long long* sink) {
    uint32_t x = mix32(blockIdx.x * 1315423911u + threadIdx.x);
    unsigned long long acc = 0;
    for (int it = 0; it < 256; it++) {
        x = x * 1664525u + 1013904223u;
        uint32_t slot = (x >> 8) & n_mask;
        if (mode == 0) acc += atomicAdd(&tbl[slot], 1ull);
        else atomicAdd(&tbl[slot], 1ull);
    }
    if (acc == 12345) sink[0] = acc;
}

int main() {
    unsigned long long *sink, *tbl;
    cudaMalloc(&sink, 8);
    const uint32_t n = 1u << 20;  // 8 MB of counters: L2 resident
    cudaMalloc(&tbl, (size_t)n * 8);
    cudaMemset(tbl, 0, (size_t)n * 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"smem atomicCAS64 random", "smem atomicAdd32 random", "smem STS64+LDS64 random", "smem atomicCAS32 random"};
    for (int ctas = 1; ctas <= 4; ctas *= 2)
        for (int mode = 0; mode < 4; mode++) {
            smem_kernel<<<sms * ctas, 256>>>(mode, sink);
            cudaEventRecord(e0);
            smem_kernel<<<sms * ctas, 256>>>(mode, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)sms * ctas * 256 * ITERS;
            printf("%-28s ctas/SM=%d : %.3f ms  %.2f Gops/s total  %.3f ops/clk/SM (at 1.9 GHz)\n", names[mode], ctas, ms,
                   ops / ms / 1e6, ops / (ms * 1e-3) / sms / 1.9e9);
        }
    for (int mode = 0; mode < 2; mode++) {
        gmem_kernel<<<sms * 8, 256>>>(mode, tbl, n - 1, sink);
        cudaEventRecord(e0);
        gmem_kernel<<<sms * 8, 256>>>(mode, tbl, n - 1, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * 8 * 256 * 256;
        printf("gmem atomicAdd64 %s random over 8 MB : %.3f ms  %.2f Gops/s\n", mode ? "(RED)" : "(return)", ms, ops / ms / 1e6);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
