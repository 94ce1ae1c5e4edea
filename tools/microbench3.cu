// tools/microbench3.cu -- same-address global atomicAdd (64-bit, with return) throughput: how many output-range
// reservations per second can one counter serve?  One lane per warp issues, like the leaf kernel's reservation.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void same_addr(unsigned long long* ctr, int iters, unsigned long long* sink) {
    unsigned long long acc = 0;
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < iters; i++) acc += atomicAdd(ctr, 1ull + (acc & 1));
    if (acc == 12345) sink[0] = acc;
}
int main() {
    unsigned long long *ctr, *sink; cudaMalloc(&ctr, 8); cudaMalloc(&sink, 8); cudaMemset(ctr, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas : {148, 592, 740}) {
        const int iters = 2000;
        same_addr<<<ctas, 256>>>(ctr, iters, sink);
        cudaEventRecord(e0);
        same_addr<<<ctas, 256>>>(ctr, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)ctas * 8 * iters;
        printf("same-address atomicAdd64 with return, %d CTAs x 8 warps (dependent chain per warp): %.3f ms  %.2f Gops/s  (%.2f us per op per warp)\n",
               ctas, ms, n / ms / 1e6, ms * 1e3 / iters);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
