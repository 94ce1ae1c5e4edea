// rows.cu -- row bookkeeping kernels: row-start bit mask, short-row detection, per-tile row index.
#include "kernels.cuh"

namespace kmer {

// One thread per row (plus one for the end sentinel).  A row shorter than k is the reference's
// "Invalid KMER Length" (generate_kmers, kmer.c:310-313); the lowest such row index is kept.
__global__ void rows_prepare_kernel(const uint64_t* __restrict__ off, uint64_t n_rows, uint64_t n_bases, int k,
                                    uint32_t* __restrict__ mask, DevStatus* status) {
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) {  // sentinel: nothing may extend past the end of the input
        atomicOr(&mask[n_bases >> 5], 1u << (n_bases & 31));
        return;
    }
    uint64_t lo = off[r], hi = off[r + 1];
    bool bad = hi < lo || hi > n_bases || (hi - lo) < (uint64_t)k;
    if (bad) atomicMin(&status->short_row, (unsigned long long)r);
    if (lo <= n_bases) atomicOr(&mask[lo >> 5], 1u << (lo & 31));
}

void launch_rows_prepare(const uint64_t* d_off, uint64_t n_rows, uint64_t n_bases, int k, uint32_t* d_mask,
                         uint64_t mask_words, DevStatus* d_status, cudaStream_t st) {
    cudaMemsetAsync(d_mask, 0, mask_words * sizeof(uint32_t), st);
    uint64_t n = n_rows + 1;
    rows_prepare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_off, n_rows, n_bases, k, d_mask, d_status);
}

// tile_row[t] = index of the row that contains flat position t*TILE  (= #rows starting at <= pos, minus 1)
__global__ void tile_row_base_kernel(const uint64_t* __restrict__ off, uint64_t n_rows, uint64_t n_tiles,
                                     uint32_t* __restrict__ tile_row) {
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    uint64_t pos = t * TILE;
    uint64_t lo = 0, hi = n_rows;  // first index with off[idx] > pos
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    tile_row[t] = (uint32_t)(lo - 1);
}

void launch_tile_row_base(const uint64_t* d_off, uint64_t n_rows, uint64_t n_tiles, uint32_t* d_tile_row, cudaStream_t st) {
    if (!n_tiles) return;
    tile_row_base_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, st>>>(d_off, n_rows, n_tiles, d_tile_row);
}

}  // namespace kmer
