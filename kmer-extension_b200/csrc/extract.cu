// extract.cu -- K1: generate_kmers for a whole column (reference kmer.c:289-351).
// Output: every window of every row as a uint64 code, rows in order, positions in order
// (the order a sequential scan of the SRF produces, kmer-tests.sql:263-274).
#include "kernels.cuh"

namespace kmer {

// (Round 2 tried staging a tile's codes in shared memory and storing 16 bytes per lane: 3.3 ms instead of 2.7 ms for 1 GB -- two
// more barriers per tile and 32 KB of shared memory cost more than the wider stores gain; the 8-byte stores of consecutive lanes
// below are already whole 256-byte segments.)
__global__ void __launch_bounds__(NT) extract_kernel(ScanArgs a, const uint32_t* __restrict__ tile_row,
                                                     uint64_t* __restrict__ out, uint64_t capacity) {
    __shared__ ScanSmem s;
    __shared__ uint32_t wpre[BND_WORDS + 1];
    __shared__ uint32_t wtot[NT / 32];
    TileScanner sc(a, s);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint64_t km1 = (uint64_t)(a.k - 1);
    while (sc.next()) {
        const uint32_t* bnd = sc.bnd();
        // exclusive prefix popcount of the row-start words of this tile
        uint32_t v = (t < BND_WORDS - 2) ? __popc(bnd[t]) : 0;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += wtot[w];
        if (t < BND_WORDS - 1) wpre[t] = base + incl - v;
        __syncthreads();
        // row containing t0 already counts a row start exactly at t0
        const uint64_t rb = (uint64_t)tile_row[sc.tile] - (bnd[0] & 1u);
#pragma unroll 4
        for (int j = 0; j < TILE / NT; j++) {
            int i = j * NT + t;
            if (sc.valid(i)) {
                uint64_t g = sc.t0 + i;
                uint32_t within = __popc(bnd[i >> 5] & (0xffffffffu >> (31 - (i & 31))));
                uint64_t row = rb + wpre[i >> 5] + within;
                uint64_t idx = g - row * km1;
                if (idx < capacity) out[idx] = sc.code(i);
                else a.status->out_overflow = 1;
            }
        }
        sc.release();
    }
}

void launch_extract(const DeviceInfo& di, const ScanArgs& a, const uint32_t* d_tile_row, uint64_t* d_codes,
                    uint64_t capacity, cudaStream_t st) {
    uint64_t n_tiles = (a.n_bases + TILE - 1) / TILE;
    if (!n_tiles) return;
    uint64_t grid = (uint64_t)di.sm_count * 6;
    if (grid > n_tiles) grid = n_tiles;
    extract_kernel<<<(unsigned)grid, NT, 0, st>>>(a, d_tile_row, d_codes, capacity);
}

}  // namespace kmer
