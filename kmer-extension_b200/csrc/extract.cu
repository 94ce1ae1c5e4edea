// extract.cu -- K1: generate_kmers for a whole column (reference kmer.c:289-351).
// Output: every window of every row as a uint64 code, rows in order, positions in order
// (the order a sequential scan of the SRF produces, kmer-tests.sql:263-274).
#include "kernels.cuh"

namespace kmer {

// The valid windows of a tile map to ONE contiguous range of the output (a window that would cross a row end has no index, the
// next valid one continues the numbering), so the tile's codes are first placed in shared memory at (index - lowest index of the
// tile) and then leave as 16-byte stores, every lane of a warp writing 512 contiguous bytes -- 8-byte stores straight from the
// windows ran at half the copy bandwidth (one L2 request per 256 bytes).
__global__ void __launch_bounds__(NT, 4) extract_kernel(ScanArgs a, const uint32_t* __restrict__ tile_row,
                                                     uint64_t* __restrict__ out, uint64_t capacity) {
    __shared__ ScanSmem s;
    __shared__ uint32_t wpre[BND_WORDS + 1];
    __shared__ uint32_t wtot[NT / 32];
    __shared__ __align__(16) uint64_t stage[TILE + 2];
    __shared__ int32_t s_lo, s_hi;
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
        if (t == 0) { s_lo = 0x7fffffff; s_hi = -0x7fffffff; }
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += wtot[w];
        if (t < BND_WORDS - 1) wpre[t] = base + incl - v;
        __syncthreads();
        // row containing t0 already counts a row start exactly at t0
        const uint64_t rb = (uint64_t)tile_row[sc.tile] - (bnd[0] & 1u);
        // output index of every valid window RELATIVE to base0 = index a window at t0 in t0's row would have (32-bit);
        // the tile's lowest and highest
        const uint32_t row0 = wpre[0];
        const uint64_t base0 = sc.t0 - (rb + row0) * km1;
        int32_t rel[TILE / NT];
        int32_t lo = 0x7fffffff, hi = -0x7fffffff;
#pragma unroll
        for (int j = 0; j < TILE / NT; j++) {
            const int i = j * NT + t;
            rel[j] = -0x7fffffff;
            if (sc.valid(i)) {
                const uint32_t within = __popc(bnd[i >> 5] & (0xffffffffu >> (31 - (i & 31))));
                rel[j] = i - (int32_t)((wpre[i >> 5] + within - row0) * (uint32_t)km1);
                lo = min(lo, rel[j]);
                hi = max(hi, rel[j] + 1);
            }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0 && hi > lo) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
        __syncthreads();
        const int32_t tlo = s_lo, thi = s_hi;
        if (thi > tlo && thi - tlo <= TILE) {                        // the tile has valid windows
            const uint64_t olo = base0 + (int64_t)tlo, ohi = base0 + (int64_t)thi;
            const uint32_t shift = (uint32_t)(olo & 1ull);           // keep the parity of the output index: 16-byte alignment
#pragma unroll
            for (int j = 0; j < TILE / NT; j++)
                if (rel[j] != -0x7fffffff) stage[rel[j] - tlo + (int32_t)shift] = sc.code(j * NT + t);
            __syncthreads();
            if (ohi > capacity) {
                if (t == 0) a.status->out_overflow = 1;
            } else {
                const uint32_t n = (uint32_t)(thi - tlo);
                uint32_t p = 0;
                if (shift) { if (t == 0) out[olo] = stage[1]; p = 1; }   // odd first index: one scalar store
                const uint32_t pairs = (n - p) / 2;
                const ulonglong2* src = reinterpret_cast<const ulonglong2*>(stage + p + shift);
                ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + olo + p);
                for (uint32_t q = t; q < pairs; q += NT) dst[q] = src[q];
                if (((n - p) & 1u) && t == 0) out[ohi - 1] = stage[n - 1 + shift];
            }
        }
        sc.release();
    }
}

void launch_extract(const DeviceInfo& di, const ScanArgs& a, const uint32_t* d_tile_row, uint64_t* d_codes,
                    uint64_t capacity, cudaStream_t st) {
    uint64_t n_tiles = (a.n_bases + TILE - 1) / TILE;
    if (!n_tiles) return;
    uint64_t grid = (uint64_t)di.sm_count * 4;
    if (grid > n_tiles) grid = n_tiles;
    extract_kernel<<<(unsigned)grid, NT, 0, st>>>(a, d_tile_row, d_codes, capacity);
}

}  // namespace kmer
