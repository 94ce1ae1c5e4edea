// scan.cuh -- the tile front-end shared by extract / count kernels.
//
// The `dna` column arrives as one flat ASCII stream (1 byte per base, rows back to back) plus a
// row-start bit mask (bit p set <=> a row starts at flat position p; bit n_bases set as sentinel).
// A CTA walks a contiguous range of TILE-base tiles.  Per tile:
//   1. the raw ASCII tile (+32-byte halo) and its slice of the row-start mask are brought into
//      shared memory by the TMA engine (cp.async.bulk, double-buffered behind an mbarrier), so the
//      next tile streams in while this one is processed;
//   2. every thread folds case, validates and 2-bit packs 16 bytes (SWAR, enc16) into `packed`;
//   3. the sink reads any window with two funnel shifts: code(i) = window64(packed, i) >> (64-2k),
//      valid(i) <=> no row start in (i, i+k-1]  (k-mers never span rows: each row is its own
//      generate_kmers call, reference kmer.c:297-328).
#pragma once
#include "common.cuh"

namespace kmer {

constexpr int TILE = 4096;   // bases per tile
constexpr int HALO = 32;     // >= k-1, multiple of 16
constexpr int NT = 256;      // threads per CTA
constexpr int RAW_BYTES = TILE + HALO;               // 4128
constexpr int PACKED_WORDS = RAW_BYTES / 16;         // 258
constexpr int BND_WORDS = 132;                       // (TILE+HALO)/32 + 1 = 130, rounded to 16 bytes
constexpr int MASK_PAD_WORDS = 256;                  // slack words behind the row-start mask

struct ScanArgs {
    const uint8_t* seq;        // flat ASCII, 16-byte aligned
    uint64_t n_bases;
    const uint32_t* row_mask;  // row-start bits, ceil((n_bases+1)/32) + MASK_PAD_WORDS words
    int k;
    DevStatus* status;
    uint64_t tile_begin = 0, tile_end = 0;   // tiles this launch walks (tile_end == 0: all) -- a column that arrives in pieces
};

struct alignas(16) ScanSmem {
    uint8_t raw[2][RAW_BYTES];
    uint32_t bnd[2][BND_WORDS];
    uint32_t packed[PACKED_WORDS + 2];
    uint64_t mbar[2];
};

struct TileScanner {
    const ScanArgs& a;
    ScanSmem& s;
    uint64_t tile, tile_end;   // current / one-past-last tile of this CTA
    uint64_t t0;               // flat position of the current tile's first base
    int stage;
    uint32_t phases;           // bit st = parity to wait for on mbar[st]
    uint32_t kmask;            // (1 << (k-1)) - 1

    __device__ TileScanner(const ScanArgs& a_, ScanSmem& s_) : a(a_), s(s_) {
        const uint64_t all_tiles = (a.n_bases + TILE - 1) / TILE;
        const uint64_t first = a.tile_end ? a.tile_begin : 0, n_tiles = (a.tile_end ? a.tile_end : all_tiles) - first;
        tile = first + n_tiles * blockIdx.x / gridDim.x;
        tile_end = first + n_tiles * (blockIdx.x + 1) / gridDim.x;
        stage = 0;
        phases = 0;
        kmask = (a.k > 1) ? ((1u << (a.k - 1)) - 1u) : 0u;
        if (threadIdx.x == 0) {
            mbar_init(&s.mbar[0], 1);
            mbar_init(&s.mbar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && tile < tile_end) issue(tile, 0);
        tile--;  // next() pre-increments
    }

    // thread 0: start the bulk copies of tile t into stage st
    __device__ void issue(uint64_t t, int st) {
        uint64_t p0 = t * TILE;
        uint64_t avail = ((a.n_bases - p0) + 15) & ~15ull;    // buffer is readable up to the next 16 B
        uint32_t raw_bytes = avail < RAW_BYTES ? (uint32_t)avail : RAW_BYTES;
        mbar_arrive_expect_tx(&s.mbar[st], raw_bytes + BND_WORDS * 4);
        bulk_g2s_stream(s.raw[st], a.seq + p0, raw_bytes, &s.mbar[st], l2_evict_first_policy());
        bulk_g2s(s.bnd[st], a.row_mask + p0 / 32, BND_WORDS * 4, &s.mbar[st]);
    }

    // advance to the next tile; returns false when the CTA's range is exhausted.
    __device__ bool next() {
        tile++;
        if (tile >= tile_end) return false;
        t0 = tile * TILE;
        // everyone is done with the other stage and with `packed` (release() of the previous tile)
        if (threadIdx.x == 0 && tile + 1 < tile_end) issue(tile + 1, stage ^ 1);
        mbar_wait(&s.mbar[stage], (phases >> stage) & 1u);
        phases ^= 1u << stage;
        // 2-bit pack + validate
        for (int c = threadIdx.x; c < PACKED_WORDS; c += NT) {
            uint64_t g = t0 + (uint64_t)c * 16;
            uint32_t word = 0;
            if (g < a.n_bases) {
                uint4 v = *reinterpret_cast<const uint4*>(&s.raw[stage][c * 16]);
                uint32_t bad;
                word = enc16(v, bad);
                if (bad) {
                    int j = first_bad_byte(v);
                    if (g + j < a.n_bases) atomicMin(&a.status->bad_char_pos, (unsigned long long)(g + j));
                }
            }
            s.packed[c] = word;
        }
        if (threadIdx.x < 2) s.packed[PACKED_WORDS + threadIdx.x] = 0;
        __syncthreads();
        return true;
    }

    // all threads: done reading this tile's shared memory
    __device__ void release() {
        __syncthreads();
        stage ^= 1;
    }

    __device__ __forceinline__ const uint32_t* bnd() const { return s.bnd[stage]; }

    // k-mer code of the window starting at tile-relative base i (garbage if !valid(i))
    __device__ __forceinline__ uint64_t code(int i) const { return window64(s.packed, i) >> (64 - 2 * a.k); }

    // window [i, i+k) lies inside one row and inside the input
    __device__ __forceinline__ bool valid(int i) const {
        return (t0 + (uint64_t)i < a.n_bases) && ((bits32(s.bnd[stage], i + 1) & kmask) == 0);
    }
};

}  // namespace kmer
