// count_part.cu -- K2c: GROUP BY k-mer / count(*) for 14 <= k <= 32 by minimizer partitioning.
//
// HBM cannot afford one random 32-byte sector per k-mer (what a global hash table costs), and the
// SMs cannot afford one atomic per k-mer in the partitioning pass.  So the column is regrouped in
// two kernels with a compact intermediate:
//
//  partition_kernel   tile scanner (TMA-staged ASCII -> 2-bit) -> for every window the minimum
//                     hashed m-mer over its first W m-mers -> bucket = mix(min) * NB >> 32.
//                     Identical k-mers have identical minimizers, hence the same bucket.
//                     Consecutive windows that share a bucket are emitted as ONE super-k-mer record
//                     (2 bits per base + a length), appended to the bucket's region with a single
//                     64-bit atomicAdd per record (about 0.3 records per k-mer on random DNA).
//  bucket_count_kernel one CTA per bucket: expand the records, insert every k-mer into a 4096-slot
//                     shared-memory hash table (64-bit atomicCAS), append the distinct
//                     (k-mer, count) pairs to the result with one global atomicAdd per bucket.
//
// HBM traffic: read N bases once, write + read ~2.3 B per k-mer of records, write 16 B per group.
//
// Anything that does not fit (a bucket region overflowing, or a bucket holding more k-mers than its
// shared-memory table takes: highly repetitive input) is only counted in DevStatus::n_overflow; the
// caller then recounts the batch with the global-hash-table path (count_hash.cu).  No result is
// produced from a partially counted batch.
#include "kernels.cuh"

namespace kmer {

constexpr int LEAF_SLOTS = 4096;          // shared-memory table slots per bucket
constexpr int LEAF_MAX_KMERS = 3400;      // refuse buckets above this load (0.83)
constexpr int LEAF_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// record formats
//   RECW == 1 (k <= 26): one uint64 : bases 0..29 in bits 63..4 (first base highest), (L-1) in bits 3..0
//   RECW == 2 (k >= 27): two uint64 : hi = bases 0..31 ; lo = bases 32..60 in bits 63..6, (L-1) in bits 5..0
// L = number of k-mers of the record (1..16), which covers L+k-1 bases.

template <int RECW>
struct Rec;
template <>
struct Rec<1> {
    uint64_t v;
};
template <>
struct alignas(16) Rec<2> {
    uint64_t hi, lo;
};

// ---------------------------------------------------------------------------------------------
// partition

template <int W, int RECW>
__global__ void __launch_bounds__(NT) partition_kernel(ScanArgs a, PartitionPlan plan, unsigned long long* __restrict__ fill,
                                                       Rec<RECW>* __restrict__ recs) {
    __shared__ ScanSmem s;
    __shared__ uint32_t sbucket[TILE];
    TileScanner sc(a, s);
    const int t = threadIdx.x;
    const int k = a.k;
    const int m = plan.m;
    const int rmax = plan.rmax;
    const uint32_t mshift = 32 - 2 * m;
    unsigned long long overflow_kmers = 0;

    while (sc.next()) {
        const uint32_t* bnd = sc.bnd();
        // this thread's 16 windows start at tile-relative bases 16t .. 16t+15 and need bases up to 16t+46
        uint32_t w[4];
        w[0] = s.packed[t]; w[1] = s.packed[t + 1]; w[2] = s.packed[t + 2]; w[3] = s.packed[t + 3];
        // validity of the 16 windows: no row start in (i, i+k-1], and inside the input
        uint32_t vmask = 0;
        {
            const int b0 = 16 * t + 1;
            uint32_t lo = bits32(bnd, b0), hi = bits32(bnd, b0 + 32);
            uint64_t bw = ((uint64_t)hi << 32) | lo;
            uint64_t remaining = a.n_bases > sc.t0 + 16ull * t ? a.n_bases - (sc.t0 + 16ull * t) : 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                bool ok = (((uint32_t)(bw >> j) & sc.kmask) == 0) && ((uint64_t)j < remaining);
                vmask |= (uint32_t)ok << j;
            }
        }
        if (vmask) {
            // hashed m-mers at bases 0 .. 15+W-1 of this chunk
            uint32_t h[16 + W - 1];
#pragma unroll
            for (int j = 0; j < 16 + W - 1; j++) {
                const int q = j >> 4, sh = (j & 15) * 2;
                uint32_t top = __funnelshift_l(w[q + 1], w[q], sh);   // 16 bases starting at base j
                uint32_t mm = top >> mshift;
                uint32_t x = mm * 0x9E3779B1u;
                h[j] = x ^ (x >> 15);
            }
            // sliding minimum over W consecutive m-mers (log-step, W is a power of two)
#pragma unroll
            for (int step = 1; step < W; step <<= 1) {
#pragma unroll
                for (int j = 0; j < 16 + W - 1 - step; j++) h[j] = min(h[j], h[j + step]);
            }
            // bucket of every window; run starts
            uint32_t starts = 0, prevb = 0;
            int len = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t b = __umulhi(mix32(h[j]), plan.n_buckets);
                sbucket[16 * t + j] = b;
                bool v = (vmask >> j) & 1u;
                bool st = v && (len == 0 || b != prevb || len == rmax);
                len = st ? 1 : (v ? len + 1 : 0);
                starts |= (uint32_t)st << j;
                prevb = b;
            }
            // emit one record per run
            const uint32_t stops = starts | ~vmask | 0x10000u;   // a run ends before the next start / invalid / chunk end
            while (starts) {
                const int j = __ffs(starts) - 1;
                starts &= starts - 1;
                const int L = __ffs(stops >> (j + 1));             // 1..16 windows
                const uint32_t b = sbucket[16 * t + j];
                const int sh = 2 * j;
                const uint32_t r0 = __funnelshift_l(w[1], w[0], sh), r1 = __funnelshift_l(w[2], w[1], sh),
                               r2 = __funnelshift_l(w[3], w[2], sh);
                const int nb = L + k - 1;                          // bases covered
                unsigned long long old = atomicAdd(&fill[b], ((unsigned long long)L << 32) | 1ull);
                uint32_t slot = (uint32_t)old;
                if (slot < plan.cap) {
                    if (RECW == 1) {
                        uint64_t v = ((uint64_t)r0 << 32) | r1;
                        v &= ~0ull << (64 - 2 * nb);               // nb <= 30
                        reinterpret_cast<uint64_t*>(recs)[(uint64_t)b * plan.cap + slot] = v | (uint64_t)(L - 1);
                    } else {
                        uint64_t hi = ((uint64_t)r0 << 32) | r1;
                        uint64_t lo = (uint64_t)r2 << 32;          // bases 32..47 (nb <= 47)
                        if (nb <= 32) { hi &= ~0ull << (64 - 2 * nb); lo = 0; }
                        else lo &= ~0ull << (128 - 2 * nb);
                        ulonglong2 o; o.x = hi; o.y = lo | (uint64_t)(L - 1);
                        reinterpret_cast<ulonglong2*>(recs)[(uint64_t)b * plan.cap + slot] = o;
                    }
                } else {
                    overflow_kmers += L;
                }
            }
        }
        sc.release();
    }
    if (overflow_kmers) atomicAdd(&a.status->n_overflow, overflow_kmers);
}

// ---------------------------------------------------------------------------------------------
// per-bucket counting

template <int RECW>
__global__ void __launch_bounds__(LEAF_THREADS) bucket_count_kernel(PartitionPlan plan, int k,
                                                                    const unsigned long long* __restrict__ fill,
                                                                    const Rec<RECW>* __restrict__ recs,
                                                                    kmer_count_pair* __restrict__ out, uint64_t capacity,
                                                                    DevStatus* status) {
    extern __shared__ __align__(16) unsigned char leaf_dyn[];
    unsigned long long* tbl = reinterpret_cast<unsigned long long*>(leaf_dyn);             // [LEAF_SLOTS]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(leaf_dyn + LEAF_SLOTS * sizeof(unsigned long long));  // [LEAF_SLOTS]
    __shared__ uint32_t s_own[LEAF_THREADS / 32];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_cursor;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int kshift = 64 - 2 * k;
    unsigned long long special = 0, total_kmers = 0, skipped = 0;

    for (uint32_t b = blockIdx.x; b < plan.n_buckets; b += gridDim.x) {
        const unsigned long long f = fill[b];
        const uint32_t nrec_all = (uint32_t)f, nk = (uint32_t)(f >> 32);
        if (nrec_all == 0) continue;
        if (nrec_all > plan.cap || nk > LEAF_MAX_KMERS) {   // uniform across the CTA
            if (t == 0) skipped += nk;       // this bucket's k-mers are not counted here: the batch is recounted
            continue;
        }
        for (int i = t; i < LEAF_SLOTS; i += LEAF_THREADS) { tbl[i] = kEmpty; cnt[i] = 0; }
        if (t == 0) s_cursor = 0;
        __syncthreads();
        uint32_t own = 0;
        const Rec<RECW>* base = recs + (uint64_t)b * plan.cap;
        for (uint32_t r = t; r < nrec_all; r += LEAF_THREADS) {
            uint64_t hi, lo = 0;
            int L;
            if (RECW == 1) {
                hi = ld_nc_u64(reinterpret_cast<const uint64_t*>(base) + r);
                L = (int)(hi & 15u) + 1;
            } else {
                uint4 raw = ld_nc_u128(base + r);
                hi = ((uint64_t)raw.y << 32) | raw.x;
                lo = ((uint64_t)raw.w << 32) | raw.z;
                L = (int)(lo & 63u) + 1;
            }
            for (int o = 0; o < L; o++) {
                uint64_t win = (RECW == 1 || o == 0) ? (hi << (2 * o)) : ((hi << (2 * o)) | (lo >> (64 - 2 * o)));
                uint64_t key = win >> kshift;
                if (key == kEmpty) { special++; continue; }           // k == 32, 't'*32
                uint32_t hsh = (uint32_t)(mix64(key)) & (LEAF_SLOTS - 1);
                for (;;) {
                    unsigned long long old = atomicCAS(&tbl[hsh], kEmpty, key);
                    if (old == kEmpty) { own++; break; }
                    if (old == key) { atomicAdd(&cnt[hsh], 1u); break; }
                    hsh = (hsh + 1) & (LEAF_SLOTS - 1);
                }
            }
            total_kmers += L;
        }
        // distinct keys of this bucket -> one reservation in the result
        for (int d = 16; d; d >>= 1) own += __shfl_xor_sync(0xffffffffu, own, d);
        if (lane == 0) s_own[warp] = own;
        __syncthreads();
        if (t == 0) {
            uint32_t tot = 0;
            for (int i = 0; i < LEAF_THREADS / 32; i++) tot += s_own[i];
            s_base = tot ? atomicAdd(&status->n_distinct, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        const unsigned long long obase = s_base;
        for (int i = t; i < LEAF_SLOTS; i += LEAF_THREADS) {
            unsigned long long key = tbl[i];
            bool occ = key != kEmpty;
            uint32_t m = __ballot_sync(0xffffffffu, occ);
            if (!m) continue;
            uint32_t pos = 0;
            if (lane == 0) pos = atomicAdd(&s_cursor, (uint32_t)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (occ) {
                uint64_t idx = obase + pos + __popc(m & ((1u << lane) - 1));
                if (idx < capacity) {
                    ulonglong2 o; o.x = key; o.y = 1ull + cnt[i];
                    reinterpret_cast<ulonglong2*>(out)[idx] = o;
                } else status->out_overflow = 1;
            }
        }
        __syncthreads();   // table is re-initialised for the next bucket
    }
    for (int d = 16; d; d >>= 1) {
        special += __shfl_xor_sync(0xffffffffu, special, d);
        total_kmers += __shfl_xor_sync(0xffffffffu, total_kmers, d);
    }
    if (lane == 0) {
        if (special) atomicAdd(&status->special_count, special);
        if (total_kmers) atomicAdd(&status->n_kmers, total_kmers);
    }
    if (skipped) atomicAdd(&status->n_overflow, skipped);
}

// appends the k == 32 all-ones key, whose occurrences were kept out of the tables
__global__ void append_special_kernel(kmer_count_pair* out, uint64_t capacity, DevStatus* status) {
    unsigned long long sc = status->special_count;
    if (!sc) return;
    unsigned long long idx = atomicAdd(&status->n_distinct, 1ull);
    if (idx < capacity) { out[idx].code = kEmpty; out[idx].count = sc; }
    else status->out_overflow = 1;
}

// ---------------------------------------------------------------------------------------------
// host side

PartitionPlan make_partition_plan(uint64_t n_kmers, int k) {
    PartitionPlan p{};
    p.w = k <= 17 ? 4 : (k <= 26 ? 8 : 16);
    p.recw = k <= 26 ? 1 : 2;
    int m = k - p.w + 1;
    p.m = m > 16 ? 16 : m;
    p.rmax = p.recw == 1 ? (30 - k + 1 > 16 ? 16 : 30 - k + 1) : 16;
    uint64_t nb = (n_kmers + 2399) / 2400;
    if (nb < 1) nb = 1;
    if (nb > 0x7fffffffull) nb = 0x7fffffffull;
    p.n_buckets = (uint32_t)nb;
    p.cap = p.recw == 1 ? 1536u : 1024u;
    return p;
}

size_t partition_record_bytes(const PartitionPlan& p) { return (size_t)p.n_buckets * p.cap * (p.recw == 1 ? 8 : 16); }

void launch_count_partition(const DeviceInfo& di, const ScanArgs& a, const PartitionPlan& p, unsigned long long* d_fill,
                            void* d_recs, kmer_count_pair* d_pairs, uint64_t capacity, cudaStream_t st,
                            void (*mark)(void*, const char*), void* mark_arg) {
    cudaMemsetAsync(d_fill, 0, (size_t)p.n_buckets * sizeof(unsigned long long), st);
    uint64_t n_tiles = (a.n_bases + TILE - 1) / TILE;
    uint64_t grid = (uint64_t)di.sm_count * 6;
    if (grid > n_tiles) grid = n_tiles;
    if (n_tiles) {
        if (p.w == 4) partition_kernel<4, 1><<<(unsigned)grid, NT, 0, st>>>(a, p, d_fill, (Rec<1>*)d_recs);
        else if (p.w == 8) partition_kernel<8, 1><<<(unsigned)grid, NT, 0, st>>>(a, p, d_fill, (Rec<1>*)d_recs);
        else partition_kernel<16, 2><<<(unsigned)grid, NT, 0, st>>>(a, p, d_fill, (Rec<2>*)d_recs);
    }
    if (mark) mark(mark_arg, "minimizer_partition");
    uint64_t lgrid = (uint64_t)di.sm_count * 4;
    if (lgrid > p.n_buckets) lgrid = p.n_buckets;
    const size_t leaf_smem = LEAF_SLOTS * (sizeof(unsigned long long) + sizeof(uint32_t));
    cudaFuncSetAttribute(bucket_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem);
    cudaFuncSetAttribute(bucket_count_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem);
    if (p.recw == 1)
        bucket_count_kernel<1><<<(unsigned)lgrid, LEAF_THREADS, leaf_smem, st>>>(p, a.k, d_fill, (const Rec<1>*)d_recs, d_pairs, capacity, a.status);
    else
        bucket_count_kernel<2><<<(unsigned)lgrid, LEAF_THREADS, leaf_smem, st>>>(p, a.k, d_fill, (const Rec<2>*)d_recs, d_pairs, capacity, a.status);
    append_special_kernel<<<1, 1, 0, st>>>(d_pairs, capacity, a.status);
    if (mark) mark(mark_arg, "bucket_count");
}

}  // namespace kmer
