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
// What does not fit is handled in tiers (skewed / repetitive input):
//   tier 2  a bucket whose region overflowed (extra records go to a spill list) or whose distinct
//           k-mers overflow the shared table is put on a failed list and emits nothing; one more
//           kernel counts exactly those buckets' records in a global hash table and appends them;
//   tier 3  if even the spill list overflows, DevStatus::n_overflow is set and the caller recounts
//           the whole batch with the global-hash-table path (count_hash.cu).
#include <cstdlib>

#include "kernels.cuh"

namespace kmer {

constexpr int LEAF_SLOTS = 4096;          // shared-memory table slots per bucket
constexpr int LEAF_THREADS = 256;
constexpr uint32_t TARGET_KMERS_PER_BUCKET = 2000;   // mean load 0.49 of the table; the tail is handled by tier 2

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
                                                       Rec<RECW>* __restrict__ recs, Rec<RECW>* __restrict__ spill) {
    __shared__ ScanSmem s;
    __shared__ unsigned long long runs[TILE];   // (bucket << 32) | (L << 16) | tile-relative start base
    __shared__ uint32_t wtot[NT / 32];
    TileScanner sc(a, s);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int k = a.k;
    const int m = plan.m;
    const int rmax = plan.rmax;
    const uint32_t mshift = 32 - 2 * m;
    unsigned long long overflow_kmers = 0;

    while (sc.next()) {
        const uint32_t* bnd = sc.bnd();
        // this thread's 16 windows start at tile-relative bases 16t .. 16t+15 and need bases up to 16t+46
        uint32_t w[3];
        w[0] = s.packed[t]; w[1] = s.packed[t + 1]; w[2] = s.packed[t + 2];
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
        uint32_t starts = 0, stops = 0x10000u;
        uint32_t bk[16];
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
            uint32_t prevb = 0;
            int len = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint32_t b = __umulhi(mix32(h[j]), plan.n_buckets);
                bk[j] = b;
                bool v = (vmask >> j) & 1u;
                bool st = v && (len == 0 || b != prevb || len == rmax);
                len = st ? 1 : (v ? len + 1 : 0);
                starts |= (uint32_t)st << j;
                prevb = b;
            }
            stops = starts | ~vmask | 0x10000u;   // a run ends before the next start / invalid / chunk end
        }
        // ---- the tile's runs, compacted into one list (block-wide exclusive scan of the per-thread run counts)
        const uint32_t nrun = __popc(starts);
        uint32_t incl = nrun;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        uint32_t rbase = incl - nrun, n_tile_runs = 0;
#pragma unroll
        for (int q = 0; q < NT / 32; q++) {
            uint32_t v = wtot[q];
            if (q < warp) rbase += v;
            n_tile_runs += v;
        }
        if (vmask) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if ((starts >> j) & 1u) {
                    const uint32_t L = __ffs(stops >> (j + 1));            // 1..16 windows
                    runs[rbase++] = ((unsigned long long)bk[j] << 32) | (L << 16) | (uint32_t)(16 * t + j);
                }
            }
        }
        __syncthreads();
        // ---- flat emission: run r of the tile is handled by thread r % NT; four slot reservations
        //      (64-bit atomicAdd with return) are in flight per thread before any of them is consumed
        for (uint32_t r0 = 0; r0 < n_tile_runs; r0 += 4 * NT) {
            unsigned long long d[4], oldq[4];
            bool act[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t r = r0 + q * NT + t;
                act[q] = r < n_tile_runs;
                d[q] = act[q] ? runs[r] : 0ull;
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                oldq[q] = !act[q] ? 0ull
                          : (plan.debug & 2) ? (unsigned long long)((mix32((uint32_t)d[q] + (uint32_t)sc.tile) >> 8) % plan.cap)
                                             : atomicAdd(&fill[(uint32_t)(d[q] >> 32)], ((d[q] & 0xff0000ull) << 16) | 1ull);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (!act[q]) continue;
                const uint32_t b = (uint32_t)(d[q] >> 32);
                const int L = (int)((d[q] >> 16) & 0xffu);
                const int p = (int)(d[q] & 0xffffu);
                const int c = p >> 4, sh = 2 * (p & 15);
                const uint32_t w0 = s.packed[c], w1 = s.packed[c + 1], w2 = s.packed[c + 2], w3 = s.packed[c + 3];
                const uint32_t r0w = __funnelshift_l(w1, w0, sh), r1w = __funnelshift_l(w2, w1, sh), r2w = __funnelshift_l(w3, w2, sh);
                const int nb = L + k - 1;                                  // bases covered
                const uint32_t slot = (uint32_t)oldq[q];
                Rec<RECW>* dst = nullptr;
                if (slot < plan.cap) dst = recs + ((uint64_t)b * plan.cap + slot);
                else {                                                      // region full: spill list (tier 2), else recount
                    unsigned long long si = atomicAdd(&a.status->n_spill, 1ull);
                    if (si < plan.spill_cap) dst = spill + si;
                    else overflow_kmers += L;
                }
                if (dst && !(plan.debug & 1)) {
                    if (RECW == 1) {
                        uint64_t v = ((uint64_t)r0w << 32) | r1w;
                        v &= ~0ull << (64 - 2 * nb);                       // nb <= 30
                        reinterpret_cast<uint64_t*>(dst)[0] = v | (uint64_t)(L - 1);
                    } else {
                        uint64_t hi = ((uint64_t)r0w << 32) | r1w;
                        uint64_t lo = (uint64_t)r2w << 32;                  // bases 32..47 (nb <= 47)
                        if (nb <= 32) { hi &= ~0ull << (64 - 2 * nb); lo = 0; }
                        else lo &= ~0ull << (128 - 2 * nb);
                        ulonglong2 o; o.x = hi; o.y = lo | (uint64_t)(L - 1);
                        reinterpret_cast<ulonglong2*>(dst)[0] = o;
                    }
                }
            }
        }
        sc.release();
    }
    if (overflow_kmers) atomicAdd(&a.status->n_overflow, overflow_kmers);
}

// ---------------------------------------------------------------------------------------------
// per-bucket counting
//
// One CTA (256 threads, 48 KB of shared memory -> 4 CTAs per SM) per bucket, two block barriers per
// bucket, no intermediate key array:
//   probe : thread t owns records t, t+256, ... of the bucket (up to four are fetched into registers
//           up front, so the global-load latency is paid once per bucket, not inside the loop) and
//           walks their k-mers with PERSISTENT-LANE probing: every loop iteration is one 64-bit
//           shared atomicCAS (double hashing); a thread whose key is placed, or found (then a shared
//           red.add bumps its counter), moves straight on to its next k-mer / next record, so nobody
//           waits for the longest probe sequence of a warp.  The last warp to finish reserves the
//           bucket's output range with ONE global atomicAdd.
//   emit  : the 4096-slot table is scanned two slots per lane (16-byte loads); occupied slots are
//           written as coalesced 16-byte (k-mer, count) pairs and reset on the spot, so the table is
//           clean for the next bucket without a separate initialisation pass.
// Shared memory is addressed through explicit 32-bit shared addresses (ld/st/atom.shared PTX).
// A bucket whose distinct keys overflow the table, or whose region overflowed in the partition pass,
// is appended to the failed list and emits nothing (its k-mers are counted by the tier-2 kernel).

__device__ __forceinline__ unsigned long long atoms_cas64(uint32_t a, unsigned long long cmp, unsigned long long val) {
    unsigned long long old;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t atoms_add32(uint32_t a, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add32(uint32_t a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t a, unsigned long long& x, unsigned long long& y) {
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x), "=l"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

__device__ __forceinline__ uint32_t leaf_hash(uint64_t key) {
    uint32_t h = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA6Bu);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}

struct LeafCounters {
    uint32_t nwin, arrived, failed, special, cursor, pad;
};

template <int RECW>
struct RecRegs {
    uint64_t hi, lo;
};

// record g of bucket b in the flattened (segment-major) order; n_src == 1 is the single-GPU layout
template <int RECW>
__device__ __forceinline__ void load_record(const Rec<RECW>* __restrict__ recs, const unsigned long long* __restrict__ fill,
                                            const PartitionPlan& plan, int n_src, uint32_t b, uint32_t g, uint64_t& hi,
                                            uint64_t& lo) {
    uint32_t seg = 0;
    if (n_src > 1) {
        for (;;) {   // g is below the bucket's total, so this terminates inside the segments
            uint32_t n = (uint32_t)fill[(uint64_t)seg * plan.n_buckets + b];
            if (g < n) break;
            g -= n;
            seg++;
        }
    }
    const Rec<RECW>* p = recs + ((uint64_t)seg * plan.n_buckets + b) * plan.cap + g;
    if (RECW == 1) {
        hi = ld_nc_u64(reinterpret_cast<const uint64_t*>(p));
        lo = 0;
    } else {
        uint4 raw = ld_nc_u128(p);
        hi = ((uint64_t)raw.y << 32) | raw.x;
        lo = ((uint64_t)raw.w << 32) | raw.z;
    }
}

template <int RECW>
__global__ void __launch_bounds__(LEAF_THREADS) bucket_count_kernel(PartitionPlan plan, int k,
                                                                    const unsigned long long* __restrict__ fill,
                                                                    const Rec<RECW>* __restrict__ recs,
                                                                    kmer_count_pair* __restrict__ out, uint64_t capacity,
                                                                    uint32_t* __restrict__ failed_ids, DevStatus* status,
                                                                    int n_src) {
    // n_src > 1 (sharded counting): bucket b's records arrive as n_src segments, one per source GPU:
    // segment s is recs[(s * n_buckets + b) * cap ..] with fill[s * n_buckets + b].
    extern __shared__ __align__(16) unsigned char leaf_dyn[];
    const uint32_t tbl_s = smem_u32(leaf_dyn);                 // u64[LEAF_SLOTS]
    const uint32_t cnt_s = tbl_s + LEAF_SLOTS * 8;             // u32[LEAF_SLOTS]
    __shared__ unsigned long long s_base[2];
    __shared__ LeafCounters s_ctr[2];          // double-buffered by bucket parity: reset while the other set is live
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const int kshift = 64 - 2 * k;
    unsigned long long special_total = 0, kmers_total = 0;
    for (int i = t; i < LEAF_SLOTS / 2; i += LEAF_THREADS) sts128(tbl_s + 16 * i, ~0u, ~0u, ~0u, ~0u);
    for (int i = t; i < LEAF_SLOTS / 4; i += LEAF_THREADS) sts128(cnt_s + 16 * i, 0u, 0u, 0u, 0u);
    if (t < 2) { s_ctr[t].nwin = 0; s_ctr[t].arrived = 0; s_ctr[t].failed = 0; s_ctr[t].special = 0; s_ctr[t].cursor = 0; }
    __syncthreads();
    uint32_t par = 0;

    for (uint32_t b = blockIdx.x; b < plan.n_buckets; b += gridDim.x) {
        uint32_t nrec_all = 0, nk = 0;
        bool seg_overflow = false;
        for (int sI = 0; sI < n_src; sI++) {
            const unsigned long long f = fill[(uint64_t)sI * plan.n_buckets + b];
            nrec_all += (uint32_t)f;
            nk += (uint32_t)(f >> 32);
            seg_overflow |= (uint32_t)f > plan.cap;
        }
        if (nrec_all == 0) continue;                                    // uniform across the CTA
        if (seg_overflow) {                                             // region overflowed in the partition pass: tier 2
            if (t == 0) {
                uint32_t idx = (uint32_t)atomicAdd(&status->n_failed, 1ull);
                failed_ids[idx] = b;
                atomicAdd(&status->failed_kmers, (unsigned long long)nk);
            }
            continue;
        }
        LeafCounters& C = s_ctr[par];
        const uint32_t nwin_s = smem_u32(&C.nwin), arrived_s = smem_u32(&C.arrived), cursor_s = smem_u32(&C.cursor);
        // ---- probe
        uint32_t own = 0, special = 0;
        for (uint32_t g0 = t; g0 < nrec_all; g0 += 4 * LEAF_THREADS) {
            // up to four records of this thread in registers (one global-load latency for all of them)
            uint64_t rh[4], rl[4];
            int nrec = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                rh[q] = 0; rl[q] = 0;
                const uint32_t g = g0 + q * LEAF_THREADS;
                if (g < nrec_all) { load_record<RECW>(recs, fill, plan, n_src, b, g, rh[q], rl[q]); nrec = q + 1; }
            }
            int q = 0, o = 0;
            uint64_t hi = rh[0], lo = rl[0];
            int L = (RECW == 1) ? (int)(hi & 15u) + 1 : (int)(lo & 63u) + 1;
            uint64_t key = hi >> kshift;
            uint32_t hf = leaf_hash(key);
            uint32_t h = hf & (LEAF_SLOTS - 1), step = ((hf >> 12) | 1u) & (LEAF_SLOTS - 1), tries = 0;
            for (;;) {
                bool done = true;
                if (key == kEmpty) special++;                           // k == 32, 't'*32: kept out of the table
                else {
                    unsigned long long old = atoms_cas64(tbl_s + 8 * h, kEmpty, key);
                    if (old == kEmpty) own++;
                    else if (old == key) reds_add32(cnt_s + 4 * h, 1u);
                    else {
                        h = (h + step) & (LEAF_SLOTS - 1);              // double hashing: odd step visits every slot
                        done = false;
                        if (++tries >= LEAF_SLOTS) { C.failed = 1; done = true; }   // table full
                    }
                }
                if (done) {
                    if (++o >= L) {                                     // next record of this thread
                        if (++q >= nrec) break;
                        hi = q == 1 ? rh[1] : (q == 2 ? rh[2] : rh[3]);
                        lo = q == 1 ? rl[1] : (q == 2 ? rl[2] : rl[3]);
                        L = (RECW == 1) ? (int)(hi & 15u) + 1 : (int)(lo & 63u) + 1;
                        o = 0;
                    }
                    const uint64_t win = (RECW == 1 || o == 0) ? (hi << (2 * o)) : ((hi << (2 * o)) | (lo >> (64 - 2 * o)));
                    key = win >> kshift;
                    hf = leaf_hash(key);
                    h = hf & (LEAF_SLOTS - 1);
                    step = ((hf >> 12) | 1u) & (LEAF_SLOTS - 1);
                    tries = 0;
                }
            }
        }
        // ---- the last warp to arrive reserves the bucket's output range
        __syncwarp();
        for (int d = 16; d; d >>= 1) {
            own += __shfl_xor_sync(0xffffffffu, own, d);
            special += __shfl_xor_sync(0xffffffffu, special, d);
        }
        if (lane == 0) {
            if (special) atomicAdd(&C.special, special);
            atoms_add32(nwin_s, own);
            __threadfence_block();
            if (atoms_add32(arrived_s, 1u) == LEAF_THREADS / 32 - 1) {
                const uint32_t total = lds32(nwin_s);
                const bool failed = *reinterpret_cast<volatile uint32_t*>(&C.failed) != 0;
                s_base[par] = (!failed && total) ? atomicAdd(&status->n_distinct, (unsigned long long)total) : 0ull;
            }
        }
        __syncthreads();                                                // (B)
        if (t == 0) {                                                   // the other counter set is idle now: reset it
            LeafCounters& N = s_ctr[par ^ 1];
            N.nwin = 0; N.arrived = 0; N.failed = 0; N.special = 0; N.cursor = 0;
        }
        const bool failed = C.failed != 0;
        if (t == 0) {
            if (failed) {
                uint32_t idx = (uint32_t)atomicAdd(&status->n_failed, 1ull);
                failed_ids[idx] = b;
                atomicAdd(&status->failed_kmers, (unsigned long long)nk);
            } else {
                special_total += C.special;
                kmers_total += nk - C.special;
            }
        }
        // ---- emit + reset: two slots per lane per iteration
        const unsigned long long obase = s_base[par];
        for (int i = t; i < LEAF_SLOTS / 2; i += LEAF_THREADS) {
            unsigned long long k0, k1;
            lds128(tbl_s + 16 * i, k0, k1);
            const bool o0 = k0 != kEmpty, o1 = k1 != kEmpty;
            const uint32_t m0 = __ballot_sync(0xffffffffu, o0), m1 = __ballot_sync(0xffffffffu, o1);
            if (!(m0 | m1)) continue;
            uint32_t pos = 0;
            if (lane == 0) pos = atoms_add32(cursor_s, (uint32_t)(__popc(m0) + __popc(m1)));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (o0 | o1) sts128(tbl_s + 16 * i, ~0u, ~0u, ~0u, ~0u);
            if (o0) {
                const uint32_t c = lds32(cnt_s + 8 * i);
                if (c) sts32(cnt_s + 8 * i, 0u);
                const uint64_t idx = obase + pos + __popc(m0 & lane_lt);
                if (!failed) {
                    if (idx < capacity) { ulonglong2 v; v.x = k0; v.y = 1ull + c; reinterpret_cast<ulonglong2*>(out)[idx] = v; }
                    else status->out_overflow = 1;
                }
            }
            if (o1) {
                const uint32_t c = lds32(cnt_s + 8 * i + 4);
                if (c) sts32(cnt_s + 8 * i + 4, 0u);
                const uint64_t idx = obase + pos + __popc(m0) + __popc(m1 & lane_lt);
                if (!failed) {
                    if (idx < capacity) { ulonglong2 v; v.x = k1; v.y = 1ull + c; reinterpret_cast<ulonglong2*>(out)[idx] = v; }
                    else status->out_overflow = 1;
                }
            }
        }
        __syncthreads();                                                // (D) the table is clean again
        par ^= 1;
    }
    if (t == 0) {
        if (special_total) atomicAdd(&status->special_count, special_total);
        if (kmers_total) atomicAdd(&status->n_kmers, kmers_total);
    }
}

// tier 2: every k-mer of the failed buckets (their in-region records) and of the spill list goes into
// a global open-addressing table (count_hash.cu layout); hash_compact then appends it to the result.
// Failed buckets and spilled records hold k-mers of the same buckets only, so nothing here can also
// have been emitted by bucket_count_kernel.
__device__ __forceinline__ void global_table_add(kmer_count_pair* slots, uint64_t mask, uint64_t code, unsigned long long c) {
    uint64_t h = mix64(code) & mask;
    for (;;) {
        unsigned long long prev = atomicCAS((unsigned long long*)&slots[h].code, kEmpty, code);
        if (prev == kEmpty || prev == code) { atomicAdd((unsigned long long*)&slots[h].count, c); return; }
        h = (h + 1) & mask;
    }
}

template <int RECW>
__device__ __forceinline__ void tier2_add_record(const Rec<RECW>* p, int kshift, kmer_count_pair* slots, uint64_t mask,
                                                 DevStatus* status, bool active) {
    uint64_t hi = 0, lo = 0;
    int L = 0;
    if (active) {
        if (RECW == 1) { hi = reinterpret_cast<const uint64_t*>(p)[0]; L = (int)(hi & 15u) + 1; }
        else { hi = reinterpret_cast<const uint64_t*>(p)[0]; lo = reinterpret_cast<const uint64_t*>(p)[1]; L = (int)(lo & 63u) + 1; }
    }
    const int lane = threadIdx.x & 31;
    int maxL = L;
    for (int d = 16; d; d >>= 1) maxL = max(maxL, __shfl_xor_sync(0xffffffffu, maxL, d));
    for (int o = 0; o < maxL; o++) {
        bool v = o < L;
        uint64_t win = (RECW == 1 || o == 0) ? (hi << (2 * o)) : ((hi << (2 * o)) | (lo >> (64 - 2 * o)));
        uint64_t key = win >> kshift;
        uint32_t vm = __ballot_sync(0xffffffffu, v);
        uint32_t same = __match_any_sync(0xffffffffu, key) & vm;     // repetitive input: aggregate equal keys
        if (v && lane == __ffs(same) - 1) {
            unsigned long long c = __popc(same);
            if (key == kEmpty) atomicAdd(&status->special_count, c);
            else global_table_add(slots, mask, key, c);
        }
    }
}

template <int RECW>
__global__ void __launch_bounds__(256) tier2_insert_kernel(PartitionPlan plan, int k, const unsigned long long* __restrict__ fill,
                                                           const Rec<RECW>* __restrict__ recs, const uint32_t* __restrict__ failed_ids,
                                                           const Rec<RECW>* __restrict__ spill, kmer_count_pair* __restrict__ slots,
                                                           uint64_t mask, DevStatus* status, int n_src) {
    const int kshift = 64 - 2 * k;
    const uint32_t n_failed = (uint32_t)status->n_failed;
    for (uint32_t fi = blockIdx.x; fi < n_failed; fi += gridDim.x) {
        const uint32_t b = failed_ids[fi];
        for (int sI = 0; sI < n_src; sI++) {
            const uint32_t nrec = min((uint32_t)fill[(uint64_t)sI * plan.n_buckets + b], plan.cap);
            const Rec<RECW>* base = recs + ((uint64_t)sI * plan.n_buckets + b) * plan.cap;
            for (uint32_t r0 = 0; r0 < nrec; r0 += blockDim.x) {
                uint32_t r = r0 + threadIdx.x;
                tier2_add_record<RECW>(base + min(r, nrec - 1), kshift, slots, mask, status, r < nrec);
            }
        }
    }
    const uint64_t n_spill = min((uint64_t)status->n_spill, (uint64_t)plan.spill_cap);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r0 = (uint64_t)blockIdx.x * blockDim.x; r0 < n_spill; r0 += stride) {
        uint64_t r = r0 + threadIdx.x;
        tier2_add_record<RECW>(spill + min(r, n_spill - 1), kshift, slots, mask, status, r < n_spill);
    }
}

// appends the k == 32 all-ones key, whose occurrences were kept out of the tables
__global__ void append_special_kernel(kmer_count_pair* out, uint64_t capacity, DevStatus* status) {
    unsigned long long sc = status->special_count;
    if (!sc) return;
    unsigned long long idx = atomicAdd(&status->n_distinct, 1ull);
    if (idx < capacity) { out[idx].code = kEmpty; out[idx].count = sc; }
    else status->out_overflow = 1;
    atomicAdd(&status->n_kmers, sc);
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
    uint64_t nb = (n_kmers + TARGET_KMERS_PER_BUCKET - 1) / TARGET_KMERS_PER_BUCKET;
    if (nb < 1) nb = 1;
    if (nb > 0x7fffffffull) nb = 0x7fffffffull;
    p.n_buckets = (uint32_t)nb;
    p.cap = p.recw == 1 ? 1280u : 896u;
    uint64_t sc = (uint64_t)p.n_buckets * p.cap / 8;          // spill list: 1/8 of the bucket regions
    p.spill_cap = sc < 4096 ? 4096 : sc;
    const char* dbg = getenv("KMER_CUDA_DEBUG_PARTITION");   // profiling experiments only (bit0: no record stores, bit1: no slot atomics)
    p.debug = dbg ? atoi(dbg) : 0;
    return p;
}

size_t partition_record_bytes(const PartitionPlan& p) { return (size_t)p.n_buckets * p.cap * (p.recw == 1 ? 8 : 16); }
size_t partition_spill_bytes(const PartitionPlan& p) { return (size_t)p.spill_cap * (p.recw == 1 ? 8 : 16); }

void launch_partition(const DeviceInfo& di, const ScanArgs& a, const PartitionPlan& p, unsigned long long* d_fill,
                      void* d_recs, void* d_spill, cudaStream_t st) {
    cudaMemsetAsync(d_fill, 0, (size_t)p.n_buckets * sizeof(unsigned long long), st);
    uint64_t n_tiles = (a.n_bases + TILE - 1) / TILE;
    uint64_t grid = (uint64_t)di.sm_count * 5;
    if (grid > n_tiles) grid = n_tiles;
    if (!n_tiles) return;
    if (p.w == 4) partition_kernel<4, 1><<<(unsigned)grid, NT, 0, st>>>(a, p, d_fill, (Rec<1>*)d_recs, (Rec<1>*)d_spill);
    else if (p.w == 8) partition_kernel<8, 1><<<(unsigned)grid, NT, 0, st>>>(a, p, d_fill, (Rec<1>*)d_recs, (Rec<1>*)d_spill);
    else partition_kernel<16, 2><<<(unsigned)grid, NT, 0, st>>>(a, p, d_fill, (Rec<2>*)d_recs, (Rec<2>*)d_spill);
}

// p.n_buckets = buckets counted HERE (all of them on one GPU, the owned range when sharded)
void launch_bucket_count(const DeviceInfo& di, const PartitionPlan& p, int k, int n_src, const unsigned long long* d_fill,
                         const void* d_recs, uint32_t* d_failed_ids, kmer_count_pair* d_pairs, uint64_t capacity,
                         DevStatus* d_status, cudaStream_t st) {
    const size_t leaf_smem = LEAF_SLOTS * (8 + 4);   // 32 + 16 = 48 KB -> 4 CTAs per SM
    uint64_t lgrid = (uint64_t)di.sm_count * 4;
    if (lgrid > p.n_buckets) lgrid = p.n_buckets;
    if (!lgrid) return;
    if (p.recw == 1) {
        cudaFuncSetAttribute(bucket_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem);
        cudaFuncSetAttribute(bucket_count_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        bucket_count_kernel<1><<<(unsigned)lgrid, LEAF_THREADS, leaf_smem, st>>>(p, k, d_fill, (const Rec<1>*)d_recs, d_pairs,
                                                                                capacity, d_failed_ids, d_status, n_src);
    } else {
        cudaFuncSetAttribute(bucket_count_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem);
        cudaFuncSetAttribute(bucket_count_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        bucket_count_kernel<2><<<(unsigned)lgrid, LEAF_THREADS, leaf_smem, st>>>(p, k, d_fill, (const Rec<2>*)d_recs, d_pairs,
                                                                                capacity, d_failed_ids, d_status, n_src);
    }
}

void launch_count_partition(const DeviceInfo& di, const ScanArgs& a, const PartitionPlan& p, unsigned long long* d_fill,
                            void* d_recs, void* d_spill, uint32_t* d_failed_ids, kmer_count_pair* d_pairs, uint64_t capacity,
                            cudaStream_t st, void (*mark)(void*, const char*), void* mark_arg) {
    launch_partition(di, a, p, d_fill, d_recs, d_spill, st);
    if (mark) mark(mark_arg, "minimizer_partition");
    launch_bucket_count(di, p, a.k, 1, d_fill, d_recs, d_failed_ids, d_pairs, capacity, a.status, st);
    if (mark) mark(mark_arg, "bucket_count");
}

// tier 2 (only when the host saw n_failed or n_spill): slots must be cleared to 0xFF (launch_hash_clear)
void launch_partition_tier2(const DeviceInfo& di, const PartitionPlan& p, int k, int n_src, const unsigned long long* d_fill,
                            const void* d_recs, const void* d_spill, const uint32_t* d_failed_ids, kmer_count_pair* d_slots,
                            uint64_t n_slots, DevStatus* d_status, cudaStream_t st) {
    unsigned grid = (unsigned)di.sm_count * 8;
    if (p.recw == 1)
        tier2_insert_kernel<1><<<grid, 256, 0, st>>>(p, k, d_fill, (const Rec<1>*)d_recs, d_failed_ids, (const Rec<1>*)d_spill,
                                                     d_slots, n_slots - 1, d_status, n_src);
    else
        tier2_insert_kernel<2><<<grid, 256, 0, st>>>(p, k, d_fill, (const Rec<2>*)d_recs, d_failed_ids, (const Rec<2>*)d_spill,
                                                     d_slots, n_slots - 1, d_status, n_src);
}

void launch_append_special(kmer_count_pair* d_pairs, uint64_t capacity, DevStatus* d_status, cudaStream_t st) {
    append_special_kernel<<<1, 1, 0, st>>>(d_pairs, capacity, d_status);
}

}  // namespace kmer
