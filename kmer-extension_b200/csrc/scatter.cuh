// scatter.cuh -- the two write-combining passes that regroup the column's super-k-mer records by minimizer bucket
// (included by count_part.cu; 14 <= k <= 32).
//
// Measured on B200 (tools/microbench4.cu, profiles/r02_microbench_store_width.txt): appending 8-byte records to their
// regions one store at a time runs at 23-40 G stores/s whatever the number of regions -- the L2 takes one request per
// store -- while whole 32-byte sectors written by one thread run at 160 G records/s and 128-byte lines at 570-630 G
// records/s (5 TB/s), even with 817 k regions.  So records are never stored one by one: they are staged in shared memory
// per destination and leave as whole sectors.  Shared memory bounds the fan-out of one pass (destinations x a few sectors
// each), hence two passes:
//
//   scatter_kernel   tile scanner (TMA-staged ASCII -> 2-bit) -> minimizer of every window -> super-k-mer records ->
//                    D coarse partitions (D <= 1024).  Every CTA writes its OWN segment of every partition, so there is
//                    no global atomic and no cursor outside shared memory / registers.
//   refine2_kernel   one CTA per coarse partition: reads the partition's segments (one per scatter CTA), splits it into
//                    its F fine buckets (F <= 1024), which only this CTA writes: again no global atomics.
//
// Staging: `caps` slots per destination; a record takes slot atomicAdd(cnt[dest]) -- after the tile's (round's) barrier
// the thread that owns the destination writes its whole sectors (st.global.v4.u64) and moves what is left (less than a
// sector) to the front.  A record that finds its destination's slots taken waits in a register for a second round
// behind the flush; if that fails too (a burst only skewed input produces) it goes to the spill list and POISONS its
// fine bucket (bit 31 of the bucket's fill word): the bucket is then counted by tier 2 together with the spilled
// records, so the result stays exact.
#pragma once

namespace kmer {

constexpr unsigned long long kPoison = 0x80000000ull;   // fill word, bit 31: a record of this bucket is on the spill list
constexpr int SCAT_RUNCAP = 256;                          // run-list entries per warp and pass
constexpr int SCAT_DPT = 4;                               // destinations per thread in the flush: D <= 4 * NT
#ifndef SCAT_MINB
#define SCAT_MINB 3
#endif
constexpr int RF2_THREADS = 1024;
constexpr int RF2_RQ = 2;                                 // records per thread and round

// global fine bucket of a record: the bucket of the minimizer of its first window
template <int W, int RECW>
__device__ __forceinline__ uint32_t record_bucket(uint64_t hi, const PartitionPlan& plan) {
    const uint32_t himask = 0xffffffffu << (32 - 2 * plan.m);
    uint32_t hmin = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < W; j++) {
        const uint32_t top = (uint32_t)((hi << (2 * j)) >> 32);
        hmin = min(hmin, mmer_hash(top, himask, plan.m));
    }
    const uint32_t pos = bucket_position<W>(hmin, plan.even);
    return (coarse_bucket(pos, plan.hash_buckets, plan.fine_shift) << plan.fine_shift) | fine_in_coarse(pos, plan.fine_shift);
}

template <int RECW>
__device__ __forceinline__ uint32_t rec_kmers(const Rec<RECW>& r) {
    if constexpr (RECW == 1) return (uint32_t)(r.v & 15u) + 1;
    else return (uint32_t)(r.lo & 63u) + 1;
}

// the record goes to tier 2: spill list + poisoned bucket (or, if even the spill list is full, a full recount)
template <int W, int RECW>
__device__ __forceinline__ void spill_record(const Rec<RECW>& r, const PartitionPlan& plan, unsigned long long* fill, Rec<RECW>* spill,
                                             DevStatus* status) {
    uint64_t hi;
    if constexpr (RECW == 1) hi = r.v; else hi = r.hi;
    const uint32_t g = record_bucket<W, RECW>(hi, plan);
    atomicOr(&fill[g], kPoison);
    const unsigned long long si = atomicAdd(&status->n_spill, 1ull);
    if (si < plan.spill_cap) spill[si] = r;
    else atomicAdd(&status->n_overflow, (unsigned long long)rec_kmers<RECW>(r));
}

// ---------------------------------------------------------------------------------------------
// staging area in shared memory:  cnt[D] records staged | gpos[D] records already in the destination's region |
// rdy[2][D] + rdy_n[2] destinations with a whole sector, per round parity | slots [caps][D] (slot-major)
// A "round" is: every thread put()s its records -- barrier -- flush_round().  The lane whose record completes a destination's
// first sector of the round lists the destination, so the flush walks a DENSE list (every lane busy) instead of polling all
// destinations.
template <int RECW>
struct Stage {
    uint32_t cnt_s, gpos_s, rdy_s, rdyn_s, slot_s, D, caps, rnd;
    static constexpr uint32_t RECB = RECW * 8;
    static constexpr uint32_t SECT = 4 / RECW;            // records per 32-byte sector
    __device__ __forceinline__ void init(uint32_t base_s, uint32_t D_, uint32_t caps_) {
        D = D_; caps = caps_; rnd = 0;
        cnt_s = base_s;
        gpos_s = cnt_s + D * 4;
        rdy_s = gpos_s + D * 4;                            // u16 [2][D]
        rdyn_s = rdy_s + D * 4;                            // u32 [2] (+ pad)
        slot_s = (rdyn_s + 8 + 15) & ~15u;
    }
    __device__ __forceinline__ static uint32_t ld32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
    __device__ __forceinline__ static void st32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
    // all threads of the CTA; followed by a barrier of the caller
    __device__ __forceinline__ void reset(uint32_t t, uint32_t nthreads) {
        for (uint32_t d = t; d < D; d += nthreads) { st32(cnt_s + 4 * d, 0u); st32(gpos_s + 4 * d, 0u); }
        if (t < 2) st32(rdyn_s + 4 * t, 0u);
        rnd = 0;
    }
    __device__ __forceinline__ uint32_t slot_addr(uint32_t slot, uint32_t d) const { return slot_s + (slot * D + d) * RECB; }
    __device__ __forceinline__ void store(uint32_t a, const Rec<RECW>& r) const {
        if constexpr (RECW == 1) asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(r.v) : "memory");
        else asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(a), "l"(r.hi), "l"(r.lo) : "memory");
    }
    __device__ __forceinline__ Rec<RECW> load(uint32_t a) const {
        Rec<RECW> r;
        if constexpr (RECW == 1) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r.v) : "r"(a) : "memory");
        else asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.hi), "=l"(r.lo) : "r"(a) : "memory");
        return r;
    }
    // true: staged; false: the destination's slots are taken (the caller retries in the next round)
    __device__ __forceinline__ bool put(uint32_t d, const Rec<RECW>& r) const {
        uint32_t pos;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(cnt_s + 4 * d) : "memory");
        if (pos >= caps) return false;
        store(slot_addr(pos, d), r);
        if (pos == SECT - 1) {                             // fewer than SECT were left over: this happens once per round and destination
            uint32_t i;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(i) : "r"(rdyn_s + 4 * (rnd & 1u)) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(rdy_s + 2 * ((rnd & 1u) * D + i)), "h"((uint16_t)d) : "memory");
        }
        return true;
    }
    // after the round's barrier: the listed destinations' whole sectors -> their regions (region(d) = first record of
    // destination d's region, cap records long); what does not fit the region is spilled.  Fewer than SECT records stay staged.
    template <int W, typename RegionFn>
    __device__ __forceinline__ void flush_round(uint32_t t, uint32_t nthreads, RegionFn region, uint32_t cap, const PartitionPlan& plan,
                                                unsigned long long* fill, Rec<RECW>* spill, DevStatus* status) {
        const uint32_t par = rnd & 1u;
        const uint32_t n_ready = ld32(rdyn_s + 4 * par);
        if (t == 0) st32(rdyn_s + 4 * (par ^ 1u), 0u);     // the other list was consumed a barrier ago
        for (uint32_t i = t; i < n_ready; i += nthreads) {
            uint32_t d;
            { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(rdy_s + 2 * (par * D + i)) : "memory"); d = v; }
            const uint32_t n = min(ld32(cnt_s + 4 * d), caps);
            const uint32_t q = n / SECT;
            uint32_t gpos = ld32(gpos_s + 4 * d);
            Rec<RECW>* const dst = region(d);
            for (uint32_t s = 0; s < q; s++) {
                if constexpr (RECW == 1) {
                    const Rec<1> r0 = load(slot_addr(4 * s, d)), r1 = load(slot_addr(4 * s + 1, d)), r2 = load(slot_addr(4 * s + 2, d)),
                                 r3 = load(slot_addr(4 * s + 3, d));
                    if (gpos + 4 <= cap) {
                        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + gpos), "l"(r0.v), "l"(r1.v), "l"(r2.v), "l"(r3.v) : "memory");
                        gpos += 4;
                    } else {
                        spill_record<W, 1>(r0, plan, fill, spill, status); spill_record<W, 1>(r1, plan, fill, spill, status);
                        spill_record<W, 1>(r2, plan, fill, spill, status); spill_record<W, 1>(r3, plan, fill, spill, status);
                    }
                } else {
                    const Rec<2> r0 = load(slot_addr(2 * s, d)), r1 = load(slot_addr(2 * s + 1, d));
                    if (gpos + 2 <= cap) {
                        asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst + gpos), "l"(r0.hi), "l"(r0.lo), "l"(r1.hi), "l"(r1.lo) : "memory");
                        gpos += 2;
                    } else {
                        spill_record<W, 2>(r0, plan, fill, spill, status); spill_record<W, 2>(r1, plan, fill, spill, status);
                    }
                }
            }
            const uint32_t left = n - q * SECT;
            for (uint32_t j = 0; j < left; j++) store(slot_addr(j, d), load(slot_addr(q * SECT + j, d)));
            st32(cnt_s + 4 * d, left);
            st32(gpos_s + 4 * d, gpos);
        }
        rnd++;
    }
    // after the last round (and a barrier): destination d's last, partial sector; returns the records now in the region
    template <int W>
    __device__ __forceinline__ uint32_t finish(uint32_t d, Rec<RECW>* dst, uint32_t cap, const PartitionPlan& plan, unsigned long long* fill,
                                               Rec<RECW>* spill, DevStatus* status) const {
        const uint32_t n = min(ld32(cnt_s + 4 * d), caps);
        uint32_t gpos = ld32(gpos_s + 4 * d);
        for (uint32_t j = 0; j < n; j++) {
            const Rec<RECW> r = load(slot_addr(j, d));
            if (gpos < cap) dst[gpos++] = r;
            else spill_record<W, RECW>(r, plan, fill, spill, status);
        }
        return gpos;
    }
};

// ---------------------------------------------------------------------------------------------
// pass 1: column -> coarse partitions.  The run detection is partition_kernel's (count_part.cu); only the way the
// records leave the SM differs.
template <int W, int RECW>
__global__ void __launch_bounds__(NT, SCAT_MINB) scatter_kernel(ScanArgs a, PartitionPlan plan, ScatterPlan sp, uint32_t* __restrict__ segfill,
                                                                 Rec<RECW>* __restrict__ seg, unsigned long long* __restrict__ fill,
                                                                 Rec<RECW>* __restrict__ spill) {
    __shared__ ScanSmem s;
    __shared__ uint32_t wruns_h[NT / 32][SCAT_RUNCAP];     // minimizer hash of the run ...
    __shared__ uint16_t wruns_p[NT / 32][SCAT_RUNCAP];     // ... and its tile-relative start base
    __shared__ uint32_t bdm[TILE / 32 + 2];                // bit p: a run cannot continue through window p (run start or invalid window)
    extern __shared__ __align__(16) unsigned char scat_dyn[];
    TileScanner sc(a, s);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int k = a.k;
    const uint32_t rmax = plan.rmax;
    const uint32_t himask = 0xffffffffu << (32 - 2 * plan.m);
    const uint32_t D = sp.n_coarse;
    Stage<RECW> st;
    st.init(smem_u32(scat_dyn), D, sp.caps);
    st.reset(t, NT);
    if (t < 2) bdm[TILE / 32 + t] = 0xffffffffu;           // the tile end ends every run
    Rec<RECW>* const myseg = seg + (uint64_t)blockIdx.x * D * sp.seg_cap;
    uint32_t* const wh = wruns_h[warp];
    uint16_t* const wp = wruns_p[warp];
    __syncthreads();

    // the super-k-mer record of L windows starting at tile-relative base p
    auto make_record = [&](int p, int L) -> Rec<RECW> {
        const int c = p >> 4, sh = 2 * (p & 15);
        const uint32_t w0 = s.packed[c], w1 = s.packed[c + 1], w2 = s.packed[c + 2], w3 = s.packed[c + 3];
        const uint32_t r0w = __funnelshift_l(w1, w0, sh), r1w = __funnelshift_l(w2, w1, sh), r2w = __funnelshift_l(w3, w2, sh);
        const int nb = L + k - 1;                                  // bases covered
        Rec<RECW> r;
        if constexpr (RECW == 1) {
            uint64_t v = ((uint64_t)r0w << 32) | r1w;
            v &= ~0ull << (64 - 2 * nb);                           // nb <= 30
            r.v = v | (uint64_t)(L - 1);
        } else {
            uint64_t hi = ((uint64_t)r0w << 32) | r1w;
            uint64_t lo = (uint64_t)r2w << 32;                      // bases 32..47 (nb <= 47)
            if (nb <= 32) { hi &= ~0ull << (64 - 2 * nb); lo = 0; }
            else lo &= ~0ull << (128 - 2 * nb);
            r.hi = hi;
            r.lo = lo | (uint64_t)(L - 1);
        }
        return r;
    };
    auto region = [&](uint32_t d) { return myseg + (uint64_t)d * sp.seg_cap; };
    auto flush_all = [&]() { st.template flush_round<W>(t, NT, region, sp.seg_cap, plan, fill, spill, a.status); };

    while (sc.next()) {
        const uint32_t* bnd = sc.bnd();
        // this thread's 16 windows start at tile-relative bases 16t .. 16t+15 and need bases up to 16t+46;
        // window 16t-1 (the previous thread's last) is looked at as well, so that runs continue across threads
        uint32_t w[3];
        w[0] = s.packed[t]; w[1] = s.packed[t + 1]; w[2] = s.packed[t + 2];
        const uint32_t wm1 = t ? s.packed[t - 1] : 0u;
        uint32_t vmask = 0;      // bit j+1: window j is valid (j = -1 .. 15): no row start in (i, i+k-1], inside the input
        {
            const int b0 = 16 * t;                                   // bit q of bw: a row starts at base 16t + q
            const uint32_t lo = bits32(bnd, b0), hi = bits32(bnd, b0 + 32);
            const uint64_t bw = ((uint64_t)hi << 32) | lo;
            const uint64_t remaining = a.n_bases > sc.t0 + 16ull * t ? a.n_bases - (sc.t0 + 16ull * t) : 0;
            if (remaining >= 16 && ((bw >> 1) & ((1ull << (15 + k - 1)) - 1ull)) == 0) vmask = 0x1fffeu;   // the common case
            else {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    bool ok = (((uint32_t)(bw >> (j + 1)) & sc.kmask) == 0) && ((uint64_t)j < remaining);
                    vmask |= (uint32_t)ok << (j + 1);
                }
            }
            if (t && remaining && (lo & sc.kmask) == 0) vmask |= 1u;  // window 16t-1: bits 16t .. 16t+k-2
        }
        uint32_t starts = 0;
        uint32_t h[17 + W - 1];                                       // h[j+1]: minimizer hash of window j
        if (vmask >> 1) {
#pragma unroll
            for (int j = -1; j < 16 + W - 1; j++) {                   // hashed m-mers at bases -1 .. 15+W-1 of this chunk
                uint32_t top;                                         // 16 bases starting at base j
                if (j < 0) top = __funnelshift_l(w[0], wm1, 30);
                else {
                    const int q = j >> 4, sh = (j & 15) * 2;
                    top = __funnelshift_l(w[q + 1], w[q], sh);
                }
                h[j + 1] = mmer_hash(top, himask, plan.m);
            }
            // sliding minimum over W consecutive m-mers: log-steps up to the largest power of two P <= W, then two
            // overlapping P-windows cover a W-window
            constexpr int P = W >= 16 ? 16 : (W >= 8 ? 8 : (W >= 4 ? 4 : 2));
#pragma unroll
            for (int step = 1; step < P; step <<= 1) {
#pragma unroll
                for (int j = 0; j < 17 + W - 1 - step; j++) h[j] = min(h[j], h[j + step]);
            }
            if (W > P) {
#pragma unroll
                for (int j = 0; j < 17; j++) h[j] = min(h[j], h[j + W - P]);
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {                            // a run starts where the minimizer changes
                const bool v = (vmask >> (j + 1)) & 1u, pv = (vmask >> j) & 1u;
                starts |= (uint32_t)(v && (!pv || h[j + 1] != h[j])) << j;
            }
        }
        // A run is emitted by ONE lane, piece by piece: keep it short.  It may run on from the previous thread's chunk only if
        // it began there, and never across a warp edge -- so a homopolymer costs every lane one short run, not one lane thousands.
        {
            const uint32_t prev_starts = __shfl_up_sync(0xffffffffu, starts, 1);
            if (((vmask >> 1) & 1u) && (lane == 0 || prev_starts == 0)) starts |= 1u;
        }
        const uint32_t bd = (starts | ~(vmask >> 1)) & 0xffffu;
        reinterpret_cast<uint16_t*>(bdm)[t] = (uint16_t)bd;
        // ---- the warp's runs, numbered by a warp-wide exclusive scan of the run counts
        const uint32_t nrun = __popc(starts);
        uint32_t incl = nrun;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        const uint32_t n_warp_runs = __shfl_sync(0xffffffffu, incl, 31);
        __syncthreads();                                              // boundary bits of the whole tile are visible
        auto run_length = [&](uint32_t p0) {                          // the run ends before the next boundary bit after its first window
            uint32_t p = p0 + 1, R = 1;
            for (;;) {
                const uint32_t nb32 = bits32(bdm, p);
                if (nb32) { R += __ffs(nb32) - 1; break; }
                R += 32; p += 32;
            }
            return R;
        };
        Rec<RECW> pend{};                                             // a record whose destination had no free slot
        uint32_t pend_d = 0xffffffffu;
        auto put = [&](uint32_t d, const Rec<RECW>& r) {
            if (st.put(d, r)) return;
            if (pend_d == 0xffffffffu) { pend = r; pend_d = d; }
            else spill_record<W, RECW>(r, plan, fill, spill, a.status);
        };
        // ---- the warp's run list is written and consumed SCAT_RUNCAP runs at a time (one pass unless the text is adversarial)
        for (uint32_t pass0 = 0; pass0 < n_warp_runs; pass0 += SCAT_RUNCAP) {
            if (starts) {
                uint32_t rbase = incl - nrun - pass0;                 // may wrap below zero: compared unsigned
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if ((starts >> j) & 1u) {
                        if (rbase < (uint32_t)SCAT_RUNCAP) { wh[rbase] = h[j + 1]; wp[rbase] = (uint16_t)(16 * t + j); }
                        rbase++;
                    }
            }
            __syncwarp();
            const uint32_t npass = min(n_warp_runs - pass0, (uint32_t)SCAT_RUNCAP);
            for (uint32_t r = lane; r < npass; r += 32) {             // run r of the pass: emitted by lane r % 32
                const uint32_t dest = coarse_bucket(bucket_position<W>(wh[r], plan.even), plan.hash_buckets, plan.fine_shift), p0 = wp[r];
                const uint32_t R = run_length(p0);
                for (uint32_t off = 0; off < R; off += rmax) put(dest, make_record((int)(p0 + off), (int)min(R - off, rmax)));
            }
            __syncwarp();
        }
        // ---- flush: whole sectors leave; a record that found no slot gets a second round behind it
        const int any_pending = __syncthreads_or(pend_d != 0xffffffffu);
        flush_all();
        if (any_pending) {
            __syncthreads();
            if (pend_d != 0xffffffffu && !st.put(pend_d, pend)) spill_record<W, RECW>(pend, plan, fill, spill, a.status);
            __syncthreads();
            flush_all();
        }
        sc.stage ^= 1;                                                // release(): the barriers above already ordered the tile's reads
    }
    __syncthreads();
    for (uint32_t d = t; d < D; d += NT)
        segfill[(uint64_t)blockIdx.x * D + d] = st.template finish<W>(d, region(d), sp.seg_cap, plan, fill, spill, a.status);
}

// ---------------------------------------------------------------------------------------------
// pass 2: coarse partition c (its n_src segments) -> fine buckets c*F .. c*F+F-1 of the leaf's layout
// (recs [n_buckets][cap], fill[bucket] = k-mers << 32 | records in the region | poison).
template <int W, int RECW>
__global__ void __launch_bounds__(RF2_THREADS, 1) refine2_kernel(PartitionPlan plan, ScatterPlan sp, const uint32_t* __restrict__ segfill,
                                                                  const Rec<RECW>* __restrict__ seg, unsigned long long* __restrict__ fill,
                                                                  Rec<RECW>* __restrict__ recs, Rec<RECW>* __restrict__ spill, DevStatus* status) {
    extern __shared__ __align__(16) unsigned char rf2_dyn[];
    const uint32_t F = 1u << plan.fine_shift, D = sp.n_coarse;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NW = RF2_THREADS / 32;
    const uint32_t gk_s = smem_u32(rf2_dyn);                      // u32[F]: k-mers per fine bucket
    Stage<RECW> st;
    st.init(gk_s + ((F * 4 + 15) & ~15u), F, sp.caps2);
    for (uint32_t c = blockIdx.x; c < D; c += gridDim.x) {
        if ((uint32_t)t < F) asm volatile("st.shared.u32 [%0], %1;" ::"r"(gk_s + 4 * t), "r"(0u) : "memory");
        st.reset(t, RF2_THREADS);
        __syncthreads();
        auto region = [&](uint32_t f) { return recs + ((uint64_t)c * F + f) * plan.cap; };
        // warp w reads the segments w, w+NW, ... of the partition, 32*RF2_RQ records per round
        uint32_t src = warp, off = 0, n_s = 0;
        auto next_segment = [&]() {                               // skip empty segments
            for (; src < sp.n_src; src += NW) {
                n_s = min(segfill[(uint64_t)src * D + c], sp.seg_cap);
                if (n_s) break;
            }
        };
        next_segment();
        for (;;) {
            Rec<RECW> rec[RF2_RQ];
            bool have[RF2_RQ];
            if (src < sp.n_src) {
                const Rec<RECW>* base = seg + ((uint64_t)src * D + c) * sp.seg_cap;
#pragma unroll
                for (int q = 0; q < RF2_RQ; q++) {
                    const uint32_t i = off + q * 32 + lane;
                    have[q] = i < n_s;
                    if (have[q]) {
                        if constexpr (RECW == 1) rec[q].v = ld_nc_u64(reinterpret_cast<const uint64_t*>(base + i));
                        else {
                            const uint4 raw = ld_nc_u128(base + i);
                            rec[q].hi = ((uint64_t)raw.y << 32) | raw.x;
                            rec[q].lo = ((uint64_t)raw.w << 32) | raw.z;
                        }
                    }
                }
                off += 32 * RF2_RQ;
                if (off >= n_s) { src += NW; off = 0; next_segment(); }
            } else {
#pragma unroll
                for (int q = 0; q < RF2_RQ; q++) have[q] = false;
            }
            Rec<RECW> pend{};
            uint32_t pend_f = 0xffffffffu;
#pragma unroll
            for (int q = 0; q < RF2_RQ; q++) {
                if (have[q]) {
                    uint64_t hi;
                    if constexpr (RECW == 1) hi = rec[q].v; else hi = rec[q].hi;
                    const uint32_t f = record_bucket<W, RECW>(hi, plan) & (F - 1);
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(gk_s + 4 * f), "r"(rec_kmers<RECW>(rec[q])) : "memory");
                    if (!st.put(f, rec[q])) {
                        if (pend_f == 0xffffffffu) { pend = rec[q]; pend_f = f; }
                        else spill_record<W, RECW>(rec[q], plan, fill, spill, status);
                    }
                }
            }
            const int any_pending = __syncthreads_or(pend_f != 0xffffffffu);
            st.template flush_round<W>(t, RF2_THREADS, region, plan.cap, plan, fill, spill, status);
            if (any_pending) {
                __syncthreads();
                if (pend_f != 0xffffffffu && !st.put(pend_f, pend)) spill_record<W, RECW>(pend, plan, fill, spill, status);
                __syncthreads();
                st.template flush_round<W>(t, RF2_THREADS, region, plan.cap, plan, fill, spill, status);
            }
            if (!__syncthreads_or(src < sp.n_src)) break;
        }
        if ((uint32_t)t < F) {
            const uint32_t total = st.template finish<W>(t, region(t), plan.cap, plan, fill, spill, status);
            uint32_t gk;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(gk) : "r"(gk_s + 4 * t) : "memory");
            atomicOr(&fill[(uint64_t)c * F + t], ((unsigned long long)gk << 32) | total);   // keeps a poison bit set by either pass
        }
        __syncthreads();
    }
}

}  // namespace kmer
