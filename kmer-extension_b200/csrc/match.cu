// match.cu -- K3/K5: batched equals / starts_with / contains over a column of k-mer codes.
//
// Reference semantics (one fmgr call per row there):
//   equals      kmer_equals               kmer.c:226-245   len equal && bytes equal
//   starts_with kmer_starts_with_helper   kmer.c:44-55     len(prefix) <= len(kmer) && leading bytes equal
//   contains    kmer_query + match()      kmer.c:59-79, kmer.h:21-53   len equal && every position admits the base
//
// contains: each IUPAC letter is a 4-bit one-hot set over (a,c,g,t); the k-mer is expanded once to
// one 4-bit one-hot nibble per base (W = ceil(k/8) 32-bit words), each pattern is held as the
// complement of its set nibbles, and a pair matches iff  onehot(kmer) & ~set(pattern) == 0  -- one
// LOP3 per word.  The warp's 32 booleans of a (pattern, 32 k-mers) block are packed with
// __ballot_sync, staged in shared memory and written as full 128-byte rows of the bit matrix.
#include "kernels.cuh"

namespace kmer {

constexpr int MT = 256;             // threads per CTA
constexpr int KPT = 8;              // k-mers per thread -> 2048 k-mers = 64 output words per CTA step (8 loads in flight per thread)
constexpr int CT = 32;              // constants per shared-memory output tile

struct KmerRegs {
    uint64_t code;
    uint32_t len;
    uint32_t oh[4];                 // one-hot nibbles, base j from the END in nibble j
};

__device__ __forceinline__ void expand_onehot(uint64_t code, uint32_t len, uint32_t oh[4]) {
#pragma unroll
    for (int w = 0; w < 4; w++) {
        uint32_t x = (uint32_t)(code >> (16 * w)) & 0xffffu;   // 8 bases, 2 bits each
        x = (x | (x << 8)) & 0x00ff00ffu;                       // spread each 2-bit field to its own nibble
        x = (x | (x << 4)) & 0x0f0f0f0fu;
        x = (x | (x << 2)) & 0x33333333u;
        uint32_t lo = x & 0x11111111u, hi = (x >> 1) & 0x11111111u;
        // nibble value v -> one-hot (1 << v)
        oh[w] = (~(hi | lo) & 0x11111111u) | ((lo & ~hi) << 1) | ((hi & ~lo) << 2) | ((hi & lo) << 3);
    }
    // clear nibbles beyond the k-mer's length
#pragma unroll
    for (int w = 0; w < 4; w++) {
        int nb = (int)len - 8 * w;
        uint32_t keep = nb >= 8 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << (4 * nb)) - 1u));
        oh[w] &= keep;
    }
}

struct ConstSmem {
    uint64_t code;
    uint32_t nset[4];   // contains: complement of the admitted-set nibbles
    uint32_t len;
    uint32_t op;
};

template <bool kHasLens, bool kOneHot>
__global__ void __launch_bounds__(MT) match_kernel(int op_default, const uint64_t* __restrict__ codes,
                                                    const uint8_t* __restrict__ lens, uint64_t m, int k,
                                                    const MatchConst* __restrict__ consts, const int* __restrict__ ops,
                                                    uint32_t n_consts, uint32_t* __restrict__ bits, uint64_t wpr,
                                                    unsigned long long* __restrict__ hits) {
    extern __shared__ unsigned char dyn[];
    ConstSmem* cs = reinterpret_cast<ConstSmem*>(dyn);                            // [n_consts]
    uint32_t* hit_acc = reinterpret_cast<uint32_t*>(cs + n_consts);               // [n_consts]
    __shared__ uint32_t tile[CT][KPT * MT / 32];                                  // [32][32]

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (uint32_t c = t; c < n_consts; c += MT) {
        MatchConst mc = consts[c];
        ConstSmem o;
        o.code = mc.code;
        o.len = mc.len;
        o.op = ops ? (uint32_t)ops[c] : (uint32_t)op_default;
        // per-base admit planes -> complement one-hot nibbles
        const uint64_t pl[4] = {mc.m0, mc.m1, mc.m2, mc.m3};
        for (int w = 0; w < 4; w++) {
            uint32_t acc = 0;
            for (int j = 0; j < 8; j++) {
                int pos = 8 * w + j;
                uint32_t nib = 0;
                for (int b = 0; b < 4; b++) nib |= (uint32_t)((pl[b] >> pos) & 1ull) << b;
                acc |= nib << (4 * j);
            }
            o.nset[w] = ~acc;
        }
        cs[c] = o;
        hit_acc[c] = 0;
    }
    __syncthreads();

    const uint64_t n_steps = (m + (uint64_t)KPT * MT - 1) / ((uint64_t)KPT * MT);
    for (uint64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
        const uint64_t base = step * (uint64_t)(KPT * MT);
        KmerRegs r[KPT];
#pragma unroll
        for (int q = 0; q < KPT; q++) {
            // warp `warp` owns output words warp*KPT + q ; lane = bit
            uint64_t i = base + (uint64_t)(warp * KPT + q) * 32 + lane;
            bool in = i < m;
            r[q].code = in ? ld_nc_u64(codes + i) : 0;
            r[q].len = in ? (kHasLens ? (uint32_t)lens[i] : (uint32_t)k) : 0xffffffffu;  // never matches
            if (kOneHot) expand_onehot(r[q].code, in ? r[q].len : 0, r[q].oh);
            else r[q].oh[0] = r[q].oh[1] = r[q].oh[2] = r[q].oh[3] = 0;
        }
        for (uint32_t c0 = 0; c0 < n_consts; c0 += CT) {
            const uint32_t nc = min((uint32_t)CT, n_consts - c0);
            for (uint32_t c = 0; c < nc; c++) {
                const ConstSmem& K = cs[c0 + c];
#pragma unroll
                for (int q = 0; q < KPT; q++) {
                    bool ok;
                    if (K.op == KMER_OP_EQUALS) {
                        ok = r[q].len == K.len && r[q].code == K.code;
                    } else if (K.op == KMER_OP_STARTS_WITH) {
                        uint32_t d = r[q].len - K.len;              // wraps when the prefix is longer
                        ok = r[q].len != 0xffffffffu && d <= 32u && ((d >= 32u ? 0ull : (r[q].code >> (2 * d))) == K.code);
                    } else if (!kOneHot) {
                        ok = false;
                    } else {
                        uint32_t bad = (r[q].oh[0] & K.nset[0]) | (r[q].oh[1] & K.nset[1]) |
                                       (r[q].oh[2] & K.nset[2]) | (r[q].oh[3] & K.nset[3]);
                        ok = bad == 0 && r[q].len == K.len;
                    }
                    uint32_t word = __ballot_sync(0xffffffffu, ok);
                    if (lane == 0) tile[c][warp * KPT + q] = word;
                }
            }
            __syncthreads();
            // rows of the bit matrix: 32 words = 128 contiguous bytes per constant
            for (uint32_t c = warp; c < nc; c += MT / 32) {
                uint32_t pc = 0;
#pragma unroll
                for (int h = 0; h < KPT * MT / 32; h += 32) {
                    uint64_t word_idx = base / 32 + h + lane;
                    uint32_t wv = tile[c][h + lane];
                    if (word_idx < wpr) bits[(uint64_t)(c0 + c) * wpr + word_idx] = wv;
                    pc += __popc(wv);
                }
                for (int d = 16; d; d >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, d);
                if (lane == 0 && pc) hit_acc[c0 + c] += pc;   // row c is always handled by the same warp
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (uint32_t c = t; c < n_consts; c += MT)
        if (hit_acc[c]) atomicAdd(&hits[c], (unsigned long long)hit_acc[c]);
}

// ---------------------------------------------------------------------------------------------
// Many constants against a column of equal-length k-mers: bit-parallel over the CONSTANTS.
//
// Testing 1000 patterns pair by pair costs ~10 instructions per (pattern, k-mer).  Every predicate here is
// "length rule AND every position admits the base" (equals: one base per position; starts_with: the prefix
// positions; contains: the IUPAC sets, match() kmer.h:21-53), so per group of 1024 constants a table
//     T[chunk][value of the chunk's 3 bases][1024 bits]  =  constants that admit these 3 bases at these positions
// is built in shared memory once, and the 1024 results of one k-mer are the AND of ceil(k/3) table rows.
// A warp takes 32 k-mers; lane l owns result bits 32l .. 32l+31 (a table row is 32 words: the row read is
// conflict free), keeps the 32 words of its 32 k-mers in registers, transposes that 32x32 bit block in
// registers, and so holds 32 finished words of the bit matrix (one per constant, 32 k-mers wide).  They are
// staged in shared memory and leave as 64-byte row segments.
constexpr int TB_THREADS = 256;
constexpr int TB_KSTEP = 512;           // k-mers per CTA step: 16 output words per constant
constexpr int TB_GROUP = 1024;          // constants per pass: 32 lanes x 32 bits
constexpr int TB_MAXCH = 11;            // ceil(32 / 3)

__device__ __forceinline__ void transpose32(uint32_t w[32]) {   // afterwards w[j] bit i == (before) w[i] bit j
    uint32_t msk = 0x0000ffffu;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
#pragma unroll
        for (int a = 0; a < 32; a++)
            if (!(a & s)) {
                const uint32_t t = ((w[a] >> s) ^ w[a + s]) & msk;
                w[a] ^= t << s;
                w[a + s] ^= t;
            }
        msk ^= msk << (s >> 1);
    }
}

constexpr int TB_COLS = TB_KSTEP / 32;  // output words per constant and CTA step
constexpr int TB_STRIDE = TB_COLS + 1;  // staging row stride in words: odd, so that 32 consecutive rows hit 32 banks

template <int NCH>
__global__ void __launch_bounds__(TB_THREADS) match_table_kernel(int op_default, const uint64_t* __restrict__ codes, uint64_t m, int k,
                                                                 const MatchConst* __restrict__ consts, const int* __restrict__ ops,
                                                                 uint32_t n_consts, uint32_t* __restrict__ bits, uint64_t wpr,
                                                                 unsigned long long* __restrict__ hits) {
    extern __shared__ __align__(16) unsigned char dyn[];
    constexpr int nch = NCH;                                                  // chunks of 3 bases: ceil(k / 3)
    uint32_t* table = reinterpret_cast<uint32_t*>(dyn);                       // [nch][64][32]
    uint32_t* tile = table + nch * 64 * 32;                                   // [1024][TB_STRIDE]
    uint32_t* hit = tile + TB_GROUP * TB_STRIDE;                              // [1024]
    uint16_t* sets = reinterpret_cast<uint16_t*>(hit + TB_GROUP);             // [nch][1024]: 3 positions x 4-bit admitted set
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint64_t n_steps = (m + TB_KSTEP - 1) / TB_KSTEP;

    for (uint32_t g0 = 0; g0 < n_consts; g0 += TB_GROUP) {
        // ---- 1. admitted sets per (constant, position); position j counts from the LAST base of the k-mer
#pragma unroll 1
        for (uint32_t c = t; c < TB_GROUP; c += TB_THREADS) {
            const uint32_t idx = g0 + c;
            MatchConst mc;
            int op = 0;
            bool live = idx < n_consts;
            if (live) {
                mc = consts[idx];
                op = ops ? ops[idx] : op_default;
                // length rules: equals / contains need equal lengths (kmer.c:240, kmer.c:64-66), a prefix must not be
                // longer than the k-mer (kmer.c:48)
                live = op == KMER_OP_STARTS_WITH ? (int)mc.len <= k : (int)mc.len == k;
            }
            const int shift = live ? k - (int)mc.len : 0;                     // starts_with: the prefix ends at position `shift`
#pragma unroll 1
            for (int ch = 0; ch < nch; ch++) {
                uint32_t s3 = 0;
#pragma unroll 1
                for (int q = 0; q < 3; q++) {
                    const int j = 3 * ch + q;
                    uint32_t set;
                    if (!live) set = 0;
                    else if (j >= k) set = 0xf;                               // beyond the k-mer: its code bits are zero
                    else if (op == KMER_OP_CONTAINS)
                        set = (uint32_t)((mc.m0 >> j) & 1) | (uint32_t)((mc.m1 >> j) & 1) << 1 | (uint32_t)((mc.m2 >> j) & 1) << 2 |
                              (uint32_t)((mc.m3 >> j) & 1) << 3;
                    else if (j < shift) set = 0xf;                            // behind the prefix
                    else set = 1u << ((mc.code >> (2 * (j - shift))) & 3);
                    s3 |= set << (4 * q);
                }
                sets[ch * TB_GROUP + c] = (uint16_t)s3;
            }
            hit[c] = 0;
        }
        __syncthreads();
        // ---- 2. the table: bit i of word l of row (chunk, v) stands for constant 32 i + l of the group
#pragma unroll 1
        for (int w = t; w < nch * 64 * 32; w += TB_THREADS) {
            const int ch = w >> 11, v = (w >> 5) & 63, l = w & 31;
            const int b0 = v & 3, b1 = (v >> 2) & 3, b2 = (v >> 4) & 3;
            uint32_t word = 0;
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const uint32_t s3 = sets[ch * TB_GROUP + 32 * i + l];
                word |= ((s3 >> b0) & (s3 >> (4 + b1)) & (s3 >> (8 + b2)) & 1u) << i;
            }
            table[w] = word;
        }
        __syncthreads();
        // ---- 3. the column
        const uint32_t n_here = min((uint32_t)TB_GROUP, n_consts - g0);
        for (uint64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
            const uint64_t base = step * TB_KSTEP;
#pragma unroll 1
            for (int r = 0; r < TB_COLS / (TB_THREADS / 32); r++) {
                const int col = warp * (TB_COLS / (TB_THREADS / 32)) + r;     // output word of this round
                const uint64_t i = base + (uint64_t)col * 32 + lane;
                const uint64_t code = i < m ? ld_nc_u64(codes + i) : 0ull;
                const uint32_t vm = __ballot_sync(0xffffffffu, i < m);
                const uint32_t clo = (uint32_t)code, chi = (uint32_t)(code >> 32);
                uint32_t w[32];
#pragma unroll
                for (int q = 0; q < 32; q++) {
                    const uint32_t qlo = __shfl_sync(0xffffffffu, clo, q);
                    const uint32_t qhi = k > 16 ? __shfl_sync(0xffffffffu, chi, q) : 0u;
                    const uint64_t cq = ((uint64_t)qhi << 32) | qlo;
                    uint32_t acc = (vm >> q) & 1u ? 0xffffffffu : 0u;
                    const uint32_t* row = table + lane;
                    uint32_t tv[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) tv[ch] = row[(ch << 11) + (((uint32_t)(cq >> (6 * ch)) & 63u) << 5)];
#pragma unroll
                    for (int ch = 0; ch < NCH; ch++) acc &= tv[ch];
                    w[q] = acc;
                }
                transpose32(w);          // w[j]: constant 32 j + lane of the group, bit i = k-mer i of this round
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    tile[(32 * j + lane) * TB_STRIDE + col] = w[j];
                    if (w[j]) atomicAdd(&hit[32 * j + lane], (uint32_t)__popc(w[j]));
                }
            }
            __syncthreads();
            // rows of the bit matrix: 16 words = 64 contiguous bytes per constant; a warp writes two rows per pass
            {
                const uint32_t cw = t & (TB_COLS - 1);
                const uint64_t word_idx = base / 32 + cw;
                if (word_idx < wpr) {
                    uint32_t* dst = bits + (uint64_t)(g0 + (t >> 4)) * wpr + word_idx;
                    const uint64_t dstep = (uint64_t)(TB_THREADS / TB_COLS) * wpr;
                    const uint32_t* src = tile + (t >> 4) * TB_STRIDE + cw;
#pragma unroll 4
                    for (uint32_t row = t >> 4; row < n_here; row += TB_THREADS / TB_COLS) {
                        *dst = *src;
                        dst += dstep;
                        src += (TB_THREADS / TB_COLS) * TB_STRIDE;
                    }
                }
            }
            __syncthreads();
        }
        for (uint32_t c = t; c < TB_GROUP; c += TB_THREADS)
            if (g0 + c < n_consts && hit[c]) atomicAdd(&hits[g0 + c], (unsigned long long)hit[c]);
        __syncthreads();
    }
}

template <bool L, bool O>
static void launch_match_t(uint64_t grid, size_t dyn, int op, const uint64_t* d_codes, const uint8_t* d_lens, uint64_t m,
                           int k, const MatchConst* d_consts, const int* d_ops, uint32_t n_consts, uint32_t* d_bits,
                           uint64_t wpr, unsigned long long* d_hits, cudaStream_t st) {
    cudaFuncSetAttribute(match_kernel<L, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    match_kernel<L, O><<<(unsigned)grid, MT, dyn, st>>>(op, d_codes, d_lens, m, k, d_consts, d_ops, n_consts, d_bits, wpr, d_hits);
}

void launch_match(const DeviceInfo& di, int op, const int* d_ops, bool any_contains, const uint64_t* d_codes,
                  const uint8_t* d_lens, uint64_t m, int k, const MatchConst* d_consts, uint32_t n_consts,
                  uint32_t* d_bits, uint64_t words_per_row, unsigned long long* d_hits, cudaStream_t st) {
    cudaMemsetAsync(d_hits, 0, (size_t)n_consts * sizeof(unsigned long long), st);
    if (!m || !n_consts) return;
    if (!d_lens && n_consts >= 96 && k >= 1) {   // many constants, one k-mer length: bit-parallel over the constants
        const int nch = (k + 2) / 3;
        const size_t dyn = (size_t)nch * 64 * 32 * 4 + (size_t)TB_GROUP * TB_STRIDE * 4 + (size_t)TB_GROUP * 4 +
                           (size_t)nch * TB_GROUP * 2;
        const uint64_t n_steps_t = (m + TB_KSTEP - 1) / TB_KSTEP;
        int per_sm = (int)((size_t)227 * 1024 / (dyn + 1024));
        if (per_sm > 2) per_sm = 2;
        uint64_t grid_t = (uint64_t)di.sm_count * per_sm;
        if (grid_t > n_steps_t) grid_t = n_steps_t;
#define KMER_TB_LAUNCH(N)                                                                                                     \
    case N:                                                                                                                   \
        cudaFuncSetAttribute(match_table_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);                   \
        match_table_kernel<N><<<(unsigned)grid_t, TB_THREADS, dyn, st>>>(op, d_codes, m, k, d_consts, d_ops, n_consts, d_bits, \
                                                                         words_per_row, d_hits);                              \
        break;
        switch (nch) {
            KMER_TB_LAUNCH(1) KMER_TB_LAUNCH(2) KMER_TB_LAUNCH(3) KMER_TB_LAUNCH(4) KMER_TB_LAUNCH(5) KMER_TB_LAUNCH(6)
            KMER_TB_LAUNCH(7) KMER_TB_LAUNCH(8) KMER_TB_LAUNCH(9) KMER_TB_LAUNCH(10) KMER_TB_LAUNCH(11)
        }
#undef KMER_TB_LAUNCH
        return;
    }
    uint64_t n_steps = (m + (uint64_t)KPT * MT - 1) / ((uint64_t)KPT * MT);
    // few constants: a pure stream of the column (8 bytes per k-mer in, bits out).  5 CTAs/SM x 8 loads per thread (48 registers)
    // keep ~80 KB of loads in flight per SM; with 4 CTAs x 4 loads the stream ran at 45 % of the copy bandwidth.
    uint64_t grid = (uint64_t)di.sm_count * (n_consts <= 8 ? 5 : 3);
    if (grid > n_steps) grid = n_steps;
    size_t dyn = (size_t)n_consts * (sizeof(ConstSmem) + sizeof(uint32_t));
    if (d_lens) {
        if (any_contains) launch_match_t<true, true>(grid, dyn, op, d_codes, d_lens, m, k, d_consts, d_ops, n_consts, d_bits, words_per_row, d_hits, st);
        else launch_match_t<true, false>(grid, dyn, op, d_codes, d_lens, m, k, d_consts, d_ops, n_consts, d_bits, words_per_row, d_hits, st);
    } else {
        if (any_contains) launch_match_t<false, true>(grid, dyn, op, d_codes, d_lens, m, k, d_consts, d_ops, n_consts, d_bits, words_per_row, d_hits, st);
        else launch_match_t<false, false>(grid, dyn, op, d_codes, d_lens, m, k, d_consts, d_ops, n_consts, d_bits, words_per_row, d_hits, st);
    }
}

}  // namespace kmer
