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
constexpr int KPT = 4;              // k-mers per thread -> 1024 k-mers = 32 output words per CTA step
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
                uint64_t word_idx = base / 32 + lane;
                uint32_t wv = tile[c][lane];
                if (word_idx < wpr) bits[(uint64_t)(c0 + c) * wpr + word_idx] = wv;
                uint32_t pc = __popc(wv);
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
    uint64_t n_steps = (m + (uint64_t)KPT * MT - 1) / ((uint64_t)KPT * MT);
    uint64_t grid = (uint64_t)di.sm_count * 4;
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
