// synth.cu -- seeded synthetic reads generated in HBM: the 10 GB benchmark shapes need no
// host-side generation and no 10 GB host-to-device copy.  Distribution restated from the reference's data_generator.py:4-11
// (characters i.i.d. uniform over "ACGT", upper case like the script's output; the reference folds case on input,
// kmer.c:28-29), made seeded and counter-based: base g of the whole table depends on (seed, g) only, so any row range can be
// produced on any device and the table does not depend on how many GPUs share it.
//   word(w) = splitmix64(seed + (w + 1) * 0x9E3779B97F4A7C15)      w = g / 32
//   base(g) = "ACGT"[(word(g / 32) >> (2 * (g % 32))) & 3]
// The numpy restatement is kmer-extension_b200/datagen.py: synth_reads_counter (tests compare the two bit for bit).
#include "kernels.cuh"

namespace kmer {

__device__ __forceinline__ uint64_t synth_word(uint64_t seed, uint64_t w) {
    uint64_t z = seed + (w + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// One thread writes 16 bases as one 16-byte store; consecutive threads write consecutive 16-byte pieces (coalesced).
// g0 = first global base of the shard (a multiple of 16 keeps a thread inside one word half; any g0 is handled).
__global__ void synth_reads_kernel(uint64_t seed, uint64_t g0, uint64_t n_bases, uint8_t* __restrict__ seq) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_vec = (n_bases + 15) / 16;          // the buffer is padded to 16 bytes (+64) by the caller
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n_vec; v += stride) {
        const uint64_t g = g0 + v * 16;
        uint32_t out[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t bytes = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t gi = g + q * 4 + j;
                const uint32_t b = (uint32_t)(synth_word(seed, gi >> 5) >> (2 * (gi & 31))) & 3u;
                bytes |= ((0x54474341u >> (8 * b)) & 0xffu) << (8 * j);   // 'A' 'C' 'G' 'T'
            }
            out[q] = bytes;
        }
        if (v * 16 + 16 <= n_bases)
            *reinterpret_cast<uint4*>(seq + v * 16) = make_uint4(out[0], out[1], out[2], out[3]);
        else
            for (uint64_t i = v * 16; i < n_bases; i++) seq[i] = (uint8_t)(out[(i & 15) >> 2] >> (8 * (i & 3)));
    }
}

__global__ void synth_offsets_kernel(uint64_t n_rows, uint64_t read_len, uint64_t* __restrict__ off) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r <= n_rows; r += stride) off[r] = r * read_len;
}

void launch_synth_reads(const DeviceInfo& di, uint64_t seed, uint64_t first_row, uint64_t n_rows, uint64_t read_len, char* d_seq,
                        uint64_t* d_row_off, cudaStream_t st) {
    const uint64_t n_bases = n_rows * read_len;
    const unsigned grid = (unsigned)di.sm_count * 8;
    if (n_bases) synth_reads_kernel<<<grid, 256, 0, st>>>(seed, first_row * read_len, n_bases, reinterpret_cast<uint8_t*>(d_seq));
    synth_offsets_kernel<<<grid, 256, 0, st>>>(n_rows, read_len, d_row_off);
}

}  // namespace kmer
