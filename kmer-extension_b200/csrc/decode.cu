// decode.cu -- text <-> code conversion of k-mer columns.
//   decode: kmer_out (kmer.c:131-138): lower-case text, optionally with the 1-byte short varlena
//           header generate_kmers writes (SET_VARSIZE_SHORT, kmer.c:341-342).
//   encode: kmer_in (kmer.c:109-129): fold case, accept acgt only.
#include "kernels.cuh"

namespace kmer {

// one thread per output byte: fully coalesced byte stores, the 8-byte code loads hit L1/L2
__global__ void decode_kernel(const uint64_t* __restrict__ codes, uint64_t n, int k, int with_header,
                              char* __restrict__ text) {
    const uint32_t rec = (uint32_t)k + (with_header ? 1u : 0u);
    const uint64_t total = n * rec;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t o = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; o < total; o += stride) {
        uint64_t i = o / rec;
        uint32_t j = (uint32_t)(o - i * rec);
        char c;
        if (with_header) {
            if (j == 0) { text[o] = (char)(((rec) << 1) | 1u); continue; }
            j--;
        }
        uint64_t code = codes[i];
        uint32_t b = (uint32_t)(code >> (2 * (k - 1 - (int)j))) & 3u;
        c = (char)((0x74676361u >> (8 * b)) & 0xffu);   // "acgt"
        text[o] = c;
    }
}

void launch_decode(const DeviceInfo& di, const uint64_t* d_codes, uint64_t n, int k, int with_header, char* d_text,
                   cudaStream_t st) {
    uint64_t total = n * (uint64_t)(k + (with_header ? 1 : 0));
    if (!total) return;
    uint64_t blocks = (total + 255) / 256;
    uint64_t maxb = (uint64_t)di.sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    decode_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_codes, n, k, with_header, d_text);
}

__global__ void encode_kernel(const char* __restrict__ text, const uint8_t* __restrict__ lens, uint64_t n, int stride,
                              uint64_t* __restrict__ codes, DevStatus* status) {
    const uint64_t gs = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += gs) {
        int len = lens ? (int)lens[i] : stride;
        if (len > stride) len = stride;
        const unsigned char* p = reinterpret_cast<const unsigned char*>(text) + i * (uint64_t)stride;
        uint64_t v = 0;
        bool bad = false;
        for (int j = 0; j < len; j++) {
            uint32_t c = p[j] | 0x20u;
            uint32_t b = c == 'a' ? 0u : c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 4u;
            bad |= b == 4u;
            v = (v << 2) | (b & 3u);
        }
        codes[i] = v;
        if (bad) atomicMin(&status->bad_char_pos, (unsigned long long)i);   // row index for this op
    }
}

void launch_encode(const DeviceInfo& di, const char* d_text, const uint8_t* d_lens, uint64_t n, int stride,
                   uint64_t* d_codes, DevStatus* d_status, cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    uint64_t maxb = (uint64_t)di.sm_count * 16;
    if (blocks > maxb) blocks = maxb;
    encode_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_text, d_lens, n, stride, d_codes, d_status);
}

// codes[n] (uint64) -> n little-endian integers of `nbytes` bytes each (nbytes = ceil(2k/8)): the transport form of a
// column of k-mers when every byte that crosses PCIe counts.  A CTA packs 1024 codes per step into shared memory and
// writes them out as 16-byte vectors (1024 * nbytes is a multiple of 16).
__global__ void __launch_bounds__(256) pack_codes_kernel(const uint64_t* __restrict__ codes, uint64_t n, int nbytes,
                                                         uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t buf[1024 * 8];
    const uint64_t n_steps = (n + 1023) / 1024;
    for (uint64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
        const uint64_t base = step * 1024;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t j = q * 256 + threadIdx.x;
            const uint64_t i = base + j;
            uint64_t c = i < n ? codes[i] : 0ull;
            for (int b = 0; b < nbytes; b++) { buf[j * nbytes + b] = (uint8_t)c; c >>= 8; }
        }
        __syncthreads();
        const uint64_t left = n - base < 1024 ? n - base : 1024;
        const uint32_t bytes = (uint32_t)left * (uint32_t)nbytes;
        uint8_t* dst = out + base * (uint64_t)nbytes;
        for (uint32_t o = threadIdx.x * 16; o < bytes; o += 256 * 16) {
            if (o + 16 <= bytes) *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(buf + o);
            else for (uint32_t z = o; z < bytes; z++) dst[z] = buf[z];
        }
        __syncthreads();
    }
}

void launch_pack_codes(const DeviceInfo& di, const uint64_t* d_codes, uint64_t n, int nbytes, uint8_t* d_out, cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 1023) / 1024;
    uint64_t maxb = (uint64_t)di.sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    pack_codes_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_codes, n, nbytes, d_out);
}

}  // namespace kmer
