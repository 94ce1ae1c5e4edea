// api.cu -- the C ABI of libkmer_cuda.so (include/kmer_cuda.h): context, workspaces, error mapping,
// host-buffer batch submit and the device-resident entry points.  No CPU compute path exists here:
// every operation is a sequence of CUDA kernels; without a device the calls fail.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace kmer;

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct PinnedBuf {
    void* p;
    size_t cap;
    bool in_use;
};

enum PendingOp { OP_NONE = 0, OP_EXTRACT, OP_COUNT, OP_MATCH, OP_DECODE, OP_ENCODE, OP_SHARD_PART, OP_SHARD_COUNT, OP_DENSE_TABLE };

}  // namespace

struct kmer_cuda_ctx {
    DeviceInfo di{};
    cudaStream_t stream = nullptr;
    kmer_cuda_error err{};
    uint64_t launches = 0;
    DevStatus* d_status = nullptr;
    DevStatus* h_status = nullptr;  // pinned
    // device workspaces, grown on demand and kept between calls
    Buf seq, off, mask, tile_row, table, consts, ops, codes, pairs, bits, hits, lens, text, fill, recs, spill, failed, seg, segfill;
    uint64_t last_tier2 = 0;      // k-mers counted by the tier-2 kernel in the last count
    uint64_t last_overflow = 0;   // k-mers the partition counter could not place (batch was recounted)
    std::vector<PinnedBuf> pinned;
    // the operation kmer_cuda_dev_finish() has to report on
    PendingOp pending = OP_NONE;
    uint64_t p_n_bases = 0, p_n_rows = 0;
    int p_k = 0;
    uint64_t p_expected_kmers = 0;
    // optional phase timing (bench.py's per-kernel roofline): events recorded after each phase
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<const char*> ev_names;   // ev_names[i] = phase that ENDS at event i ("" for the start mark)
    size_t ev_used = 0;
    std::vector<float> phase_ms;
    std::vector<const char*> phase_names;
};

static void mark(kmer_cuda_ctx* c, cudaStream_t st, const char* name) {
    if (!c->profiling) return;
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->ev_pool.push_back(e);
        c->ev_names.push_back("");
    }
    c->ev_names[c->ev_used] = name;
    cudaEventRecord(c->ev_pool[c->ev_used++], st);
}

static kmer_cuda_error g_init_error;  // error of a failed kmer_cuda_init (no ctx to hold it)

// ------------------------------------------------------------------------------------------------
// errors

static int set_error(kmer_cuda_error* e, int status, const char* sqlstate, const char* msg, const char* detail,
                     int64_t row) {
    e->status = status;
    snprintf(e->sqlstate, sizeof(e->sqlstate), "%s", sqlstate);
    snprintf(e->message, sizeof(e->message), "%s", msg);
    snprintf(e->detail, sizeof(e->detail), "%s", detail ? detail : "");
    e->row = row;
    return status;
}

// the reference's own errors: same SQLSTATE, same text (kmer.c:33-36,117-119,151-153,179-181,311-313)
static int ref_error(kmer_cuda_error* e, int status, int64_t row) {
    switch (status) {
        case KMER_ERR_INVALID_DNA:
            return set_error(e, status, "22P02", "Invalid DNA Sequence",
                             "Valid characters are A, C, G, T (case-insensitive).", row);
        case KMER_ERR_KMER_TOO_LONG:
            return set_error(e, status, "22001", "KMer Sequence larger than length 32", "", row);
        case KMER_ERR_INVALID_QKMER:
            return set_error(e, status, "22P02", "Invalid QKMer Sequence", "", row);
        case KMER_ERR_INVALID_K:
            return set_error(e, status, "22023", "Invalid KMER Length", "", row);
        case KMER_ERR_QKMER_TOO_LONG:
            return set_error(e, status, "22001", "QKMer Sequence larger than length 32", "", row);
    }
    return set_error(e, status, "XX000", "kmer_cuda: internal error", "", row);
}

static int cuda_error(kmer_cuda_ctx* c, cudaError_t ce, const char* what) {
    char msg[160];
    snprintf(msg, sizeof(msg), "kmer_cuda: %s failed: %s", what, cudaGetErrorString(ce));
    bool oom = ce == cudaErrorMemoryAllocation;
    return set_error(c ? &c->err : &g_init_error, oom ? KMER_ERR_OOM : KMER_ERR_CUDA, oom ? "53200" : "XX000", msg, "", -1);
}

static int bad_arg(kmer_cuda_ctx* c, const char* what) {
    char msg[160];
    snprintf(msg, sizeof(msg), "kmer_cuda: bad argument: %s", what);
    return set_error(&c->err, KMER_ERR_BAD_ARGUMENT, "XX000", msg, "", -1);
}

#define CU(call, what)                                      \
    do {                                                    \
        cudaError_t ce__ = (call);                          \
        if (ce__ != cudaSuccess) return cuda_error(c, ce__, what); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// memory

static int ws(kmer_cuda_ctx* c, Buf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return KMER_OK;
    if (b.p) {
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t ce = cudaMalloc(&b.p, want);
    if (ce != cudaSuccess) {
        b.p = nullptr;
        cudaGetLastError();
        return cuda_error(c, ce, "cudaMalloc");
    }
    b.cap = want;
    return KMER_OK;
}

static void* pinned_get(kmer_cuda_ctx* c, size_t bytes) {
    if (bytes == 0) bytes = 16;
    int best = -1;
    for (size_t i = 0; i < c->pinned.size(); i++)
        if (!c->pinned[i].in_use && c->pinned[i].cap >= bytes && (best < 0 || c->pinned[i].cap < c->pinned[best].cap))
            best = (int)i;
    if (best >= 0) {
        c->pinned[best].in_use = true;
        return c->pinned[best].p;
    }
    // drop idle buffers that are too small before growing (keeps the pinned footprint bounded)
    for (size_t i = 0; i < c->pinned.size();) {
        if (!c->pinned[i].in_use) {
            cudaFreeHost(c->pinned[i].p);
            c->pinned.erase(c->pinned.begin() + i);
        } else
            i++;
    }
    void* p = nullptr;
    cudaError_t ce = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        cuda_error(c, ce, "cudaHostAlloc");
        return nullptr;
    }
    c->pinned.push_back({p, bytes, true});
    return p;
}

static bool is_device_accessible_host(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// ------------------------------------------------------------------------------------------------
// small kernels of the API layer

__global__ void status_reset_kernel(DevStatus* s) {
    s->bad_char_pos = kNoError;
    s->short_row = kNoError;
    s->n_kmers = 0;
    s->n_distinct = 0;
    s->n_overflow = 0;
    s->special_count = 0;
    s->out_overflow = 0;
    s->pad = kNoError;
    s->n_spill = 0;
    s->n_failed = 0;
    s->failed_kmers = 0;
    s->n_unique = 0;
}

// pad := row containing bad_char_pos (so the host never needs the offsets)
__global__ void resolve_bad_row_kernel(DevStatus* s, const uint64_t* off, uint64_t n_rows) {
    unsigned long long pos = s->bad_char_pos;
    if (pos == kNoError) return;
    uint64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    s->pad = lo - 1;
}

// ------------------------------------------------------------------------------------------------
// constants: kmer / qkmer literals -> MatchConst   (kmer_in kmer.c:109-129, qkmer_in kmer.c:141-190, match() kmer.h:21-53)

static int base_of(int ch) {
    switch (ch | 0x20) {
        case 'a': return 0;
        case 'c': return 1;
        case 'g': return 2;
        case 't': return 3;
    }
    return -1;
}

static int iupac_set(int ch) {  // bit b set <=> base b admitted ; -1 invalid letter
    if (!((ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z'))) return -1;
    switch (ch | 0x20) {
        case 'a': return 1;
        case 'c': return 2;
        case 'g': return 4;
        case 't': return 8;
        case 'u': return 0;  // accepted by qkmer_in (kmer.c:165) but matches nothing (kmer.h:50-51)
        case 'r': return 1 | 4;
        case 'y': return 2 | 8;
        case 'k': return 4 | 8;
        case 'm': return 1 | 2;
        case 's': return 4 | 2;
        case 'w': return 1 | 8;
        case 'b': return 2 | 4 | 8;
        case 'd': return 1 | 4 | 8;
        case 'h': return 1 | 2 | 8;
        case 'v': return 1 | 2 | 4;
        case 'n': return 15;
    }
    return -1;
}

static int compile_const(kmer_cuda_ctx* c, int op, const char* text, int64_t idx, MatchConst* out) {
    memset(out, 0, sizeof(*out));
    if (!text) return bad_arg(c, "NULL constant (the SQL functions are STRICT: filter NULLs in the caller)");
    size_t len = strlen(text);
    if (op == KMER_OP_CONTAINS) {
        if (len > KMER_CUDA_MAX_K) return ref_error(&c->err, KMER_ERR_QKMER_TOO_LONG, idx);
        uint64_t pl[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < len; i++) {
            int s = iupac_set((unsigned char)text[i]);
            if (s < 0) return ref_error(&c->err, KMER_ERR_INVALID_QKMER, idx);
            size_t j = len - 1 - i;
            for (int b = 0; b < 4; b++)
                if (s & (1 << b)) pl[b] |= 1ull << j;
        }
        out->m0 = pl[0]; out->m1 = pl[1]; out->m2 = pl[2]; out->m3 = pl[3];
    } else if (op == KMER_OP_EQUALS || op == KMER_OP_STARTS_WITH) {
        if (len > KMER_CUDA_MAX_K) return ref_error(&c->err, KMER_ERR_KMER_TOO_LONG, idx);
        uint64_t v = 0;
        for (size_t i = 0; i < len; i++) {
            int b = base_of((unsigned char)text[i]);
            if (b < 0) return ref_error(&c->err, KMER_ERR_INVALID_DNA, idx);
            v = (v << 2) | (uint64_t)b;
        }
        out->code = v;
    } else
        return bad_arg(c, "unknown match op");
    out->len = (uint32_t)len;
    return KMER_OK;
}

// ------------------------------------------------------------------------------------------------
// lifecycle

extern "C" int kmer_cuda_abi_version(void) { return KMER_CUDA_ABI_VERSION; }

extern "C" int kmer_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int kmer_cuda_init(kmer_cuda_ctx** out, int device) {
    kmer_cuda_ctx* c = nullptr;
    if (!out) return set_error(&g_init_error, KMER_ERR_BAD_ARGUMENT, "XX000", "kmer_cuda: bad argument: ctx", "", -1);
    *out = nullptr;
    int n = kmer_cuda_device_count();
    if (n <= 0)
        return set_error(&g_init_error, KMER_ERR_NO_DEVICE, "XX000",
                         "kmer_cuda: no CUDA device available (this library has no CPU path)", "", -1);
    if (device < 0 || device >= n)
        return set_error(&g_init_error, KMER_ERR_BAD_ARGUMENT, "XX000", "kmer_cuda: bad argument: device index", "", -1);
    CU(cudaSetDevice(device), "cudaSetDevice");
    c = new (std::nothrow) kmer_cuda_ctx();
    if (!c) return set_error(&g_init_error, KMER_ERR_OOM, "53200", "kmer_cuda: out of host memory", "", -1);
    cudaDeviceProp prop;
    cudaError_t ce = cudaGetDeviceProperties(&prop, device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&c->d_status, sizeof(DevStatus));
    if (ce == cudaSuccess) ce = cudaHostAlloc((void**)&c->h_status, sizeof(DevStatus), cudaHostAllocDefault);
    if (ce != cudaSuccess) {
        int rc = cuda_error(nullptr, ce, "context setup");
        delete c;
        return rc;
    }
    c->di.device = device;
    c->di.sm_count = prop.multiProcessorCount;
    c->di.total_mem = prop.totalGlobalMem;
    if (prop.major < 10) {
        delete c;
        return set_error(&g_init_error, KMER_ERR_NO_DEVICE, "XX000", "kmer_cuda: built for sm_100a (B200) only", "", -1);
    }
    c->err.status = KMER_OK;
    c->err.row = -1;
    *out = c;
    return KMER_OK;
}

static void buf_free(Buf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" void kmer_cuda_shutdown(kmer_cuda_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->di.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    Buf* all[] = {&c->seq, &c->off, &c->mask, &c->tile_row, &c->table, &c->consts, &c->ops,
                  &c->codes, &c->pairs, &c->bits, &c->hits, &c->lens, &c->text, &c->fill, &c->recs, &c->spill, &c->failed, &c->seg, &c->segfill};
    for (Buf* b : all) buf_free(*b);
    for (auto& p : c->pinned) cudaFreeHost(p.p);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->d_status) cudaFree(c->d_status);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" const kmer_cuda_error* kmer_cuda_last_error(const kmer_cuda_ctx* c) { return c ? &c->err : &g_init_error; }

extern "C" void kmer_cuda_release(kmer_cuda_ctx* c, void* result) {
    if (!c || !result) return;
    for (auto& p : c->pinned)
        if (p.p == result) p.in_use = false;  // kept for reuse; freed at shutdown
}

extern "C" void kmer_cuda_set_profiling(kmer_cuda_ctx* c, int on) {
    if (c) c->profiling = on != 0;
}

extern "C" int kmer_cuda_get_phases(const kmer_cuda_ctx* c, const char** names, float* ms, int capacity) {
    if (!c) return 0;
    int n = (int)c->phase_ms.size();
    for (int i = 0; i < n && i < capacity; i++) {
        if (names) names[i] = c->phase_names[i];
        if (ms) ms[i] = c->phase_ms[i];
    }
    return n;
}

extern "C" uint64_t kmer_cuda_launch_count(const kmer_cuda_ctx* c) { return c ? c->launches : 0; }

extern "C" uint64_t kmer_cuda_max_kmers(uint64_t n_bases, uint64_t n_rows, int k) {
    if (k < 1 || k > KMER_CUDA_MAX_K) return 0;
    uint64_t sub = n_rows * (uint64_t)(k - 1);
    return n_bases > sub ? n_bases - sub : 0;
}

// ------------------------------------------------------------------------------------------------
// device-resident operations

// stream argument of the C ABI: NULL = the CUDA default stream (what a caller that never created a
// stream is using); KMER_OWN_STREAM = the context's private stream (used by the submit_* calls).
#define KMER_OWN_STREAM ((void*)(uintptr_t)1)
static cudaStream_t pick_stream(kmer_cuda_ctx* c, void* stream) {
    return stream == KMER_OWN_STREAM ? c->stream : (cudaStream_t)stream;
}

static int begin_op(kmer_cuda_ctx* c, cudaStream_t st) {
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    c->ev_used = 0;
    mark(c, st, "");
    status_reset_kernel<<<1, 1, 0, st>>>(c->d_status);
    c->launches++;
    c->pending = OP_NONE;
    return KMER_OK;
}

// row mask + short-row check shared by extract and count.  Returns KMER_ERR_INVALID_K immediately for
// k outside 1..32 (generate_kmers, kmer.c:310: `window_size <= 0 || window_size > MAX_KMER_LENGTH`).
static int prepare_rows(kmer_cuda_ctx* c, const uint64_t* d_off, uint64_t n_bases, uint64_t n_rows, int k,
                        cudaStream_t st, ScanArgs* a, const char* d_seq) {
    if (n_rows && (k < 1 || k > KMER_CUDA_MAX_K)) return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
    if ((reinterpret_cast<uintptr_t>(d_seq) & 15) != 0) return bad_arg(c, "d_seq must be 16-byte aligned");
    uint64_t mask_words = (n_bases + 1 + 31) / 32 + MASK_PAD_WORDS;
    int rc = ws(c, c->mask, mask_words * 4);
    if (rc) return rc;
    launch_rows_prepare(d_off, n_rows, n_bases, k, (uint32_t*)c->mask.p, mask_words, c->d_status, st);
    c->launches++;
    mark(c, st, "rows_prepare");
    a->seq = reinterpret_cast<const uint8_t*>(d_seq);
    a->n_bases = n_bases;
    a->row_mask = (const uint32_t*)c->mask.p;
    a->k = k;
    a->status = c->d_status;
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_extract(kmer_cuda_ctx* c, const char* d_seq, uint64_t n_bases, const uint64_t* d_row_off,
                                     uint64_t n_rows, int k, uint64_t* d_codes, uint64_t codes_capacity, void* stream) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_EXTRACT;
    c->p_n_bases = n_bases; c->p_n_rows = n_rows; c->p_k = k;
    c->p_expected_kmers = kmer_cuda_max_kmers(n_bases, n_rows, k);
    if (n_rows == 0 || n_bases == 0) {
        if (n_rows && (k < 1 || k > KMER_CUDA_MAX_K)) return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
        if (n_rows) {  // rows exist but are all empty: len 0 < k
            return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
        }
        return KMER_OK;
    }
    ScanArgs a;
    rc = prepare_rows(c, d_row_off, n_bases, n_rows, k, st, &a, d_seq);
    if (rc) return rc;
    uint64_t n_tiles = (n_bases + TILE - 1) / TILE;
    rc = ws(c, c->tile_row, n_tiles * 4);
    if (rc) return rc;
    launch_tile_row_base(d_row_off, n_rows, n_tiles, (uint32_t*)c->tile_row.p, st);
    mark(c, st, "tile_row_base");
    launch_extract(c->di, a, (const uint32_t*)c->tile_row.p, d_codes, codes_capacity, st);
    mark(c, st, "extract");
    resolve_bad_row_kernel<<<1, 1, 0, st>>>(c->d_status, d_row_off, n_rows);
    c->launches += 3;
    CU(cudaGetLastError(), "extract launch");
    return KMER_OK;
}

struct MarkArg {
    kmer_cuda_ctx* c;
    cudaStream_t st;
};
static void mark_cb(void* arg, const char* name) {
    MarkArg* m = (MarkArg*)arg;
    mark(m->c, m->st, name);
}

static uint64_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// d_uniq != nullptr: split result format (see kmer_cuda_dev_count_split)
static int dev_count_impl(kmer_cuda_ctx* c, const char* d_seq, uint64_t n_bases, const uint64_t* d_row_off, uint64_t n_rows, int k,
                          kmer_count_pair* d_pairs, uint64_t pairs_capacity, uint64_t* d_uniq, uint64_t uniq_capacity, int algo,
                          void* stream) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_COUNT;
    c->p_n_bases = n_bases; c->p_n_rows = n_rows; c->p_k = k;
    c->p_expected_kmers = kmer_cuda_max_kmers(n_bases, n_rows, k);
    if (n_rows == 0 || n_bases == 0) {
        if (n_rows) return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
        return KMER_OK;
    }
    ScanArgs a;
    rc = prepare_rows(c, d_row_off, n_bases, n_rows, k, st, &a, d_seq);
    if (rc) return rc;
    if (algo == 0) algo = (k <= 13) ? 1 : 3;
    c->last_overflow = 0;
    c->last_tier2 = 0;
    if (algo == 3) {
        if (k < 14) return bad_arg(c, "minimizer-partition counting needs k >= 14");
        PartitionPlan plan = make_partition_plan(c->p_expected_kmers, k);
        ScatterPlan sp{};
        const bool two_pass = make_scatter_plan(c->di, n_bases, c->p_expected_kmers, plan, sp);
        rc = ws(c, c->fill, (size_t)plan.n_buckets * 8);
        if (!rc) rc = ws(c, c->recs, partition_record_bytes(plan));
        if (!rc) rc = ws(c, c->spill, partition_spill_bytes(plan));
        if (!rc) rc = ws(c, c->failed, (size_t)plan.n_buckets * 4);
        if (!rc && two_pass) rc = ws(c, c->seg, scatter_seg_bytes(plan, sp));
        if (!rc && two_pass) rc = ws(c, c->segfill, scatter_segfill_bytes(sp));
        if (rc) return rc;
        MarkArg ma{c, st};
        if (two_pass) {
            launch_scatter_refine(c->di, a, plan, sp, (uint32_t*)c->segfill.p, c->seg.p, (unsigned long long*)c->fill.p, c->recs.p,
                                  c->spill.p, st, mark_cb, &ma);
            launch_bucket_count(c->di, plan, k, (const unsigned long long*)c->fill.p, c->recs.p, c->spill.p, (uint32_t*)c->failed.p, d_pairs,
                                pairs_capacity, d_uniq, uniq_capacity, c->d_status, st);
            mark(c, st, "bucket_count");
            c->launches += 3;
        } else {
            launch_count_partition(c->di, a, plan, (unsigned long long*)c->fill.p, c->recs.p, c->spill.p, (uint32_t*)c->failed.p,
                                   d_pairs, pairs_capacity, d_uniq, uniq_capacity, st, mark_cb, &ma);
            c->launches += 2;
        }
        // Did everything fit?  (One host round trip; skewed input needs tier 2 or a full recount.)
        CU(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, st), "D2H status");
        CU(cudaStreamSynchronize(st), "stream sync");
        const DevStatus hs = *c->h_status;
        if (hs.bad_char_pos != kNoError || hs.short_row != kNoError) {
            algo = -1;  // failing with an input error that finish() reports
        } else if (hs.n_overflow != 0) {
            c->last_overflow = hs.n_overflow;        // tier 3: recount the batch through the global hash table
            status_reset_kernel<<<1, 1, 0, st>>>(c->d_status);
            c->launches++;
            algo = 2;
        } else {
            if (hs.n_failed || hs.n_spill) {         // tier 2: only the buckets that did not fit
                uint64_t n_slots = next_pow2(std::max<uint64_t>(1024, (hs.failed_kmers + hs.n_spill * 16) * 2));
                rc = ws(c, c->table, n_slots * sizeof(kmer_count_pair));
                if (rc) return rc;
                launch_hash_clear((kmer_count_pair*)c->table.p, n_slots, st);
                launch_partition_tier2(c->di, plan, k, 1, (const unsigned long long*)c->fill.p, c->recs.p, c->spill.p,
                                       (const uint32_t*)c->failed.p, (kmer_count_pair*)c->table.p, n_slots, c->d_status, st);
                mark(c, st, "tier2_insert");
                // compaction appends the table and the k==32 special key and adds both to n_kmers
                launch_hash_compact(c->di, (const kmer_count_pair*)c->table.p, n_slots, k, d_pairs, pairs_capacity, c->d_status, st);
                mark(c, st, "tier2_compact");
                c->launches += 2;
                c->last_tier2 = hs.failed_kmers;
            } else {
                launch_append_special(d_pairs, pairs_capacity, c->d_status, st);
                c->launches++;
            }
            algo = -1;
        }
    }
    if (algo == 1) {
        if (k > 15) return bad_arg(c, "dense counting needs k <= 15");
        uint64_t nbins = 1ull << (2 * k);
        rc = ws(c, c->table, nbins * 8);
        if (rc) return rc;
        launch_count_dense(c->di, a, (unsigned long long*)c->table.p, d_pairs, pairs_capacity, st);
        c->launches += 2;
        mark(c, st, "count_dense+compact");
    } else if (algo == 2) {
        uint64_t maxd = c->p_expected_kmers;
        if (k < 32 && (1ull << (2 * k)) < maxd) maxd = 1ull << (2 * k);
        uint64_t n_slots = next_pow2(std::max<uint64_t>(1024, maxd * 2));
        rc = ws(c, c->table, n_slots * sizeof(kmer_count_pair));
        if (rc) return rc;
        launch_hash_clear((kmer_count_pair*)c->table.p, n_slots, st);
        mark(c, st, "hash_clear");
        launch_count_hash_insert(c->di, a, (kmer_count_pair*)c->table.p, n_slots, st);
        mark(c, st, "count_hash_insert");
        launch_hash_compact(c->di, (const kmer_count_pair*)c->table.p, n_slots, k, d_pairs, pairs_capacity, c->d_status, st);
        c->launches += 2;
        mark(c, st, "hash_compact");
    } else if (algo != -1)
        return bad_arg(c, "unknown counting algorithm");
    resolve_bad_row_kernel<<<1, 1, 0, st>>>(c->d_status, d_row_off, n_rows);
    c->launches++;
    CU(cudaGetLastError(), "count launch");
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_count(kmer_cuda_ctx* c, const char* d_seq, uint64_t n_bases, const uint64_t* d_row_off,
                                   uint64_t n_rows, int k, kmer_count_pair* d_pairs, uint64_t pairs_capacity, int algo,
                                   void* stream) {
    return dev_count_impl(c, d_seq, n_bases, d_row_off, n_rows, k, d_pairs, pairs_capacity, nullptr, 0, algo, stream);
}

extern "C" int kmer_cuda_dev_count_split(kmer_cuda_ctx* c, const char* d_seq, uint64_t n_bases, const uint64_t* d_row_off,
                                         uint64_t n_rows, int k, uint64_t* d_uniq, uint64_t uniq_capacity,
                                         kmer_count_pair* d_pairs, uint64_t pairs_capacity, void* stream) {
    if (c && !d_uniq && uniq_capacity) return bad_arg(c, "d_uniq");
    return dev_count_impl(c, d_seq, n_bases, d_row_off, n_rows, k, d_pairs, pairs_capacity, d_uniq, d_uniq ? uniq_capacity : 0, 0,
                          stream);
}

static int upload_consts(kmer_cuda_ctx* c, int op, const int* ops, const char* const* consts, uint32_t n_consts,
                         cudaStream_t st, bool* any_contains, const int** d_ops_out) {
    std::vector<MatchConst> mc(n_consts ? n_consts : 1);
    *any_contains = false;
    for (uint32_t i = 0; i < n_consts; i++) {
        int o = ops ? ops[i] : op;
        int rc = compile_const(c, o, consts[i], (int64_t)i, &mc[i]);
        if (rc) return rc;
        if (o == KMER_OP_CONTAINS) *any_contains = true;
    }
    int rc = ws(c, c->consts, sizeof(MatchConst) * (size_t)n_consts);
    if (rc) return rc;
    // staged through pinned memory so the async copy really is asynchronous and the vector may die
    size_t bytes = sizeof(MatchConst) * (size_t)n_consts + (ops ? sizeof(int) * (size_t)n_consts : 0);
    char* stage = (char*)pinned_get(c, bytes);
    if (!stage) return c->err.status;
    memcpy(stage, mc.data(), sizeof(MatchConst) * (size_t)n_consts);
    CU(cudaMemcpyAsync(c->consts.p, stage, sizeof(MatchConst) * (size_t)n_consts, cudaMemcpyHostToDevice, st), "H2D consts");
    *d_ops_out = nullptr;
    if (ops) {
        rc = ws(c, c->ops, sizeof(int) * (size_t)n_consts);
        if (rc) return rc;
        memcpy(stage + sizeof(MatchConst) * (size_t)n_consts, ops, sizeof(int) * (size_t)n_consts);
        CU(cudaMemcpyAsync(c->ops.p, stage + sizeof(MatchConst) * (size_t)n_consts, sizeof(int) * (size_t)n_consts,
                           cudaMemcpyHostToDevice, st), "H2D ops");
        *d_ops_out = (const int*)c->ops.p;
    }
    CU(cudaStreamSynchronize(st), "consts sync");
    kmer_cuda_release(c, stage);
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_match(kmer_cuda_ctx* c, int op, const int* ops, const uint64_t* d_codes, const uint8_t* d_lens,
                                   uint64_t m, int k, const char* const* consts, uint32_t n_consts, uint32_t* d_bits,
                                   uint64_t* d_hits, void* stream) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_MATCH;
    if (!d_lens && (k < 0 || k > KMER_CUDA_MAX_K)) return bad_arg(c, "k-mer length must be 0..32");
    if (n_consts > 4096) return bad_arg(c, "at most 4096 constants per call");
    bool any_contains = false;
    const int* d_ops = nullptr;
    rc = upload_consts(c, op, ops, consts, n_consts, st, &any_contains, &d_ops);
    if (rc) return rc;
    uint64_t wpr = (m + 31) / 32;
    launch_match(c->di, op, d_ops, any_contains, d_codes, d_lens, m, k, (const MatchConst*)c->consts.p, n_consts, d_bits,
                 wpr, (unsigned long long*)d_hits, st);
    c->launches++;
    mark(c, st, "match");
    CU(cudaGetLastError(), "match launch");
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_decode(kmer_cuda_ctx* c, const uint64_t* d_codes, uint64_t n, int k, int with_header,
                                    char* d_text, void* stream) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_DECODE;
    if (k < 0 || k > KMER_CUDA_MAX_K) return bad_arg(c, "k-mer length must be 0..32");
    launch_decode(c->di, d_codes, n, k, with_header, d_text, st);
    c->launches++;
    CU(cudaGetLastError(), "decode launch");
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_finish(kmer_cuda_ctx* c, void* stream, kmer_dev_result* result) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    CU(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, st), "D2H status");
    CU(cudaStreamSynchronize(st), "stream sync");
    c->phase_ms.clear();
    c->phase_names.clear();
    for (size_t i = 1; i < c->ev_used; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_pool[i - 1], c->ev_pool[i]) != cudaSuccess) cudaGetLastError();
        c->phase_ms.push_back(ms);
        c->phase_names.push_back(c->ev_names[i]);
    }
    c->ev_used = 0;
    const DevStatus& s = *c->h_status;
    if (result) {
        result->n_kmers = 0;
        result->n_distinct = 0;
        result->n_overflow = 0;
        result->n_tier2 = 0;
    }
    PendingOp op = c->pending;
    c->pending = OP_NONE;
    if (op == OP_EXTRACT || op == OP_COUNT) {
        // the first offending row decides; on the same row dna_in (text -> dna) precedes generate_kmers
        uint64_t bad_row = s.bad_char_pos == kNoError ? kNoError : s.pad;
        if (bad_row != kNoError && bad_row <= s.short_row) return ref_error(&c->err, KMER_ERR_INVALID_DNA, (int64_t)bad_row);
        if (s.short_row != kNoError) return ref_error(&c->err, KMER_ERR_INVALID_K, (int64_t)s.short_row);
        if (s.out_overflow)
            return set_error(&c->err, KMER_ERR_CAPACITY, "XX000", "kmer_cuda: output buffer too small", "", -1);
        if (result) {
            result->n_kmers = op == OP_EXTRACT ? c->p_expected_kmers : s.n_kmers;
            result->n_distinct = s.n_distinct;
            result->n_overflow = c->last_overflow;
            result->n_tier2 = c->last_tier2;
            result->n_unique = s.n_unique;
        }
        if (op == OP_COUNT && s.n_kmers != c->p_expected_kmers)
            return set_error(&c->err, KMER_ERR_CUDA, "XX000", "kmer_cuda: internal error: counted k-mers != windows", "", -1);
    } else if (op == OP_SHARD_PART || op == OP_DENSE_TABLE) {
        uint64_t bad_row = s.bad_char_pos == kNoError ? kNoError : s.pad;
        if (bad_row != kNoError && bad_row <= s.short_row) return ref_error(&c->err, KMER_ERR_INVALID_DNA, (int64_t)bad_row);
        if (s.short_row != kNoError) return ref_error(&c->err, KMER_ERR_INVALID_K, (int64_t)s.short_row);
        if (s.n_overflow)
            return set_error(&c->err, KMER_ERR_CAPACITY, "XX000",
                             "kmer_cuda: a bucket segment overflowed (input too repetitive for the sharded partition path)", "", -1);
        if (result) result->n_kmers = c->p_expected_kmers;
    } else if (op == OP_SHARD_COUNT) {
        if (s.out_overflow)
            return set_error(&c->err, KMER_ERR_CAPACITY, "XX000", "kmer_cuda: output buffer too small", "", -1);
        if (c->last_overflow)
            return set_error(&c->err, KMER_ERR_CAPACITY, "XX000",
                             "kmer_cuda: the spill list overflowed (input too repetitive for the sharded partition path)", "", -1);
        if (result) {
            result->n_kmers = s.n_kmers;
            result->n_distinct = s.n_distinct;
            result->n_tier2 = c->last_tier2;
            result->n_unique = s.n_unique;
        }
    } else if (op == OP_ENCODE) {
        if (s.bad_char_pos != kNoError) return ref_error(&c->err, KMER_ERR_INVALID_DNA, (int64_t)s.bad_char_pos);
    }
    return KMER_OK;
}

// ------------------------------------------------------------------------------------------------
// sharded counting (the caller owns the exchange)

extern "C" int kmer_cuda_shard_plan(uint64_t total_kmers, int k, uint32_t n_ranks, kmer_shard_plan* plan) {
    return kmer_cuda_shard_plan_chunked(total_kmers, k, n_ranks, 1, plan);
}

extern "C" int kmer_cuda_shard_plan_chunked(uint64_t total_kmers, int k, uint32_t n_ranks, uint32_t chunks_per_rank,
                                            kmer_shard_plan* plan) {
    if (!plan || n_ranks < 1 || n_ranks > 16 || chunks_per_rank < 1 || n_ranks * chunks_per_rank > 32 || k < 14 || k > KMER_CUDA_MAX_K)
        return KMER_ERR_BAD_ARGUMENT;
    // fine buckets as on one GPU (about 1000 k-mers each), 2^fine_shift of them per coarse partition
    PartitionPlan p = make_partition_plan(total_kmers, k);
    uint64_t fine_per_rank = ((uint64_t)p.n_buckets + n_ranks - 1) / n_ranks;
    uint32_t fine_shift = 8;
    while (fine_shift > 0 && (1ull << fine_shift) > fine_per_rank) fine_shift--;
    uint64_t coarse_per_rank = (fine_per_rank + (1ull << fine_shift) - 1) >> fine_shift;
    if (coarse_per_rank * n_ranks << fine_shift > 0x7fffffffull) return KMER_ERR_BAD_ARGUMENT;
    memset(plan, 0, sizeof(*plan));
    plan->n_ranks = n_ranks;
    plan->buckets_per_rank = (uint32_t)coarse_per_rank;
    plan->n_buckets = plan->buckets_per_rank * n_ranks;
    plan->fine_shift = fine_shift;
    plan->fine_cap = p.cap;
    plan->chunks_per_rank = chunks_per_rank;
    plan->k = k;
    plan->w = p.w; plan->m = p.m; plan->recw = p.recw; plan->rmax = p.rmax;
    plan->rec_bytes = p.recw == 1 ? 8 : 16;
    {   // records per (coarse partition, source): 1/n_ranks of a partition's k-mers, about 2.1/(w+1) records per k-mer
        double kmers_per_part = (double)total_kmers / (double)plan->n_buckets;
        double rpk = 2.1 / (p.w + 1) + (p.rmax < p.w ? 1.0 / p.rmax : 0.0);
        double mean = kmers_per_part * rpk / (n_ranks * chunks_per_rank);
        double cap = 1.15 * mean + 6.0 * sqrt(3.0 * mean) + 64.0;
        plan->cap = ((uint32_t)cap + 1u) & ~1u;   // even: every (partition, source) segment starts 16-byte aligned
    }
    plan->recs_bytes_per_peer = (uint64_t)plan->buckets_per_rank * plan->cap * plan->rec_bytes;
    plan->fill_bytes_per_peer = (uint64_t)plan->buckets_per_rank * 8;
    return KMER_OK;
}

// the source side scatters into coarse partitions; the hash range is the global number of fine buckets
static PartitionPlan coarse_partition_plan(const kmer_shard_plan* sp) {
    PartitionPlan p{};
    p.n_buckets = sp->n_buckets;
    p.hash_buckets = sp->n_buckets << sp->fine_shift;
    p.fine_shift = (int)sp->fine_shift;
    p.cap = sp->cap;
    p.w = sp->w; p.m = sp->m; p.recw = sp->recw; p.rmax = sp->rmax;
    p.spill_cap = 0;   // no spill list on the source side: a full segment is an error reported by finish()
    p.debug = 0;
    return p;
}
// the owner side: this GPU's fine buckets (local numbering), counted like a single-GPU batch
static PartitionPlan fine_partition_plan(const kmer_shard_plan* sp) {
    PartitionPlan p{};
    p.n_buckets = sp->buckets_per_rank << sp->fine_shift;
    p.hash_buckets = sp->n_buckets << sp->fine_shift;
    p.fine_shift = (int)sp->fine_shift;
    p.cap = sp->fine_cap;
    p.w = sp->w; p.m = sp->m; p.recw = sp->recw; p.rmax = sp->rmax;
    uint64_t sc = (uint64_t)p.n_buckets * p.cap / 8;
    p.spill_cap = sc < 4096 ? 4096 : sc;
    p.debug = 0;
    return p;
}

extern "C" int kmer_cuda_dev_shard_partition(kmer_cuda_ctx* c, const char* d_seq, uint64_t n_bases, const uint64_t* d_row_off,
                                             uint64_t n_rows, const kmer_shard_plan* sp, void* d_send_recs,
                                             uint64_t* d_send_fill, void* stream) {
    if (!c || !sp) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_SHARD_PART;
    c->p_n_bases = n_bases; c->p_n_rows = n_rows; c->p_k = sp->k;
    c->p_expected_kmers = kmer_cuda_max_kmers(n_bases, n_rows, sp->k);
    PartitionPlan plan = coarse_partition_plan(sp);
    if (n_rows == 0 || n_bases == 0) {
        if (n_rows) return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
        CU(cudaMemsetAsync(d_send_fill, 0, (size_t)sp->n_buckets * 8, st), "memset fill");
        return KMER_OK;
    }
    ScanArgs a;
    rc = prepare_rows(c, d_row_off, n_bases, n_rows, sp->k, st, &a, d_seq);
    if (rc) return rc;
    launch_partition(c->di, a, plan, (unsigned long long*)d_send_fill, d_send_recs, nullptr, st);
    mark(c, st, "minimizer_partition");
    resolve_bad_row_kernel<<<1, 1, 0, st>>>(c->d_status, d_row_off, n_rows);
    c->launches += 2;
    CU(cudaGetLastError(), "shard partition launch");
    return KMER_OK;
}

static int dev_shard_count_impl(kmer_cuda_ctx* c, const kmer_shard_plan* sp, const void* d_recv_recs, const uint64_t* d_recv_fill,
                                uint64_t* d_uniq, uint64_t uniq_capacity, kmer_count_pair* d_pairs, uint64_t pairs_capacity,
                                void* stream) {
    if (!c || !sp) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_SHARD_COUNT;
    c->last_overflow = 0;
    c->last_tier2 = 0;
    const int k = sp->k;
    if ((sp->cap & 1u) || sp->fine_shift > 12) return bad_arg(c, "shard plan: cap must be even, fine_shift <= 12");
    // coarse partitions (one segment per source GPU) -> this GPU's fine buckets, then counted like a single-GPU batch
    PartitionPlan plan = fine_partition_plan(sp);
    rc = ws(c, c->fill, (size_t)plan.n_buckets * 8);
    if (!rc) rc = ws(c, c->recs, partition_record_bytes(plan));
    if (!rc) rc = ws(c, c->spill, partition_spill_bytes(plan));
    if (!rc) rc = ws(c, c->failed, (size_t)plan.n_buckets * 4);
    if (rc) return rc;
    const int n_src = (int)(sp->n_ranks * (sp->chunks_per_rank ? sp->chunks_per_rank : 1u));
    launch_refine(c->di, plan, k, n_src, sp->buckets_per_rank, sp->cap, (const unsigned long long*)d_recv_fill,
                  d_recv_recs, (unsigned long long*)c->fill.p, c->recs.p, c->spill.p, c->d_status, st);
    mark(c, st, "refine");
    launch_bucket_count(c->di, plan, k, (const unsigned long long*)c->fill.p, c->recs.p, c->spill.p, (uint32_t*)c->failed.p, d_pairs,
                        pairs_capacity, d_uniq, d_uniq ? uniq_capacity : 0, c->d_status, st);
    mark(c, st, "bucket_count");
    c->launches += 2;
    CU(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, st), "D2H status");
    CU(cudaStreamSynchronize(st), "stream sync");
    const DevStatus hs = *c->h_status;
    if (hs.n_overflow != 0) {
        c->last_overflow = hs.n_overflow;   // even the spill list overflowed: reported by finish() as KMER_ERR_CAPACITY
    } else if (hs.n_failed || hs.n_spill) {   // tier 2: only the buckets that did not fit on chip
        uint64_t n_slots = next_pow2(std::max<uint64_t>(1024, (hs.failed_kmers + hs.n_spill * 16) * 2));
        rc = ws(c, c->table, n_slots * sizeof(kmer_count_pair));
        if (rc) return rc;
        launch_hash_clear((kmer_count_pair*)c->table.p, n_slots, st);
        launch_partition_tier2(c->di, plan, k, 1, (const unsigned long long*)c->fill.p, c->recs.p, c->spill.p,
                               (const uint32_t*)c->failed.p, (kmer_count_pair*)c->table.p, n_slots, c->d_status, st);
        mark(c, st, "tier2_insert");
        launch_hash_compact(c->di, (const kmer_count_pair*)c->table.p, n_slots, k, d_pairs, pairs_capacity, c->d_status, st);
        mark(c, st, "tier2_compact");
        c->launches += 2;
        c->last_tier2 = hs.failed_kmers;
    } else {
        launch_append_special(d_pairs, pairs_capacity, c->d_status, st);
        c->launches++;
    }
    CU(cudaGetLastError(), "shard count launch");
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_shard_count(kmer_cuda_ctx* c, const kmer_shard_plan* sp, const void* d_recv_recs,
                                         const uint64_t* d_recv_fill, kmer_count_pair* d_pairs, uint64_t pairs_capacity,
                                         void* stream) {
    return dev_shard_count_impl(c, sp, d_recv_recs, d_recv_fill, nullptr, 0, d_pairs, pairs_capacity, stream);
}

extern "C" int kmer_cuda_dev_shard_count_split(kmer_cuda_ctx* c, const kmer_shard_plan* sp, const void* d_recv_recs,
                                               const uint64_t* d_recv_fill, uint64_t* d_uniq, uint64_t uniq_capacity,
                                               kmer_count_pair* d_pairs, uint64_t pairs_capacity, void* stream) {
    if (c && !d_uniq && uniq_capacity) return bad_arg(c, "d_uniq");
    return dev_shard_count_impl(c, sp, d_recv_recs, d_recv_fill, d_uniq, uniq_capacity, d_pairs, pairs_capacity, stream);
}

extern "C" int kmer_cuda_dev_dense_table(kmer_cuda_ctx* c, const char* d_seq, uint64_t n_bases, const uint64_t* d_row_off,
                                         uint64_t n_rows, int k, uint64_t* d_table, void* stream) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_DENSE_TABLE;
    c->p_n_bases = n_bases; c->p_n_rows = n_rows; c->p_k = k;
    c->p_expected_kmers = kmer_cuda_max_kmers(n_bases, n_rows, k);
    if (n_rows && (k < 1 || k > 13)) {
        if (k < 1 || k > KMER_CUDA_MAX_K) return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
        return bad_arg(c, "dense counting needs k <= 13");
    }
    if (n_rows == 0 || n_bases == 0) {
        if (n_rows) return ref_error(&c->err, KMER_ERR_INVALID_K, 0);
        if (k >= 1 && k <= 13) CU(cudaMemsetAsync(d_table, 0, (size_t)8 << (2 * k), st), "memset table");
        return KMER_OK;
    }
    ScanArgs a;
    rc = prepare_rows(c, d_row_off, n_bases, n_rows, k, st, &a, d_seq);
    if (rc) return rc;
    launch_dense_table(c->di, a, (unsigned long long*)d_table, st);
    mark(c, st, "count_dense");
    resolve_bad_row_kernel<<<1, 1, 0, st>>>(c->d_status, d_row_off, n_rows);
    c->launches += 2;
    CU(cudaGetLastError(), "dense table launch");
    return KMER_OK;
}

extern "C" int kmer_cuda_dev_dense_emit(kmer_cuda_ctx* c, const uint64_t* d_table, int k, uint32_t rank, uint32_t n_ranks,
                                        kmer_count_pair* d_pairs, uint64_t pairs_capacity, void* stream) {
    if (!c) return KMER_ERR_BAD_ARGUMENT;
    cudaStream_t st = pick_stream(c, stream);
    int rc = begin_op(c, st);
    if (rc) return rc;
    c->pending = OP_SHARD_COUNT;
    c->last_overflow = 0;
    c->last_tier2 = 0;
    if (k < 1 || k > 13 || n_ranks < 1 || rank >= n_ranks) return bad_arg(c, "dense emit: k in 1..13, rank < n_ranks");
    launch_dense_emit(c->di, (const unsigned long long*)d_table, k, rank, n_ranks, d_pairs, pairs_capacity, c->d_status, st);
    mark(c, st, "dense_emit");
    c->launches++;
    CU(cudaGetLastError(), "dense emit launch");
    return KMER_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer batch submit

static int h2d(kmer_cuda_ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (!bytes) return KMER_OK;
    // pinned (or otherwise device-accessible) source: one asynchronous DMA; pageable: the driver stages it
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st), "H2D copy");
    return KMER_OK;
}

static int upload_rows(kmer_cuda_ctx* c, const char* seq, const uint64_t* row_off, uint64_t n_rows, uint64_t* n_bases) {
    if (n_rows && !row_off) return bad_arg(c, "row_off");
    *n_bases = n_rows ? row_off[n_rows] : 0;
    if (n_rows && row_off[0] != 0) return bad_arg(c, "row_off[0] must be 0");
    if (*n_bases && !seq) return bad_arg(c, "seq");
    int rc = ws(c, c->seq, ((*n_bases + 15) & ~15ull) + 64);
    if (rc) return rc;
    rc = ws(c, c->off, (n_rows + 1) * 8);
    if (rc) return rc;
    rc = h2d(c, c->seq.p, seq, *n_bases, c->stream);
    if (rc) return rc;
    if (n_rows) rc = h2d(c, c->off.p, row_off, (n_rows + 1) * 8, c->stream);
    return rc;
}

extern "C" int kmer_cuda_submit_extract(kmer_cuda_ctx* c, const char* seq, const uint64_t* row_off, uint64_t n_rows, int k,
                                        uint64_t** codes, uint64_t* n_kmers) {
    if (!c || !codes || !n_kmers) return KMER_ERR_BAD_ARGUMENT;
    *codes = nullptr;
    *n_kmers = 0;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    uint64_t n_bases = 0;
    int rc = upload_rows(c, seq, row_off, n_rows, &n_bases);
    if (rc) return rc;
    uint64_t cap = kmer_cuda_max_kmers(n_bases, n_rows, k);
    rc = ws(c, c->codes, cap * 8);
    if (rc) return rc;
    rc = kmer_cuda_dev_extract(c, (const char*)c->seq.p, n_bases, (const uint64_t*)c->off.p, n_rows, k,
                               (uint64_t*)c->codes.p, cap, KMER_OWN_STREAM);
    if (rc) return rc;
    kmer_dev_result res;
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, &res);
    if (rc) return rc;
    uint64_t* out = (uint64_t*)pinned_get(c, res.n_kmers * 8);
    if (!out) return c->err.status;
    CU(cudaMemcpyAsync(out, c->codes.p, res.n_kmers * 8, cudaMemcpyDeviceToHost, c->stream), "D2H codes");
    CU(cudaStreamSynchronize(c->stream), "stream sync");
    *codes = out;
    *n_kmers = res.n_kmers;
    return KMER_OK;
}

extern "C" int kmer_cuda_submit_count(kmer_cuda_ctx* c, const char* seq, const uint64_t* row_off, uint64_t n_rows, int k,
                                      kmer_count_pair** pairs, uint64_t* n_distinct, uint64_t* n_kmers) {
    if (!c || !pairs || !n_distinct) return KMER_ERR_BAD_ARGUMENT;
    *pairs = nullptr;
    *n_distinct = 0;
    if (n_kmers) *n_kmers = 0;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    uint64_t n_bases = 0;
    int rc = upload_rows(c, seq, row_off, n_rows, &n_bases);
    if (rc) return rc;
    uint64_t cap = kmer_cuda_max_kmers(n_bases, n_rows, k);
    if (k >= 1 && k < 32 && (1ull << (2 * k)) < cap) cap = 1ull << (2 * k);
    rc = ws(c, c->pairs, cap * sizeof(kmer_count_pair));
    if (rc) return rc;
    rc = kmer_cuda_dev_count(c, (const char*)c->seq.p, n_bases, (const uint64_t*)c->off.p, n_rows, k,
                             (kmer_count_pair*)c->pairs.p, cap, 0, KMER_OWN_STREAM);
    if (rc) return rc;
    kmer_dev_result res;
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, &res);
    if (rc) return rc;
    kmer_count_pair* out = (kmer_count_pair*)pinned_get(c, res.n_distinct * sizeof(kmer_count_pair));
    if (!out) return c->err.status;
    CU(cudaMemcpyAsync(out, c->pairs.p, res.n_distinct * sizeof(kmer_count_pair), cudaMemcpyDeviceToHost, c->stream),
       "D2H pairs");
    CU(cudaStreamSynchronize(c->stream), "stream sync");
    *pairs = out;
    *n_distinct = res.n_distinct;
    if (n_kmers) *n_kmers = res.n_kmers;
    return KMER_OK;
}

extern "C" int kmer_cuda_submit_count_split(kmer_cuda_ctx* c, const char* seq, const uint64_t* row_off, uint64_t n_rows, int k,
                                            uint64_t** uniq_codes, uint64_t* n_unique, kmer_count_pair** pairs,
                                            uint64_t* n_pairs, uint64_t* n_kmers) {
    if (!c || !uniq_codes || !n_unique || !pairs || !n_pairs) return KMER_ERR_BAD_ARGUMENT;
    *uniq_codes = nullptr; *pairs = nullptr;
    *n_unique = 0; *n_pairs = 0;
    if (n_kmers) *n_kmers = 0;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    uint64_t n_bases = 0;
    int rc = upload_rows(c, seq, row_off, n_rows, &n_bases);
    if (rc) return rc;
    uint64_t cap = kmer_cuda_max_kmers(n_bases, n_rows, k);
    if (k >= 1 && k < 32 && (1ull << (2 * k)) < cap) cap = 1ull << (2 * k);
    // a k-mer that is not unique occurs at least twice: at most cap/2 pairs next to the unique codes -- unless a fallback
    // tier (which writes pairs only) takes over, so the pair buffer keeps the full size
    rc = ws(c, c->pairs, cap * sizeof(kmer_count_pair));
    if (!rc) rc = ws(c, c->codes, cap * sizeof(uint64_t));
    if (rc) return rc;
    rc = kmer_cuda_dev_count_split(c, (const char*)c->seq.p, n_bases, (const uint64_t*)c->off.p, n_rows, k, (uint64_t*)c->codes.p, cap,
                                   (kmer_count_pair*)c->pairs.p, cap, KMER_OWN_STREAM);
    if (rc) return rc;
    kmer_dev_result res;
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, &res);
    if (rc) return rc;
    uint64_t* out_u = (uint64_t*)pinned_get(c, res.n_unique * sizeof(uint64_t));
    kmer_count_pair* out_p = (kmer_count_pair*)pinned_get(c, res.n_distinct * sizeof(kmer_count_pair));
    if (!out_u || !out_p) return c->err.status;
    CU(cudaMemcpyAsync(out_u, c->codes.p, res.n_unique * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream), "D2H codes");
    CU(cudaMemcpyAsync(out_p, c->pairs.p, res.n_distinct * sizeof(kmer_count_pair), cudaMemcpyDeviceToHost, c->stream), "D2H pairs");
    CU(cudaStreamSynchronize(c->stream), "stream sync");
    *uniq_codes = out_u;
    *n_unique = res.n_unique;
    *pairs = out_p;
    *n_pairs = res.n_distinct;
    if (n_kmers) *n_kmers = res.n_kmers;
    return KMER_OK;
}

extern "C" int kmer_cuda_submit_count_packed(kmer_cuda_ctx* c, const char* seq, const uint64_t* row_off, uint64_t n_rows, int k,
                                             uint8_t** uniq_packed, uint64_t* n_unique, int* code_bytes,
                                             kmer_count_pair** pairs, uint64_t* n_pairs, uint64_t* n_kmers) {
    if (!c || !uniq_packed || !n_unique || !code_bytes || !pairs || !n_pairs) return KMER_ERR_BAD_ARGUMENT;
    *uniq_packed = nullptr; *pairs = nullptr;
    *n_unique = 0; *n_pairs = 0; *code_bytes = 0;
    if (n_kmers) *n_kmers = 0;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    uint64_t n_bases = 0;
    int rc = upload_rows(c, seq, row_off, n_rows, &n_bases);
    if (rc) return rc;
    uint64_t cap = kmer_cuda_max_kmers(n_bases, n_rows, k);
    if (k >= 1 && k < 32 && (1ull << (2 * k)) < cap) cap = 1ull << (2 * k);
    rc = ws(c, c->pairs, cap * sizeof(kmer_count_pair));
    if (!rc) rc = ws(c, c->codes, cap * sizeof(uint64_t));
    if (rc) return rc;
    rc = kmer_cuda_dev_count_split(c, (const char*)c->seq.p, n_bases, (const uint64_t*)c->off.p, n_rows, k, (uint64_t*)c->codes.p, cap,
                                   (kmer_count_pair*)c->pairs.p, cap, KMER_OWN_STREAM);
    if (rc) return rc;
    kmer_dev_result res;
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, &res);
    if (rc) return rc;
    const int nbytes = k >= 1 ? (2 * k + 7) / 8 : 1;
    const size_t packed_bytes = (size_t)res.n_unique * (size_t)nbytes;
    rc = ws(c, c->text, packed_bytes + 16);
    if (rc) return rc;
    launch_pack_codes(c->di, (const uint64_t*)c->codes.p, res.n_unique, nbytes, (uint8_t*)c->text.p, c->stream);
    c->launches++;
    uint8_t* out_u = (uint8_t*)pinned_get(c, packed_bytes);
    kmer_count_pair* out_p = (kmer_count_pair*)pinned_get(c, res.n_distinct * sizeof(kmer_count_pair));
    if (!out_u || !out_p) return c->err.status;
    CU(cudaMemcpyAsync(out_u, c->text.p, packed_bytes, cudaMemcpyDeviceToHost, c->stream), "D2H packed codes");
    CU(cudaMemcpyAsync(out_p, c->pairs.p, res.n_distinct * sizeof(kmer_count_pair), cudaMemcpyDeviceToHost, c->stream), "D2H pairs");
    CU(cudaStreamSynchronize(c->stream), "stream sync");
    *uniq_packed = out_u;
    *n_unique = res.n_unique;
    *code_bytes = nbytes;
    *pairs = out_p;
    *n_pairs = res.n_distinct;
    if (n_kmers) *n_kmers = res.n_kmers;
    return KMER_OK;
}

extern "C" int kmer_cuda_submit_match(kmer_cuda_ctx* c, int op, const int* ops, const uint64_t* codes, const uint8_t* lens,
                                      uint64_t m, int k, const char* const* consts, uint32_t n_consts, uint32_t** bits,
                                      uint64_t* words_per_row, uint64_t** hits) {
    if (!c || !bits || !words_per_row) return KMER_ERR_BAD_ARGUMENT;
    *bits = nullptr;
    if (hits) *hits = nullptr;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    if (m && !codes) return bad_arg(c, "codes");
    if (n_consts && !consts) return bad_arg(c, "consts");
    uint64_t wpr = (m + 31) / 32;
    *words_per_row = wpr;
    int rc = ws(c, c->codes, m * 8);
    if (!rc && lens) rc = ws(c, c->lens, m);
    if (!rc) rc = ws(c, c->bits, (size_t)n_consts * wpr * 4);
    if (!rc) rc = ws(c, c->hits, (size_t)n_consts * 8);
    if (rc) return rc;
    rc = h2d(c, c->codes.p, codes, m * 8, c->stream);
    if (!rc && lens) rc = h2d(c, c->lens.p, lens, m, c->stream);
    if (rc) return rc;
    rc = kmer_cuda_dev_match(c, op, ops, (const uint64_t*)c->codes.p, lens ? (const uint8_t*)c->lens.p : nullptr, m, k,
                             consts, n_consts, (uint32_t*)c->bits.p, (uint64_t*)c->hits.p, KMER_OWN_STREAM);
    if (rc) return rc;
    size_t bits_bytes = (size_t)n_consts * wpr * 4;
    char* out = (char*)pinned_get(c, bits_bytes);
    if (!out) return c->err.status;
    if (bits_bytes) CU(cudaMemcpyAsync(out, c->bits.p, bits_bytes, cudaMemcpyDeviceToHost, c->stream), "D2H bits");
    uint64_t* hout = (uint64_t*)pinned_get(c, (size_t)n_consts * 8);
    if (!hout) return c->err.status;
    if (n_consts) CU(cudaMemcpyAsync(hout, c->hits.p, (size_t)n_consts * 8, cudaMemcpyDeviceToHost, c->stream), "D2H hits");
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, nullptr);
    if (rc) return rc;
    *bits = (uint32_t*)out;
    if (hits) *hits = hout; else kmer_cuda_release(c, hout);
    return KMER_OK;
}

extern "C" int kmer_cuda_submit_decode(kmer_cuda_ctx* c, const uint64_t* codes, uint64_t n, int k, int with_header,
                                       char** text) {
    if (!c || !text) return KMER_ERR_BAD_ARGUMENT;
    *text = nullptr;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    if (k < 0 || k > KMER_CUDA_MAX_K) return bad_arg(c, "k-mer length must be 0..32");
    size_t bytes = (size_t)n * (size_t)(k + (with_header ? 1 : 0));
    int rc = ws(c, c->codes, n * 8);
    if (!rc) rc = ws(c, c->text, bytes);
    if (rc) return rc;
    rc = h2d(c, c->codes.p, codes, n * 8, c->stream);
    if (rc) return rc;
    rc = kmer_cuda_dev_decode(c, (const uint64_t*)c->codes.p, n, k, with_header, (char*)c->text.p, KMER_OWN_STREAM);
    if (rc) return rc;
    char* out = (char*)pinned_get(c, bytes);
    if (!out) return c->err.status;
    if (bytes) CU(cudaMemcpyAsync(out, c->text.p, bytes, cudaMemcpyDeviceToHost, c->stream), "D2H text");
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, nullptr);
    if (rc) return rc;
    *text = out;
    return KMER_OK;
}

extern "C" int kmer_cuda_submit_encode(kmer_cuda_ctx* c, const char* text, const uint8_t* lens, uint64_t n, int stride,
                                       uint64_t** codes) {
    if (!c || !codes) return KMER_ERR_BAD_ARGUMENT;
    *codes = nullptr;
    CU(cudaSetDevice(c->di.device), "cudaSetDevice");
    if (stride < 0) return bad_arg(c, "stride");
    // kmer_in checks the length before the alphabet (kmer.c:115-122)
    if (lens) {
        for (uint64_t i = 0; i < n; i++)
            if (lens[i] > KMER_CUDA_MAX_K) return ref_error(&c->err, KMER_ERR_KMER_TOO_LONG, (int64_t)i);
    } else if (stride > KMER_CUDA_MAX_K && n)
        return ref_error(&c->err, KMER_ERR_KMER_TOO_LONG, 0);
    size_t bytes = (size_t)n * (size_t)stride;
    int rc = ws(c, c->text, bytes);
    if (!rc) rc = ws(c, c->codes, n * 8);
    if (!rc && lens) rc = ws(c, c->lens, n);
    if (rc) return rc;
    rc = h2d(c, c->text.p, text, bytes, c->stream);
    if (!rc && lens) rc = h2d(c, c->lens.p, lens, n, c->stream);
    if (rc) return rc;
    rc = begin_op(c, c->stream);
    if (rc) return rc;
    c->pending = OP_ENCODE;
    launch_encode(c->di, (const char*)c->text.p, lens ? (const uint8_t*)c->lens.p : nullptr, n, stride,
                  (uint64_t*)c->codes.p, c->d_status, c->stream);
    c->launches++;
    uint64_t* out = (uint64_t*)pinned_get(c, n * 8);
    if (!out) return c->err.status;
    if (n) CU(cudaMemcpyAsync(out, c->codes.p, n * 8, cudaMemcpyDeviceToHost, c->stream), "D2H codes");
    rc = kmer_cuda_dev_finish(c, KMER_OWN_STREAM, nullptr);
    if (rc) {
        kmer_cuda_release(c, out);
        return rc;
    }
    *codes = out;
    return KMER_OK;
}
