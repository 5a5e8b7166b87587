// kssd_b200.cu -- host side of libkssd_b200.so: contexts, launch sequences and the C-ABI declared in
// include/kssd_b200.h.  CUDA runtime only (no torch, no CPU compute path: every entry point needs
// the device).  Sorting/scans of the (tiny) post-sampling streams use CUB from the CUDA toolkit;
// the byte scan, the index expansion and the count/statistics kernels are ours.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/kssd_b200.h"
#include "index_dist.cuh"
#include "sketch_scan3.cuh"
#include "sketch_fastq.cuh"
#include "sketch_fastq3.cuh"
#include "sketch_buckets.cuh"
#include "set_ops.cuh"
#include "composite.cuh"
#include "dist_text.cuh"

using namespace kssd;

// ------------------------------------------------------------------------------------------------
// errors, bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return fail(KSSD_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define LAUNCHED(n) g_launches.fetch_add((n), std::memory_order_relaxed)

extern "C" const char *kssd_last_error(void) { return g_err.c_str(); }
extern "C" const char *kssd_version(void) { return "kssd-b200 0.1 (sm_100a)"; }
extern "C" uint64_t kssd_kernel_launch_count(void) { return g_launches.load(); }

static const uint32_t kPrimer[25] = {   // reference global_basic.c:74-81
    251u, 509u, 1021u, 2039u, 4093u, 8191u, 16381u, 32749u, 65521u, 131071u, 262139u, 524287u, 1048573u,
    2097143u, 4194301u, 8388593u, 16777213u, 33554393u, 67108859u, 134217689u, 268435399u, 536870909u,
    1073741789u, 2147483647u, 4294967291u};

// grow-only device scratch buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// stream-ordered scratch allocations of one call, released on every exit path
struct StreamScratch {
    cudaStream_t stream;
    std::vector<void *> ptrs;
    explicit StreamScratch(cudaStream_t s) : stream(s) {}
    ~StreamScratch() { for (void *p : ptrs) cudaFreeAsync(p, stream); }
    void *alloc(size_t bytes)
    {
        void *p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(bytes, 16), stream) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return p;
    }
    StreamScratch(const StreamScratch &) = delete;
    StreamScratch &operator=(const StreamScratch &) = delete;
};

struct kssd_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    SketchParams P{};
    kssd_ctx_info_t info{};
    uint32_t *d_prefilter = nullptr;
    uint32_t *d_prefilter3 = nullptr;            // block bitmap of the lazy scan (sketch_scan3.cuh)
    unsigned long long *d_gtab = nullptr;        // its second level: per block, the members among the three windows around it
    int scan_stride = 3;                         // bases per first-level probe of that scan (3 when 2*subk >= 12, else 1)
    int scan_impl = 3;                           // KSSD_SCAN_IMPL=2 selects the previous formulation (A/B runs)
    uint2 *d_ht = nullptr;
    uint32_t *d_flag = nullptr;                  // set by the validation kernels: data taken from disk is never trusted as an index
    // scratch
    DevBuf seq, meta, plan, keys, ords, keys2, ords2, flags, pos, runs, keep, counts, minord, cubtmp, misc;
    uint8_t *stag[2] = {nullptr, nullptr};       // pinned staging buffers of kssd_stage1_files
    uint64_t stag_cap[2] = {0, 0};
    DevBuf gzin, gztext;                         // .gz batches decoded on the GPU: compressed bytes, decoded text (stage1_files.cuh)
    // cached span plan of the last batch layout
    bool plan_valid = false;
    uint64_t plan_key = 0;
    int plan_genomes = 0;
    uint32_t plan_spans = 0;
    // bucket layout of the same batch layout (sketch_buckets.cuh): capacities from the genome lengths
    bool plan_buckets = false;
    uint32_t plan_bucket_total = 0, plan_bucket_maxcap = 0;
    DevBuf bplan, bwork;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};      // [4], [5]: the distance.out text kernels
    float last_ms[6] = {0, 0, 0, 0, 0, 0};
    char *h_text = nullptr;                      // pinned: the text of kssd_dist_text (grow-only)
    size_t h_text_cap = 0;
    bool stats_ms_pending = false;               // last_ms[4] of a sparse search: read off ev[2] .. ev[1] on demand
    bool total_ms_pending = false;               // last_ms[1] of a bucket-mode sketch: read off ev[0] .. ev[2] on demand
    // Sketch-size arrays of reference sets already on the device (Stage III): the reference's query-batch loop
    // (command_dist.c:763-790) searches the same references batch after batch, and a pageable 400 KB upload per batch
    // would make the host wait for the stream.  Keyed by content (memcmp); an entry lives while a job uses it.
    struct SizeSet { uint32_t refs = 0; bool any_empty = false; uint32_t *dev = nullptr; std::shared_ptr<std::vector<uint32_t>> host; };
    std::vector<SizeSet> size_sets;
    // pinned result slots + events of kssd_dist_stats_async (cudaMallocHost / cudaFreeHost synchronise the device: never per job)
    uint64_t *async_page = nullptr;
    std::vector<int> async_free;
    std::vector<cudaEvent_t> async_events, async_counted;
    cudaStream_t stream2 = nullptr;              // the rows pass of search b runs here while the count kernel of search b + 1 runs on `stream`
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
__global__ void count_sampled_kernel(const int32_t *__restrict__ shuf, uint32_t n, uint32_t dim_end, uint32_t *count)
{
    uint32_t c = 0;
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
        const int32_t v = shuf[d];
        c += (v >= 0 && (uint32_t)v < dim_end) ? 1u : 0u;
    }
    c = __reduce_add_sync(kFull, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// S = {d : shuf[d] < dim_end}: exact hash table d -> pf, and the prefilter bitmap of S u RC(S)
__global__ void build_sampled_kernel(const int32_t *__restrict__ shuf, uint32_t n, uint32_t dim_end, int s,
                                     uint32_t *__restrict__ prefilter, uint32_t *__restrict__ prefilter3, unsigned long long *__restrict__ gtab,
                                     int stride, uint2 *__restrict__ ht, uint32_t ht_mask)
{
    const int wbits = 4 * s;                                   // width of the inner 2s-mer
    const uint32_t reps = wbits < kPfBitShift + 5 ? 1u << (kPfBitShift + 5 - wbits) : 1u;
    const uint32_t reps3 = wbits < 20 ? 1u << (20 - wbits) : 1u;
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
        const int32_t v = shuf[d];
        if (v < 0 || (uint32_t)v >= dim_end) continue;
        const uint32_t both[2] = {d, (uint32_t)revcomp2((uint64_t)d, 2 * s)};
#pragma unroll
        for (int k = 0; k < 2; k++) {
            // first level (layout in kssd_device.cuh); probes may carry bases beyond the window in bits >= 4s,
            // so a narrow window is entered under every value of those bits
            for (uint32_t hi = 0; hi < reps; hi++) {
                const uint32_t x = both[k] | (hi << wbits);
                atomicOr(&prefilter[x & kPfWordMask], 0x80000000u >> ((x >> kPfBitShift) & 31));
            }
            const uint32_t i2 = pf2_index(both[k]);
            atomicOr(&prefilter[kPfWords + (i2 >> 5)], 1u << (i2 & 31));
            // lazy scan (sketch_scan3.cuh): the window in the scan representation; its 10-base blocks at offsets
            // 0 .. stride-1 (word = bits 0-14, flag = 0x80000000 >> bits 15-19), and per block and offset the folded rest
            const uint32_t y = (uint32_t)to_scan_repr(both[k], 2 * s);
            for (int r = 0; r < stride; r++)
                for (uint32_t hi = 0; hi < reps3; hi++) {
                    const uint32_t b = ((y >> (2 * r)) | (hi << wbits)) & 0xfffffu;
                    atomicOr(&prefilter3[b & 0x7fffu], 0x80000000u >> (b >> 15));
                    if (gtab) atomicOr(&gtab[b], 1ull << (16 * r + gtab_ext(y, r)));
                }
        }
        uint32_t h = mix32(d) & ht_mask;
        for (;;) {
            const uint32_t old = atomicCAS(&ht[h].x, kHtEmpty, d);
            if (old == kHtEmpty) { ht[h].y = (uint32_t)v; break; }
            h = (h + 1) & ht_mask;
        }
    }
}

extern "C" int kssd_ctx_create(kssd_ctx_t **out, int device, const int32_t *shuf_table, int k, int subk, int drlevel,
                               int component_sz)
{
    if (!out || !shuf_table) return fail(KSSD_E_INVAL, "kssd_ctx_create: null argument");
    if (subk < 1 || subk >= 8 || k < subk || k > 16 || drlevel < 0 || drlevel > subk || k - subk > 15 || component_sz < 1 ||
        component_sz > 7)
        return fail(KSSD_E_INVAL, "kssd_ctx_create: unsupported (k=%d, subk=%d, drlevel=%d, COMPONENT_SZ=%d)", k, subk, drlevel,
                    component_sz);
    const int pi = 4 * (k - drlevel) - 8 - 7;   // CTX_SPC_USE_L = 8, command_dist.c:220
    if (pi < 0 || pi > 24) return fail(KSSD_E_PRIMER, "get_hashsz(): primer_ind: %d out of range(0 ~ 24)", pi);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(KSSD_E_CUDA, "kssd_ctx_create: no CUDA device %d (found %d); this library has no CPU path", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(KSSD_E_CUDA, "kssd_ctx_create: device %d is sm_%d%d, need sm_100", device, prop.major, prop.minor);

    kssd_ctx *c = new kssd_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // keep stream-ordered allocations cached in the pool instead of returning them at every sync
        cudaMemPool_t pool;
        CU(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto &e : c->ev) CU(cudaEventCreate(&e));

    SketchParams &P = c->P;
    P.k = k; P.s = subk; P.L = drlevel;
    P.TL = 2 * k;
    P.out = k - subk;
    P.hist_min_n = (P.TL - 1 + 1) / 2;
    P.tupmask = ~0ull >> (64 - 4 * k);
    P.undomask = ((1ull << (2 * P.out)) - 1ull) << (2 * (k + subk));
    P.outmask = (1ull << (2 * P.out)) - 1ull;
    P.innermask = (uint32_t)((1ull << (4 * subk)) - 1ull);
    const uint64_t subspace = 1ull << (4 * (subk - drlevel));
    P.dim_end = (uint32_t)std::max<uint64_t>(subspace, 4096);   // MIN_SUBCTX_DIM_SMP_SZ
    P.comp_code_bits = (k - drlevel > component_sz) ? 4 * (k - drlevel - component_sz) : 0;
    P.comp_mask = (1u << P.comp_code_bits) - 1u;

    const uint32_t n = 1u << (4 * subk);
    int32_t *d_shuf = nullptr;
    uint32_t *d_cnt = nullptr;
    CU(cudaMalloc(&d_shuf, (size_t)n * 4));
    CU(cudaMalloc(&d_cnt, 4));
    CU(cudaMemcpyAsync(d_shuf, shuf_table, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(d_cnt, 0, 4, c->stream));
    count_sampled_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_shuf, n, P.dim_end, d_cnt);
    LAUNCHED(1);
    uint32_t n_sampled = 0;
    CU(cudaMemcpyAsync(&n_sampled, d_cnt, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    uint32_t ht_size = 1024;
    while (ht_size < 2 * (uint64_t)n_sampled + 16) ht_size <<= 1;
    P.ht_mask = ht_size - 1;
    CU(cudaMalloc(&c->d_prefilter, (kPfWords + kPf2Words) * 4));
    CU(cudaMalloc(&c->d_ht, (size_t)ht_size * sizeof(uint2)));
    CU(cudaMalloc(&c->d_flag, 16));
    CU(cudaMemsetAsync(c->d_flag, 0, 16, c->stream));
    CU(cudaMemsetAsync(c->d_prefilter, 0, (kPfWords + kPf2Words) * 4, c->stream));
    CU(cudaMemsetAsync(c->d_ht, 0xff, (size_t)ht_size * sizeof(uint2), c->stream));
    CU(cudaMalloc(&c->d_prefilter3, kPf3Words * 4));
    CU(cudaMemsetAsync(c->d_prefilter3, 0, kPf3Words * 4, c->stream));
    c->scan_stride = 2 * subk >= 12 ? 3 : 1;
    if (c->scan_stride == 3) {
        CU(cudaMalloc(&c->d_gtab, (size_t)kGtabEntries * 8));
        CU(cudaMemsetAsync(c->d_gtab, 0, (size_t)kGtabEntries * 8, c->stream));
    }
    if (const char *e = getenv("KSSD_SCAN_IMPL")) c->scan_impl = atoi(e) == 2 ? 2 : 3;
    build_sampled_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(d_shuf, n, P.dim_end, subk, c->d_prefilter, c->d_prefilter3,
                                                                c->d_gtab, c->scan_stride, c->d_ht, P.ht_mask);
    LAUNCHED(1);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    cudaFree(d_shuf);
    cudaFree(d_cnt);
    P.prefilter = c->d_prefilter;
    P.ht = c->d_ht;
    P.gtab = c->d_gtab;

    CU(cudaFuncSetAttribute(sketch_fasta32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kScanSmemBytes));
    CU(cudaFuncSetAttribute(sketch_fasta3_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScan3SmemBytes));
    CU(cudaFuncSetAttribute(sketch_fasta3_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScan3SmemBytes));
    CU(cudaFuncSetAttribute(sketch_fasta3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScan3SmemBytes));
    CU(cudaFuncSetAttribute(sketch_fasta3_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScan3SmemBytes));

    kssd_ctx_info_t &I = c->info;
    I.k = k; I.subk = subk; I.drlevel = drlevel; I.component_sz = component_sz;
    I.component_num = 1 << P.comp_code_bits;
    I.comp_code_bits = P.comp_code_bits;
    I.dim_end = P.dim_end;
    I.hashsize = kPrimer[pi];
    I.hashlimit = (uint32_t)(I.hashsize * 0.6);   // LD_FCTR
    I.n_sampled = n_sampled;
    I.device = device;
    I.sm_count = c->sm_count;
    *out = c;
    return KSSD_OK;
}

extern "C" void kssd_ctx_destroy(kssd_ctx_t *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf *b : {&c->seq, &c->meta, &c->plan, &c->keys, &c->ords, &c->keys2, &c->ords2, &c->flags, &c->pos, &c->runs, &c->keep, &c->bplan, &c->bwork, &c->counts, &c->minord,
                      &c->cubtmp, &c->misc, &c->gzin, &c->gztext})
        b->release();
    for (int b = 0; b < 2; b++) if (c->stag[b]) cudaFreeHost(c->stag[b]);
    for (auto &e : c->size_sets) cudaFree(e.dev);
    if (c->async_page) cudaFreeHost(c->async_page);
    if (c->h_text) cudaFreeHost(c->h_text);
    for (cudaEvent_t e : c->async_events) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->async_counted) if (e) cudaEventDestroy(e);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    cudaFree(c->d_prefilter);
    cudaFree(c->d_prefilter3);
    cudaFree(c->d_gtab);
    cudaFree(c->d_ht);
    cudaFree(c->d_flag);
    for (auto &e : c->ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int kssd_ctx_info(const kssd_ctx_t *c, kssd_ctx_info_t *info)
{
    if (!c || !info) return fail(KSSD_E_INVAL, "kssd_ctx_info: null argument");
    *info = c->info;
    return KSSD_OK;
}
extern "C" void *kssd_ctx_stream(const kssd_ctx_t *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int kssd_ctx_sync(const kssd_ctx_t *c)
{
    if (!c) return fail(KSSD_E_INVAL, "kssd_ctx_sync: null");
    CU(cudaStreamSynchronize(c->stream));
    if (c->stream2) CU(cudaStreamSynchronize(c->stream2));      // the rows passes of queued searches (kssd_dist_stats_async)
    return KSSD_OK;
}
extern "C" float kssd_ctx_last_ms(const kssd_ctx_t *cc, int which)
{
    kssd_ctx *c = const_cast<kssd_ctx *>(cc);
    if (!c || which < 0 || which >= 6) return -1.f;
    if (which == 4 && c->stats_ms_pending) {
        cudaSetDevice(c->device);
        if (cudaEventSynchronize(c->ev[1]) == cudaSuccess) cudaEventElapsedTime(&c->last_ms[4], c->ev[2], c->ev[1]);
        c->stats_ms_pending = false;
    }
    if (which == 1 && c->total_ms_pending) {
        cudaSetDevice(c->device);
        if (cudaEventSynchronize(c->ev[2]) == cudaSuccess) cudaEventElapsedTime(&c->last_ms[1], c->ev[0], c->ev[2]);
        c->total_ms_pending = false;
    }
    return c->last_ms[which];
}

// ------------------------------------------------------------------------------------------------
// Stage I
// ------------------------------------------------------------------------------------------------
struct kssd_sketch {
    kssd_ctx *ctx = nullptr;
    int n_genomes = 0, n_comp = 1, mode = 0;
    uint64_t n_occ = 0, total = 0;
    float scan_ms = 0;
    uint8_t *d_blob = nullptr;                   // one stream-ordered allocation: ids | ord | abund | index
    size_t blob_bytes = 0;
    uint32_t *d_ids = nullptr;
    uint16_t *d_abund = nullptr;
    uint64_t *d_ord = nullptr;
    uint64_t *d_index = nullptr;                 // n_comp * (n_genomes+1)
    std::vector<uint64_t> comp_start;            // n_comp+1 offsets into d_ids
    std::vector<uint64_t> index;                 // n_comp * (n_genomes+1)
    std::vector<int32_t> status;
    std::vector<uint64_t> n_reads;               // KSSD_MODE_BYREAD: '>' records per file
};

// Post-pass over the sorted occurrence keys, run by run (a run = one distinct (component, genome, id)).  Nothing here
// walks a run serially: a k-mer repeated a million times costs what a million distinct ones cost.
//   run_heads_kernel    flags[i] = key i opens a run                      -> exclusive scan -> run number of every key
//   run_reduce_kernel   first position of every run; first occurrence (min offset) by a segmented warp reduction
//   run_keep_kernel     multiplicity = distance to the next run, keep rule per mode, per-genome distinct-key tally
//   sketch_scatter_kernel  kept runs -> ids / abundances / first-occurrence offsets, per (component, genome) tally
__global__ void run_heads_kernel(const uint64_t *__restrict__ keys, uint32_t n, uint32_t *__restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i - 1] != keys[i]) ? 1u : 0u;
}

__global__ void run_reduce_kernel(const uint64_t *__restrict__ ords, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos, uint32_t n,
                                  uint32_t *__restrict__ runstart, unsigned long long *__restrict__ minord)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool valid = i < n;
    const uint32_t f = valid ? flags[i] : 0u;
    const uint32_t r = valid ? pos[i] + f - 1u : 0xffffffffu;
    unsigned long long v = valid ? ords[i] : ~0ull;
    if (f) runstart[r] = i;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {      // segmented min: runs are contiguous, so equal run numbers o lanes apart bound a whole segment
        const unsigned long long ov = __shfl_down_sync(kFull, v, o);
        const uint32_t orr = __shfl_down_sync(kFull, r, o);
        if (lane + o < 32 && orr == r && ov < v) v = ov;
    }
    const uint32_t pr = __shfl_up_sync(kFull, r, 1);
    if (valid && (lane == 0 || pr != r)) atomicMin(&minord[r], v);
}

__global__ void run_keep_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos,
                                const uint32_t *__restrict__ runstart, uint32_t n, int mode, int M, uint32_t *__restrict__ keep,
                                uint16_t *__restrict__ counts, uint32_t *__restrict__ distinct_pg)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t nruns = pos[n - 1] + flags[n - 1];
    if (i >= nruns) { keep[i] = 0; return; }
    const uint32_t s = runstart[i], e = i + 1 < nruns ? runstart[i + 1] : n, cnt = e - s;
    bool k = true;
    if (mode == KSSD_MODE_FASTA_UNIQ) k = cnt == 1;          // iseq2comem.c:694-695 + :540
    else if (mode == KSSD_MODE_FASTQ) k = cnt >= (uint32_t)M; // iseq2comem.c:336-346 + :514
    keep[i] = k ? 1u : 0u;
    counts[i] = (uint16_t)min(cnt, 65535u);                   // iseq2comem.c:602-604
    atomicAdd(&distinct_pg[(uint32_t)(keys[s] >> 28) & 0x0fffffffu], 1u);   // every distinct key took a slot (keycount, :262 / :689)
}

__global__ void sketch_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos,
                                      const uint32_t *__restrict__ runstart, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ outpos,
                                      const uint16_t *__restrict__ counts, const unsigned long long *__restrict__ minord, uint32_t n, int n_genomes,
                                      uint32_t *__restrict__ ids, uint16_t *__restrict__ abund, uint64_t *__restrict__ ord,
                                      uint32_t *__restrict__ per_cg)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t nruns = pos[n - 1] + flags[n - 1];
    if (i >= nruns || !keep[i]) return;
    const uint64_t key = keys[runstart[i]];
    const uint32_t p = outpos[i];
    ids[p] = (uint32_t)(key & 0x0fffffffu);
    abund[p] = counts[i];
    ord[p] = minord[i];
    const uint32_t comp = (uint32_t)(key >> 56), gid = (uint32_t)(key >> 28) & 0x0fffffffu;
    atomicAdd(&per_cg[(uint64_t)comp * n_genomes + gid], 1u);
}

// total kept = outpos[n-1] + keep[n-1], left in the slot after the per-(component, genome) counters
__global__ void sketch_total_kernel(const uint32_t *__restrict__ keep, const uint32_t *__restrict__ outpos, uint32_t n, uint32_t *__restrict__ slot)
{
    *slot = outpos[n - 1] + keep[n - 1];
}

// FASTQ: per genome, index the lines, then one thread per record.  The index is built in ONE pass over the text
// (nl_index_kernel: decoupled look-back, 32-bit offsets, counts left on the device -- no host round trip before the walk);
// files of 4 GiB and more, or with lines so short that the index outgrows its buffer (`single_pass` false after the retry),
// take the two-pass index (count, scan, fill) with 64-bit positions.
static int fastq_run(kssd_ctx *c, const uint8_t *d_seq, const uint64_t *goff, const uint64_t *glen, int n_genomes, int mode, int Q,
                     const ScanArgs &A, bool single_pass)
{
    const SketchParams &P = c->P;
    CU(cudaFuncSetAttribute(sketch_fastq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((kPfWords + kPf2Words) * 4)));
    CU(cudaFuncSetAttribute(sketch_fastq3_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFq3SmemBytes));
    CU(cudaFuncSetAttribute(sketch_fastq3_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFq3SmemBytes));
    CU(cudaFuncSetAttribute(sketch_fastq3_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFq3SmemBytes));
    CU(cudaFuncSetAttribute(sketch_fastq3_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFq3SmemBytes));
    for (int g = 0; g < n_genomes; g++) {
        const uint64_t gs = goff[g], ge = gs + glen[g];
        if (ge == gs) continue;
        const uint64_t a0 = gs & ~15ull;
        if (single_pass && ge - a0 < 0xffffffffull) {
            const uint64_t nblk = (ge - a0 + kNlxBytesPerBlock - 1) / kNlxBytesPerBlock;
            const uint64_t lines_cap = std::max<uint64_t>((ge - a0) / 8, 4096);          // lines of 8 bytes on average: 10x the usual read
            CU(c->flags.ensure(nblk * 8 + 64));
            CU(c->minord.ensure(lines_cap * 4));
            unsigned long long *state = c->flags.as<unsigned long long>();
            NlIndexOut *io = reinterpret_cast<NlIndexOut *>(state + nblk);
            CU(cudaMemsetAsync(state, 0, nblk * 8 + sizeof(NlIndexOut), c->stream));
            const bool dbg = getenv("KSSD_FASTQ_TIMING") != nullptr;      // per-kernel times of this path on stderr (profiles/fastq_scale.py)
            if (dbg && getenv("KSSD_FASTQ_FLUSH")) { CU(c->keys2.ensure(256u << 20)); CU(cudaMemsetAsync(c->keys2.p, 1, 256u << 20, c->stream)); }
            if (dbg) CU(cudaEventRecord(c->ev[2], c->stream));
            nl_index_kernel<<<(uint32_t)nblk, kNlBlock, 0, c->stream>>>(d_seq, gs, ge, a0, (uint32_t)nblk, state, io, c->minord.as<uint32_t>(), (uint32_t)lines_cap,
                                                                          A.ticket, dbg && getenv("KSSD_NLX_NOLB") ? 1 : 0);
            FastqArgs F{};
            F.seq = d_seq; F.seq_bytes = A.seq_bytes; F.gs = gs; F.ge = ge;
            F.nlpos32 = c->minord.as<uint32_t>(); F.pos_base = a0; F.idx = io;
            F.gid = (uint32_t)g;
            F.abund = mode == KSSD_MODE_FASTQ_ABUND;
            F.Q = Q;
            F.line_cap = F.abund ? 4094u : 19998u;
            F.out_keys = A.out_keys; F.out_ords = A.out_ords; F.out_cap = A.out_cap; F.out_count = A.out_count; F.gstatus = A.gstatus;
            const uint64_t want = ((ge - gs) / 64 + kFastqThreads - 1) / kFastqThreads;   // records are not counted yet: >= 64 bytes each as a guess
            if (dbg) CU(cudaEventRecord(c->ev[3], c->stream));
            // the warp walk (block prefilter, lazy codes) whenever the qualities cannot matter: -A, or -Q <= 0 and no byte >= 0x80 in the
            // file -- the second condition is known on the device only, so both kernels are queued and one of them returns at once
            const bool warp_walk = (F.abund || Q <= 0) && !getenv("KSSD_FASTQ_THREAD_WALK");
            if (warp_walk) {
                ScanArgs A3 = A;
                A3.strict_window = 1;
                const bool big = 2 * (P.TL - 1) >= 32;
                F.warp_walk = 1;
                if (c->scan_stride == 3 && big) sketch_fastq3_kernel<3, true><<<c->sm_count, kScanThreads, kFq3SmemBytes, c->stream>>>(P, A3, F, c->d_prefilter3);
                else if (c->scan_stride == 3) sketch_fastq3_kernel<3, false><<<c->sm_count, kScanThreads, kFq3SmemBytes, c->stream>>>(P, A3, F, c->d_prefilter3);
                else if (big) sketch_fastq3_kernel<1, true><<<c->sm_count, kScanThreads, kFq3SmemBytes, c->stream>>>(P, A3, F, c->d_prefilter3);
                else sketch_fastq3_kernel<1, false><<<c->sm_count, kScanThreads, kFq3SmemBytes, c->stream>>>(P, A3, F, c->d_prefilter3);
                LAUNCHED(1);
            }
            sketch_fastq_kernel<<<(uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)c->sm_count), kFastqThreads, (kPfWords + kPf2Words) * 4, c->stream>>>(P, F);
            LAUNCHED(2);
            if (dbg) {
                cudaEvent_t e4;
                CU(cudaEventCreate(&e4));
                CU(cudaEventRecord(e4, c->stream));
                CU(cudaEventSynchronize(e4));
                float t_ix = 0, t_walk = 0;
                cudaEventElapsedTime(&t_ix, c->ev[2], c->ev[3]);
                cudaEventElapsedTime(&t_walk, c->ev[3], e4);
                fprintf(stderr, "kssd fastq genome %d: line index %.3f ms, walk %.3f ms\n", g, t_ix, t_walk);
                cudaEventDestroy(e4);
            }
            continue;
        }
        const uint64_t nblk = (ge - a0 + kNlBytesPerBlock - 1) / kNlBytesPerBlock;
        if (nblk > 0x7fffffffull) return fail(KSSD_E_INVAL, "kssd_sketch_batch: FASTQ text of genome %d too large", g);
        CU(c->flags.ensure(nblk * 4));
        CU(c->pos.ensure(nblk * 4));
        nl_count_kernel<false><<<(uint32_t)nblk, kNlBlock, 0, c->stream>>>(d_seq, gs, ge, a0, c->flags.as<uint32_t>());
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), nblk, c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), nblk, c->stream));
        LAUNCHED(3);
        uint32_t lastoff = 0, lastcnt = 0;
        uint8_t lastbyte = 0;
        CU(cudaMemcpyAsync(&lastoff, c->pos.as<uint32_t>() + (nblk - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&lastcnt, c->flags.as<uint32_t>() + (nblk - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&lastbyte, d_seq + ge - 1, 1, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const uint64_t n_nl = (uint64_t)lastoff + lastcnt;
        CU(c->minord.ensure(std::max<uint64_t>(n_nl, 1) * 8));
        nl_fill_kernel<false><<<(uint32_t)nblk, kNlBlock, 0, c->stream>>>(d_seq, gs, ge, a0, c->pos.as<uint32_t>(), c->minord.as<uint64_t>());
        LAUNCHED(1);
        FastqArgs F{};
        F.seq = d_seq; F.seq_bytes = A.seq_bytes; F.gs = gs; F.ge = ge;
        F.nlpos = c->minord.as<uint64_t>();
        F.n_nl = n_nl;
        F.n_lines = n_nl + (lastbyte != '\n' ? 1 : 0);
        F.n_records = (F.n_lines + 3) / 4;
        F.gid = (uint32_t)g;
        F.abund = mode == KSSD_MODE_FASTQ_ABUND;
        F.Q = Q;
        F.line_cap = F.abund ? 4094u : 19998u;     // fgets(…, 4096 / 20000, …) keeps len+1 <= cap-1 in one piece
        F.out_keys = A.out_keys; F.out_ords = A.out_ords; F.out_cap = A.out_cap; F.out_count = A.out_count; F.gstatus = A.gstatus;
        if (F.n_records) {
            const uint64_t want = (F.n_records + kFastqThreads - 1) / kFastqThreads;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)c->sm_count);
            sketch_fastq_kernel<<<grid, kFastqThreads, (kPfWords + kPf2Words) * 4, c->stream>>>(P, F);
            LAUNCHED(1);
        }
    }
    return KSSD_OK;
}

// --byread (reference reads2mco, iseq2comem.c:78-186) after the scan: header starts of every file, occurrences in
// stream order per (component, file), the record number of each, and the per-(component, file) totals.
static int byread_finish(kssd_ctx *c, kssd_sketch *S, const uint8_t *d_seq, const uint64_t *goff, const uint64_t *glen, int n_genomes,
                         uint32_t n_occ, const uint64_t *d_goff, const int32_t *d_gstatus)
{
    const SketchParams &P = c->P;
    const int n_comp = S->n_comp;
    if (n_genomes >= (1 << 20) || n_comp > 256) return fail(KSSD_E_INVAL, "kssd_sketch_batch: --byread takes < 2^20 files and <= 256 components");
    std::vector<uint64_t> boff(n_genomes + 1, 0);
    for (int g = 0; g < n_genomes; g++) {
        if (glen[g] >> 36) return fail(KSSD_E_INVAL, "kssd_sketch_batch: --byread file %d larger than 64 GiB", g);
        const uint64_t a0 = goff[g] & ~15ull;
        boff[g + 1] = boff[g] + (glen[g] ? (goff[g] + glen[g] - a0 + kNlBytesPerBlock - 1) / kNlBytesPerBlock : 0);
    }
    const uint64_t nblk = boff[n_genomes];
    if (nblk > 0x7fffffffull) return fail(KSSD_E_INVAL, "kssd_sketch_batch: --byread batch too large");
    S->n_reads.assign(n_genomes, 0);
    std::vector<uint64_t> hdr_off(n_genomes + 1, 0);
    if (nblk) {
        CU(c->flags.ensure(nblk * 4));
        CU(c->pos.ensure(nblk * 4));
        for (int g = 0; g < n_genomes; g++)
            if (boff[g + 1] > boff[g]) {
                nl_count_kernel<true><<<(uint32_t)(boff[g + 1] - boff[g]), kNlBlock, 0, c->stream>>>(d_seq, goff[g], goff[g] + glen[g], goff[g] & ~15ull,
                                                                                                  c->flags.as<uint32_t>() + boff[g]);
                LAUNCHED(1);
            }
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), nblk, c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), nblk, c->stream));
        LAUNCHED(2);
        std::vector<uint32_t> pos(nblk);
        uint32_t lastcnt = 0;
        CU(cudaMemcpyAsync(pos.data(), c->pos.p, nblk * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&lastcnt, c->flags.as<uint32_t>() + (nblk - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const uint64_t total_hdr = (uint64_t)pos[nblk - 1] + lastcnt;
        for (int g = 0; g <= n_genomes; g++) hdr_off[g] = boff[g] < nblk ? pos[boff[g]] : total_hdr;
        for (int g = 0; g < n_genomes; g++) S->n_reads[g] = hdr_off[g + 1] - hdr_off[g];
        CU(c->minord.ensure(std::max<uint64_t>(total_hdr, 1) * 8));
        for (int g = 0; g < n_genomes; g++)
            if (boff[g + 1] > boff[g]) {
                nl_fill_kernel<true><<<(uint32_t)(boff[g + 1] - boff[g]), kNlBlock, 0, c->stream>>>(d_seq, goff[g], goff[g] + glen[g], goff[g] & ~15ull,
                                                                                                 c->pos.as<uint32_t>() + boff[g], c->minord.as<uint64_t>());
                LAUNCHED(1);
            }
    }
    S->status.assign(n_genomes, 0);
    S->comp_start.assign(n_comp + 1, 0);
    S->index.assign((size_t)n_comp * (n_genomes + 1), 0);
    std::vector<uint32_t> per_cg((size_t)n_comp * n_genomes, 0);
    const size_t nmax = std::max<size_t>(n_occ, 1);
    const size_t o_ids = 0, o_ord = (nmax * 4 + 15) & ~(size_t)15, o_idx = o_ord + nmax * 8;
    S->blob_bytes = o_idx + S->index.size() * 8;
    CU(cudaMallocAsync(&S->d_blob, S->blob_bytes, c->stream));
    S->d_ids = reinterpret_cast<uint32_t *>(S->d_blob + o_ids);
    S->d_ord = reinterpret_cast<uint64_t *>(S->d_blob + o_ord);     // record number of every occurrence in this mode
    S->d_abund = nullptr;
    S->d_index = reinterpret_cast<uint64_t *>(S->d_blob + o_idx);
    if (n_occ) {
        CU(c->keys2.ensure((size_t)n_occ * 8));
        CU(c->ords2.ensure((size_t)n_occ * 8));
        const uint32_t nb = (n_occ + 255) / 256;
        byread_rekey_kernel<<<nb, 256, 0, c->stream>>>(c->keys.as<uint64_t>(), c->ords.as<uint64_t>(), n_occ, c->keys2.as<uint64_t>(),
                                                       c->ords2.as<uint32_t>());
        int gbits = 1;
        while ((1ll << gbits) < n_genomes) gbits++;
        const int end_bit = P.comp_code_bits ? 56 + P.comp_code_bits : 36 + gbits;
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->keys2.as<uint64_t>(), c->keys.as<uint64_t>(), c->ords2.as<uint32_t>(), S->d_ids, n_occ, 0,
                                        end_bit, c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp, c->keys2.as<uint64_t>(), c->keys.as<uint64_t>(), c->ords2.as<uint32_t>(), S->d_ids,
                                           n_occ, 0, end_bit, c->stream));
        CU(c->misc.ensure(per_cg.size() * 4));
        CU(cudaMemsetAsync(c->misc.p, 0, per_cg.size() * 4, c->stream));
        CU(c->counts.ensure(hdr_off.size() * 8));
        CU(cudaMemcpyAsync(c->counts.p, hdr_off.data(), hdr_off.size() * 8, cudaMemcpyHostToDevice, c->stream));
        byread_assign_kernel<<<nb, 256, 0, c->stream>>>(c->keys.as<uint64_t>(), n_occ, d_goff, c->minord.as<uint64_t>(), c->counts.as<uint64_t>(),
                                                        n_genomes, S->d_ord, c->misc.as<uint32_t>());
        LAUNCHED(10);
        CU(cudaMemcpyAsync(per_cg.data(), c->misc.p, per_cg.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(S->status.data(), d_gstatus, 4ull * n_genomes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    uint64_t run = 0;
    for (int cc = 0; cc < n_comp; cc++) {
        S->comp_start[cc] = run;
        uint64_t *ix = &S->index[(size_t)cc * (n_genomes + 1)];
        for (int g = 0; g < n_genomes; g++) ix[g + 1] = ix[g] + per_cg[(size_t)cc * n_genomes + g];
        run += ix[n_genomes];
    }
    S->comp_start[n_comp] = run;
    S->total = run;
    for (int g = 0; g < n_genomes; g++) S->status[g] = (S->status[g] & 1) ? KSSD_E_HEADER_EOF : 0;
    CU(cudaMemcpyAsync(S->d_index, S->index.data(), S->index.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev[2], c->stream));
    CU(cudaEventSynchronize(c->ev[2]));
    CU(cudaEventElapsedTime(&c->last_ms[1], c->ev[0], c->ev[2]));
    return KSSD_OK;
}

static int sketch_run(kssd_ctx *c, const uint8_t *d_seq, size_t seq_bytes, const uint64_t *goff, const uint64_t *glen, int n_genomes,
                      const kssd_sketch_opts_t *opts, kssd_sketch_t **out)
{
    const SketchParams &P = c->P;
    const int mode = opts ? opts->mode : KSSD_MODE_FASTA;
    if (mode < KSSD_MODE_FASTA || mode > KSSD_MODE_BYREAD) return fail(KSSD_E_INVAL, "kssd_sketch_batch: unknown mode %d", mode);
    const bool is_fastq = mode == KSSD_MODE_FASTQ || mode == KSSD_MODE_FASTQ_ABUND;
    if (mode == KSSD_MODE_FASTQ && opts && (opts->M < 1 || opts->M >= 15))
        return fail(KSSD_E_INVAL, "fastq2co(): Occurence num should smaller than 15");   // iseq2comem.c:279
    if (n_genomes <= 0 || n_genomes >= (1 << 28)) return fail(KSSD_E_INVAL, "kssd_sketch_batch: n_genomes=%d", n_genomes);
    uint64_t total = 0;
    for (int g = 0; g < n_genomes; g++) {
        if (goff[g] % 16 || goff[g] + glen[g] > seq_bytes)
            return fail(KSSD_E_INVAL, "kssd_sketch_batch: genome %d extent [%llu,+%llu) misaligned or outside the buffer", g,
                        (unsigned long long)goff[g], (unsigned long long)glen[g]);
        total += glen[g];
    }
    // Spans: the unit a warp pulls from the ticket.  Guided sizes: big spans for the bulk (the per-span start-up --
    // boundary search, one masked iteration -- is ~3 us), a quarter of that for the last round and a sixteenth for the
    // last quarter round, so the tail of the dynamic schedule is short.  opts->span_bytes fixes one size (tests).
    const uint64_t warps = (uint64_t)c->sm_count * kScanWarps;
    uint32_t span = opts ? opts->span_bytes : 0;
    const bool guided = span == 0;
    if (guided) {
        uint64_t want = total / (warps * 4) + 1;
        span = 4096;
        while (span < want && span < (1u << 20)) span <<= 1;
    }
    if (span < 512 || (span & (span - 1))) return fail(KSSD_E_INVAL, "kssd_sketch_batch: span_bytes must be a power of two >= 512");
    // The span plan is a function of the batch layout only; a host that sketches batches of the same layout (a
    // streaming pipeline with fixed-size staging buffers) pays for it once: the plan and its device copy are cached.
    uint64_t lkey = 1469598103934665603ull;
    auto mixkey = [&](uint64_t v) { lkey = (lkey ^ v) * 1099511628211ull; };
    mixkey((uint64_t)n_genomes); mixkey(span); mixkey(guided);
    for (int g = 0; g < n_genomes; g++) { mixkey(goff[g]); mixkey(glen[g]); }
    const bool plan_hit = c->plan_valid && c->plan_key == lkey && c->plan_genomes == n_genomes;
    if (!plan_hit) {
        std::vector<uint32_t> span_gid;
        std::vector<uint64_t> span_nom;
        uint64_t remaining = total;
        const uint64_t round = warps * span;
        for (int g = 0; g < n_genomes; g++)
            for (uint64_t o = 0; o < glen[g];) {
                uint32_t sz = span;
                if (guided) {
                    if (remaining <= round / 4) sz = std::max<uint32_t>(span / 16, 4096);
                    else if (remaining <= round) sz = std::max<uint32_t>(span / 4, 4096);
                }
                span_gid.push_back((uint32_t)g);
                span_nom.push_back(goff[g] + o);
                const uint64_t step = std::min<uint64_t>(sz, glen[g] - o);
                o += step;
                remaining -= step;
            }
        c->plan_spans = (uint32_t)span_gid.size();
        // device plan: goff | glen | span_nom | span_gid
        const size_t p_glen = 8ull * n_genomes, p_nom = p_glen + 8ull * n_genomes, p_sgid = p_nom + 8ull * c->plan_spans,
                     p_end = p_sgid + 4ull * c->plan_spans;
        CU(c->plan.ensure(p_end));
        uint8_t *pb = c->plan.as<uint8_t>();
        CU(cudaMemcpyAsync(pb, goff, 8ull * n_genomes, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(pb + p_glen, glen, 8ull * n_genomes, cudaMemcpyHostToDevice, c->stream));
        if (c->plan_spans) {
            CU(cudaMemcpyAsync(pb + p_nom, span_nom.data(), 8ull * c->plan_spans, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(pb + p_sgid, span_gid.data(), 4ull * c->plan_spans, cudaMemcpyHostToDevice, c->stream));
        }
        {   // bucket capacities: twice the expected occurrences of a (component, genome) plus slack (sketch_buckets.cuh)
            const int n_comp_b = c->info.component_num;
            const double rate_b = (double)c->info.n_sampled / (double)(1ull << (4 * P.s)) / n_comp_b;
            const uint64_t NB = (uint64_t)n_comp_b * n_genomes;
            std::vector<uint32_t> boff(NB + 1, 0);
            uint64_t run = 0;
            uint32_t maxcap = 0;
            bool ok = NB < (1u << 24);
            for (int cc = 0; cc < n_comp_b && ok; cc++)
                for (int g = 0; g < n_genomes; g++) {
                    const double e = (double)glen[g] * rate_b;
                    const uint64_t cap = ((uint64_t)(2.0 * e + 8.0 * sqrt(e)) + 64 + 1) & ~1ull;
                    boff[(size_t)cc * n_genomes + g] = (uint32_t)run;
                    run += cap;
                    maxcap = (uint32_t)std::max<uint64_t>(maxcap, cap);
                    if (cap > kBucketMaxCap || run > 0xfffffff0ull) { ok = false; break; }
                }
            boff[NB] = (uint32_t)run;
            c->plan_buckets = ok;
            c->plan_bucket_total = (uint32_t)run;
            c->plan_bucket_maxcap = maxcap;
            if (ok) {
                CU(c->bplan.ensure((NB + 1) * 4));
                CU(cudaMemcpyAsync(c->bplan.p, boff.data(), (NB + 1) * 4, cudaMemcpyHostToDevice, c->stream));
            }
            CU(cudaStreamSynchronize(c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));      // the vectors die here
        c->plan_key = lkey; c->plan_genomes = n_genomes; c->plan_valid = true;
    }
    const uint32_t n_spans = c->plan_spans;
    uint8_t *pb = c->plan.as<uint8_t>();
    const size_t p_glen = 8ull * n_genomes, p_nom = p_glen + 8ull * n_genomes, p_sgid = p_nom + 8ull * n_spans;

    // per-call device scratch: gstatus | occurrences of code 0 dropped per genome | ticket | out_count
    const size_t m_stat = 0, m_zero = m_stat + 4ull * n_genomes, m_tick = m_zero + 4ull * n_genomes, m_cnt = m_tick + 4, m_end = m_cnt + 4;
    CU(c->meta.ensure(m_end));
    uint8_t *mb = c->meta.as<uint8_t>();
    CU(cudaMemsetAsync(mb + m_stat, 0, m_end, c->stream));

    // ---- bucket mode (sketch_buckets.cuh): no global sort, one host round trip.  FASTA modes of the lazy scan, when every
    // (component, genome) bucket fits one CTA's shared memory; an overflowing bucket sends the batch to the list mode below.
    if (!is_fastq && mode != KSSD_MODE_BYREAD && c->scan_impl == 3 && c->plan_buckets && n_spans && !getenv("KSSD_NO_BUCKETS")) {
        const int n_comp = c->info.component_num;
        const uint32_t NB = (uint32_t)n_comp * (uint32_t)n_genomes;
        // work area: bcnt[NB] | kept[NB + 1] | foff[NB + 1] | distinct[G] | overflow | n_occ
        const size_t w_cnt = 0, w_kept = w_cnt + 4ull * NB, w_foff = w_kept + 4ull * (NB + 1), w_dist = w_foff + 4ull * (NB + 1),
                     w_ovf = w_dist + 4ull * n_genomes, w_end = w_ovf + 8;
        CU(c->bwork.ensure(w_end));
        uint8_t *wb = c->bwork.as<uint8_t>();
        CU(cudaMemsetAsync(wb, 0, w_end, c->stream));
        CU(c->keys.ensure((size_t)c->plan_bucket_total * 8));
        CU(c->flags.ensure((size_t)c->plan_bucket_total * 4));
        CU(c->counts.ensure((size_t)c->plan_bucket_total * 2));
        CU(c->minord.ensure((size_t)c->plan_bucket_total * 8));
        ScanArgs A{};
        A.seq = d_seq; A.seq_bytes = seq_bytes;
        A.goff = reinterpret_cast<uint64_t *>(pb);
        A.glen = reinterpret_cast<uint64_t *>(pb + p_glen);
        A.span_nom = reinterpret_cast<uint64_t *>(pb + p_nom);
        A.span_gid = reinterpret_cast<uint32_t *>(pb + p_sgid);
        A.n_spans = n_spans;
        A.gstatus = reinterpret_cast<int32_t *>(mb + m_stat);
        A.zero_count = reinterpret_cast<uint32_t *>(mb + m_zero);
        A.ticket = reinterpret_cast<uint32_t *>(mb + m_tick);
        A.out_count = reinterpret_cast<uint32_t *>(mb + m_cnt);
        A.drop_zero = 1;
        A.bkeys = c->keys.as<uint64_t>();
        A.boff = c->bplan.as<uint32_t>();
        A.bcnt = reinterpret_cast<uint32_t *>(wb + w_cnt);
        A.boverflow = reinterpret_cast<uint32_t *>(wb + w_ovf);
        A.n_genomes = (uint32_t)n_genomes;
        CU(cudaEventRecord(c->ev[0], c->stream));
        const bool big = 2 * (P.TL - 1) >= 32;
        if (c->scan_stride == 3 && big) sketch_fasta3_kernel<3, true><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
        else if (c->scan_stride == 3) sketch_fasta3_kernel<3, false><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
        else if (big) sketch_fasta3_kernel<1, true><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
        else sketch_fasta3_kernel<1, false><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
        CU(cudaEventRecord(c->ev[1], c->stream));
        uint32_t P2 = 32;
        while (P2 < c->plan_bucket_maxcap) P2 <<= 1;
        const size_t bsm = (size_t)P2 * 10;
        CU(cudaFuncSetAttribute(bucket_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
        uint32_t *d_kept = reinterpret_cast<uint32_t *>(wb + w_kept), *d_foff = reinterpret_cast<uint32_t *>(wb + w_foff);
        bucket_finish_kernel<<<NB, kBucketThreads, bsm, c->stream>>>(A.bkeys, A.boff, A.bcnt, A.boverflow, (uint32_t)n_genomes, mode, opts ? opts->M : 1,
                                                                   c->flags.as<uint32_t>(), c->counts.as<uint16_t>(), c->minord.as<uint64_t>(), d_kept,
                                                                   reinterpret_cast<uint32_t *>(wb + w_dist), reinterpret_cast<uint32_t *>(wb + w_ovf) + 1);
        size_t tmpb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmpb, d_kept, d_foff, NB + 1, c->stream);
        CU(c->cubtmp.ensure(tmpb));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmpb, d_kept, d_foff, NB + 1, c->stream));
        LAUNCHED(4);
        std::vector<uint32_t> foff(NB + 1), distinct(n_genomes), zeros(n_genomes);
        uint32_t ovf[2] = {0, 0};
        kssd_sketch *S = new kssd_sketch();
        S->ctx = c; S->n_genomes = n_genomes; S->n_comp = n_comp; S->mode = mode;
        S->status.assign(n_genomes, 0);
        CU(cudaMemcpyAsync(foff.data(), d_foff, 4ull * (NB + 1), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(distinct.data(), wb + w_dist, 4ull * n_genomes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(ovf, wb + w_ovf, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(S->status.data(), mb + m_stat, 4ull * n_genomes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(zeros.data(), mb + m_zero, 4ull * n_genomes, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        if (ovf[0]) {
            delete S;                                     // a bucket overflowed: list mode redoes the batch (counters are reset there)
            CU(cudaMemsetAsync(mb + m_stat, 0, m_end, c->stream));
        } else {
            CU(cudaEventElapsedTime(&S->scan_ms, c->ev[0], c->ev[1]));
            c->last_ms[0] = S->scan_ms;
            const uint32_t total = foff[NB];
            S->total = total;
            S->comp_start.assign(n_comp + 1, 0);
            S->index.assign((size_t)n_comp * (n_genomes + 1), 0);
            const size_t nmax = std::max<size_t>(total, 1);
            const size_t o_ids = 0, o_ord = (nmax * 4 + 15) & ~(size_t)15, o_ab = o_ord + nmax * 8, o_idx = (o_ab + nmax * 2 + 15) & ~(size_t)15;
            S->blob_bytes = o_idx + S->index.size() * 8;
            CU(cudaMallocAsync(&S->d_blob, S->blob_bytes, c->stream));
            S->d_ids = reinterpret_cast<uint32_t *>(S->d_blob + o_ids);
            S->d_ord = reinterpret_cast<uint64_t *>(S->d_blob + o_ord);
            S->d_abund = reinterpret_cast<uint16_t *>(S->d_blob + o_ab);
            S->d_index = reinterpret_cast<uint64_t *>(S->d_blob + o_idx);
            if (total) {
                bucket_move_kernel<<<NB, 128, 0, c->stream>>>(A.boff, d_foff, c->flags.as<uint32_t>(), c->counts.as<uint16_t>(), c->minord.as<uint64_t>(),
                                                             S->d_ids, S->d_abund, S->d_ord);
                LAUNCHED(1);
            }
            for (int cc = 0; cc < n_comp; cc++) {
                const uint32_t *f = &foff[(size_t)cc * n_genomes];
                S->comp_start[cc] = f[0];
                uint64_t *ix = &S->index[(size_t)cc * (n_genomes + 1)];
                for (int g = 0; g <= n_genomes; g++) ix[g] = (uint64_t)(f[g] - f[0]);
            }
            S->comp_start[n_comp] = total;
            for (int g = 0; g < n_genomes; g++) {
                if (S->status[g] & 1) S->status[g] = KSSD_E_HEADER_EOF;
                else if ((uint64_t)distinct[g] + zeros[g] > c->info.hashlimit) S->status[g] = KSSD_E_CROWD;      // keycount (see the list mode below)
                else S->status[g] = 0;
            }
            S->n_occ = ovf[1];
            CU(cudaMemcpyAsync(S->d_index, S->index.data(), S->index.size() * 8, cudaMemcpyHostToDevice, c->stream));
            CU(cudaEventRecord(c->ev[2], c->stream));      // (kssd_ctx_last_ms(1) waits for it; the call itself does not)
            c->total_ms_pending = true;
            *out = S;
            return KSSD_OK;
        }
    }

    // occurrence buffer: expected total/|sampling| ; 4x head-room, retried on overflow
    const double rate = (double)c->info.n_sampled / (double)(1ull << (4 * P.s));
    uint64_t cap = (uint64_t)((double)total * rate * 4.0) + (1u << 16);
    kssd_sketch *S = new kssd_sketch();
    S->ctx = c; S->n_genomes = n_genomes; S->n_comp = c->info.component_num; S->mode = mode;
    uint32_t n_occ = 0;
    bool fq_single_pass = !getenv("KSSD_FASTQ_TWO_PASS");
    for (int attempt = 0;; attempt++) {
        if (cap > 0xfffffff0ull) cap = 0xfffffff0ull;
        CU(c->keys.ensure(cap * 8));
        CU(c->ords.ensure(cap * 8));
        ScanArgs A{};
        A.seq = d_seq; A.seq_bytes = seq_bytes;
        A.goff = reinterpret_cast<uint64_t *>(pb);
        A.glen = reinterpret_cast<uint64_t *>(pb + p_glen);
        A.span_nom = reinterpret_cast<uint64_t *>(pb + p_nom);
        A.span_gid = reinterpret_cast<uint32_t *>(pb + p_sgid);
        A.n_spans = n_spans;
        A.gstatus = reinterpret_cast<int32_t *>(mb + m_stat);
        A.zero_count = reinterpret_cast<uint32_t *>(mb + m_zero);
        A.ticket = reinterpret_cast<uint32_t *>(mb + m_tick);
        A.out_count = reinterpret_cast<uint32_t *>(mb + m_cnt);
        A.out_keys = c->keys.as<uint64_t>(); A.out_ords = c->ords.as<uint64_t>();
        A.out_cap = (uint32_t)cap;
        A.drop_zero = (is_fastq || mode == KSSD_MODE_BYREAD) ? 0 : 1;      // only fasta2co's hash table loses code 0
        CU(cudaMemsetAsync(mb + m_zero, 0, m_end - m_zero, c->stream));
        CU(cudaEventRecord(c->ev[0], c->stream));
        if (!is_fastq) {
            if (n_spans) {
                const bool big = 2 * (P.TL - 1) >= 32;
                if (c->scan_impl == 2) sketch_fasta32_kernel<<<c->sm_count, kScanThreads, kScanSmemBytes, c->stream>>>(P, A);
                else if (c->scan_stride == 3 && big) sketch_fasta3_kernel<3, true><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
                else if (c->scan_stride == 3) sketch_fasta3_kernel<3, false><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
                else if (big) sketch_fasta3_kernel<1, true><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
                else sketch_fasta3_kernel<1, false><<<c->sm_count, kScanThreads, kScan3SmemBytes, c->stream>>>(P, A, c->d_prefilter3);
                LAUNCHED(1);
            }
        } else {
            int rc = fastq_run(c, d_seq, goff, glen, n_genomes, mode, opts ? opts->Q : 0, A, fq_single_pass);
            if (rc) { delete S; return rc; }
        }
        CU(cudaEventRecord(c->ev[1], c->stream));
        uint32_t tick_cnt[2] = {0, 0};                          // m_tick (FASTQ: "a line index overflowed") | m_cnt
        CU(cudaMemcpyAsync(tick_cnt, mb + m_tick, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        n_occ = tick_cnt[1];
        if (is_fastq && fq_single_pass && tick_cnt[0]) { fq_single_pass = false; attempt--; continue; }      // once more with the two-pass index
        if (n_occ <= cap) break;
        if (attempt >= 2) { delete S; return fail(KSSD_E_NOMEM, "kssd_sketch_batch: occurrence buffer overflow (%u > %llu)", n_occ, (unsigned long long)cap); }
        cap = (uint64_t)n_occ + (n_occ >> 3) + 1024;
    }
    CU(cudaEventElapsedTime(&S->scan_ms, c->ev[0], c->ev[1]));
    c->last_ms[0] = S->scan_ms;
    S->n_occ = n_occ;
    if (mode == KSSD_MODE_BYREAD) {
        const int rc = byread_finish(c, S, d_seq, goff, glen, n_genomes, n_occ, reinterpret_cast<const uint64_t *>(pb),
                                     reinterpret_cast<const int32_t *>(mb + m_stat));
        if (rc) { if (S->d_blob) cudaFreeAsync(S->d_blob, c->stream); delete S; return rc; }
        *out = S;
        return KSSD_OK;
    }

    const int n_comp = S->n_comp;
    S->status.assign(n_genomes, 0);
    S->comp_start.assign(n_comp + 1, 0);
    S->index.assign((size_t)n_comp * (n_genomes + 1), 0);
    // per_cg: kept ids per (component, genome) | kept total | distinct keys per genome (before the keep rule)
    const size_t cg_total = (size_t)n_comp * n_genomes, cg_distinct = cg_total + 1;
    std::vector<uint32_t> per_cg(cg_distinct + n_genomes, 0);
    std::vector<uint32_t> zeros(n_genomes, 0);
    // one stream-ordered allocation holds ids | ord | abund | index for the life of the handle
    const size_t nmax = std::max<size_t>(n_occ, 1);
    const size_t o_ids = 0, o_ord = (nmax * 4 + 15) & ~(size_t)15, o_ab = o_ord + nmax * 8, o_idx = (o_ab + nmax * 2 + 15) & ~(size_t)15;
    S->blob_bytes = o_idx + S->index.size() * 8;
    CU(cudaMallocAsync(&S->d_blob, S->blob_bytes, c->stream));
    S->d_ids = reinterpret_cast<uint32_t *>(S->d_blob + o_ids);
    S->d_ord = reinterpret_cast<uint64_t *>(S->d_blob + o_ord);
    S->d_abund = reinterpret_cast<uint16_t *>(S->d_blob + o_ab);
    S->d_index = reinterpret_cast<uint64_t *>(S->d_blob + o_idx);
    if (n_occ) {
        // sort occurrences by (component, genome, id); carry the offset along
        CU(c->keys2.ensure((size_t)n_occ * 8));
        CU(c->ords2.ensure((size_t)n_occ * 8));
        size_t tmp_bytes = 0;
        // sort only the bits in use: id (28) + genome id, plus the component field at bit 56 when there is one
        int gbits = 1;
        while ((1ll << gbits) < n_genomes) gbits++;
        const int end_bit = P.comp_code_bits ? 56 + P.comp_code_bits : 28 + gbits;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->keys.as<uint64_t>(), c->keys2.as<uint64_t>(), c->ords.as<uint64_t>(),
                                        c->ords2.as<uint64_t>(), n_occ, 0, end_bit, c->stream);
        CU(c->cubtmp.ensure(tmp_bytes));
        CU(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp_bytes, c->keys.as<uint64_t>(), c->keys2.as<uint64_t>(), c->ords.as<uint64_t>(),
                                           c->ords2.as<uint64_t>(), n_occ, 0, end_bit, c->stream));
        LAUNCHED(8);
        CU(c->flags.ensure((size_t)n_occ * 4));
        CU(c->pos.ensure((size_t)n_occ * 4));
        CU(c->runs.ensure((size_t)n_occ * 4));
        CU(c->keep.ensure((size_t)n_occ * 8));          // keep flags | their exclusive scan
        CU(c->counts.ensure((size_t)n_occ * 2));
        CU(c->minord.ensure((size_t)n_occ * 8));
        uint32_t *d_keep = c->keep.as<uint32_t>(), *d_outpos = d_keep + n_occ;
        const uint32_t nb = (n_occ + 255) / 256;
        size_t tmp2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp2, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n_occ, c->stream);
        CU(c->cubtmp.ensure(tmp2));
        CU(c->misc.ensure(per_cg.size() * 4));
        CU(cudaMemsetAsync(c->misc.p, 0, per_cg.size() * 4, c->stream));
        CU(cudaMemsetAsync(c->minord.p, 0xff, (size_t)n_occ * 8, c->stream));
        run_heads_kernel<<<nb, 256, 0, c->stream>>>(c->keys2.as<uint64_t>(), n_occ, c->flags.as<uint32_t>());
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp2, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n_occ, c->stream));
        run_reduce_kernel<<<nb, 256, 0, c->stream>>>(c->ords2.as<uint64_t>(), c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n_occ,
                                                     c->runs.as<uint32_t>(), c->minord.as<unsigned long long>());
        run_keep_kernel<<<nb, 256, 0, c->stream>>>(c->keys2.as<uint64_t>(), c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), c->runs.as<uint32_t>(),
                                                   n_occ, mode, opts ? opts->M : 1, d_keep, c->counts.as<uint16_t>(),
                                                   c->misc.as<uint32_t>() + cg_distinct);
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp2, d_keep, d_outpos, n_occ, c->stream));
        sketch_scatter_kernel<<<nb, 256, 0, c->stream>>>(c->keys2.as<uint64_t>(), c->flags.as<uint32_t>(), c->pos.as<uint32_t>(),
                                                         c->runs.as<uint32_t>(), d_keep, d_outpos, c->counts.as<uint16_t>(),
                                                         c->minord.as<unsigned long long>(), n_occ, n_genomes, S->d_ids, S->d_abund, S->d_ord,
                                                         c->misc.as<uint32_t>());
        sketch_total_kernel<<<1, 1, 0, c->stream>>>(d_keep, d_outpos, n_occ, c->misc.as<uint32_t>() + cg_total);
        LAUNCHED(9);
        CU(cudaMemcpyAsync(per_cg.data(), c->misc.p, per_cg.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(S->status.data(), mb + m_stat, 4ull * n_genomes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(zeros.data(), mb + m_zero, 4ull * n_genomes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    S->total = per_cg[cg_total];
    // per-component combco.index (command_dist.c:331-354) and reference error conditions
    std::vector<uint64_t> per_genome(n_genomes, 0);
    uint64_t run = 0;
    for (int cc = 0; cc < n_comp; cc++) {
        S->comp_start[cc] = run;
        uint64_t *ix = &S->index[(size_t)cc * (n_genomes + 1)];
        ix[0] = 0;
        for (int g = 0; g < n_genomes; g++) {
            const uint32_t v = per_cg[(size_t)cc * n_genomes + g];
            ix[g + 1] = ix[g] + v;
            per_genome[g] += v;
        }
        run += ix[n_genomes];
    }
    S->comp_start[n_comp] = run;
    for (int g = 0; g < n_genomes; g++) {
        if (S->status[g] & 1) S->status[g] = KSSD_E_HEADER_EOF;
        else if (S->status[g] & 2) S->status[g] = KSSD_E_LONGLINE;
        // "the context space is too crowd": the reference counts every slot it takes -- every distinct key, kept later or
        // not, and in the FASTA modes once more for every occurrence of code 0, whose slot never looks taken
        // (iseq2comem.c:258-263, :685-691; -A: :595-599); fastq2co never counts (:338)
        else if (mode != KSSD_MODE_FASTQ && (uint64_t)per_cg[cg_distinct + g] + zeros[g] > c->info.hashlimit) S->status[g] = KSSD_E_CROWD;
        else S->status[g] = 0;
    }
    CU(cudaMemcpyAsync(S->d_index, S->index.data(), S->index.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev[2], c->stream));
    CU(cudaEventSynchronize(c->ev[2]));
    CU(cudaEventElapsedTime(&c->last_ms[1], c->ev[0], c->ev[2]));
    c->total_ms_pending = false;
    *out = S;
    return KSSD_OK;
}

extern "C" int kssd_sketch_batch_dev(kssd_ctx_t *c, const uint8_t *seq_dev, size_t seq_bytes, const uint64_t *goff, const uint64_t *glen,
                                     int n_genomes, const kssd_sketch_opts_t *opts, kssd_sketch_t **out)
{
    if (!c || !seq_dev || !goff || !glen || !out) return fail(KSSD_E_INVAL, "kssd_sketch_batch_dev: null argument");
    if (reinterpret_cast<uintptr_t>(seq_dev) & 31) return fail(KSSD_E_INVAL, "kssd_sketch_batch_dev: seq_dev must be 32-byte aligned");
    CU(cudaSetDevice(c->device));
    return sketch_run(c, seq_dev, seq_bytes, goff, glen, n_genomes, opts, out);
}

extern "C" int kssd_sketch_batch_host(kssd_ctx_t *c, const uint8_t *seq, size_t seq_bytes, const uint64_t *goff, const uint64_t *glen,
                                      int n_genomes, const kssd_sketch_opts_t *opts, kssd_sketch_t **out)
{
    if (!c || !seq || !goff || !glen || !out) return fail(KSSD_E_INVAL, "kssd_sketch_batch_host: null argument");
    CU(cudaSetDevice(c->device));
    CU(c->seq.ensure(seq_bytes + 1024));
    CU(cudaMemcpyAsync(c->seq.p, seq, seq_bytes, cudaMemcpyHostToDevice, c->stream));
    return sketch_run(c, c->seq.as<uint8_t>(), seq_bytes, goff, glen, n_genomes, opts, out);
}

extern "C" int64_t kssd_sketch_count(const kssd_sketch_t *s, int comp)
{
    if (!s || comp < 0 || comp >= s->n_comp) return fail(KSSD_E_INVAL, "kssd_sketch_count: bad component");
    return (int64_t)(s->comp_start[comp + 1] - s->comp_start[comp]);
}

extern "C" int kssd_sketch_status(const kssd_sketch_t *s, int32_t *status_out)
{
    if (!s || !status_out) return fail(KSSD_E_INVAL, "kssd_sketch_status: null");
    memcpy(status_out, s->status.data(), s->status.size() * 4);
    return KSSD_OK;
}

extern "C" int kssd_sketch_fetch(const kssd_sketch_t *s, int comp, uint32_t *ids, uint64_t *index, uint16_t *abund, uint64_t *ord)
{
    if (!s || comp < 0 || comp >= s->n_comp) return fail(KSSD_E_INVAL, "kssd_sketch_fetch: bad component");
    kssd_ctx *c = s->ctx;
    CU(cudaSetDevice(c->device));
    const uint64_t b = s->comp_start[comp], n = s->comp_start[comp + 1] - b;
    if (ids && n) CU(cudaMemcpyAsync(ids, s->d_ids + b, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (abund && n) CU(cudaMemcpyAsync(abund, s->d_abund + b, n * 2, cudaMemcpyDeviceToHost, c->stream));
    if (ord && n) CU(cudaMemcpyAsync(ord, s->d_ord + b, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (index) memcpy(index, &s->index[(size_t)comp * (s->n_genomes + 1)], 8ull * (s->n_genomes + 1));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

extern "C" int kssd_sketch_read_counts(const kssd_sketch_t *s, uint64_t *n_reads_out)
{
    if (!s || !n_reads_out) return fail(KSSD_E_INVAL, "kssd_sketch_read_counts: null");
    if (s->mode != KSSD_MODE_BYREAD) return fail(KSSD_E_INVAL, "kssd_sketch_read_counts: not a --byread sketch");
    memcpy(n_reads_out, s->n_reads.data(), s->n_reads.size() * 8);
    return KSSD_OK;
}

extern "C" int kssd_sketch_fetch_read_index(const kssd_sketch_t *s, int comp, int file, uint64_t *index_out)
{
    if (!s || !index_out || comp < 0 || comp >= s->n_comp || file < 0 || file >= s->n_genomes)
        return fail(KSSD_E_INVAL, "kssd_sketch_fetch_read_index: bad argument");
    if (s->mode != KSSD_MODE_BYREAD) return fail(KSSD_E_INVAL, "kssd_sketch_fetch_read_index: not a --byread sketch");
    kssd_ctx *c = s->ctx;
    CU(cudaSetDevice(c->device));
    const uint64_t *ix = &s->index[(size_t)comp * (s->n_genomes + 1)];
    const uint64_t seg_lo = s->comp_start[comp] + ix[file], seg_n = ix[file + 1] - ix[file], nr = s->n_reads[file];
    CU(c->minord.ensure((nr + 1) * 8));
    byread_index_kernel<<<(uint32_t)((nr + 1 + 255) / 256), 256, 0, c->stream>>>(s->d_ord + seg_lo, seg_n, nr, c->minord.as<uint64_t>());
    LAUNCHED(1);
    CU(cudaMemcpyAsync(index_out, c->minord.p, (nr + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return KSSD_OK;
}

extern "C" int kssd_sketch_dev_ptrs(const kssd_sketch_t *s, int comp, const uint32_t **ids_dev, const uint64_t **index_dev)
{
    if (!s || comp < 0 || comp >= s->n_comp) return fail(KSSD_E_INVAL, "kssd_sketch_dev_ptrs: bad component");
    if (ids_dev) *ids_dev = s->d_ids + s->comp_start[comp];
    if (index_dev) *index_dev = s->d_index + (size_t)comp * (s->n_genomes + 1);
    return KSSD_OK;
}

extern "C" int kssd_sketch_stats(const kssd_sketch_t *s, uint64_t *n_occurrences, float *scan_kernel_ms)
{
    if (!s) return fail(KSSD_E_INVAL, "kssd_sketch_stats: null");
    if (n_occurrences) *n_occurrences = s->n_occ;
    if (scan_kernel_ms) *scan_kernel_ms = s->scan_ms;
    return KSSD_OK;
}

extern "C" void kssd_sketch_free(kssd_sketch_t *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->d_blob) cudaFreeAsync(s->d_blob, s->ctx->stream);
    delete s;
}

#include "stage1_files.cuh"

// ------------------------------------------------------------------------------------------------
// validation of caller / on-disk data that is about to be used as device indices (a stale or mismatched sketch, index
// or pan file must end in KSSD_E_INVAL, not in an out-of-bounds write): one streaming pass, one flag
// ------------------------------------------------------------------------------------------------
__global__ void validate_below_kernel(const uint32_t *__restrict__ v, uint64_t n, uint32_t limit, uint32_t *__restrict__ flag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < n && v[i] >= limit;
    if (__any_sync(kFull, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

static int validate_below(kssd_ctx *c, const uint32_t *d_v, uint64_t n, uint64_t limit)
{
    if (!n || limit > 0xffffffffull) return KSSD_OK;
    validate_below_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(d_v, n, (uint32_t)limit, c->d_flag);
    LAUNCHED(1);
    return KSSD_OK;
}

// reads and clears the flag (synchronises the stream)
static int validation_failed(kssd_ctx *c, bool *failed)
{
    uint32_t f = 0;
    CU(cudaMemcpyAsync(&f, c->d_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (f) CU(cudaMemsetAsync(c->d_flag, 0, 4, c->stream));
    *failed = f != 0;
    return KSSD_OK;
}

// largest per-query code count of one component, from the data (counts are bounded by it, whatever sizes the caller declared)
__global__ void max_extent_kernel(const uint64_t *__restrict__ qindex, uint32_t n_qry, uint32_t *__restrict__ out)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    if (q < n_qry) { const uint64_t e = qindex[q + 1] - qindex[q]; v = e > 0xffffffffull ? 0xffffffffu : (uint32_t)e; }
    v = __reduce_max_sync(kFull, v);
    if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

static int max_query_extent(kssd_ctx *c, const uint64_t *qindex_dev, int n_qry, uint32_t *out)
{
    CU(cudaMemsetAsync(c->d_flag + 1, 0, 4, c->stream));
    max_extent_kernel<<<(n_qry + 255) / 256, 256, 0, c->stream>>>(qindex_dev, (uint32_t)n_qry, c->d_flag + 1);
    LAUNCHED(1);
    CU(cudaMemcpyAsync(out, c->d_flag + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

// ------------------------------------------------------------------------------------------------
// Stage II
// ------------------------------------------------------------------------------------------------
struct kssd_index {
    kssd_ctx *ctx = nullptr;
    int n_genomes = 0;
    uint64_t n_postings = 0, n_unique = 0, space = 0;
    uint32_t *d_ucodes = nullptr, *d_uoff = nullptr, *d_gids = nullptr;
    uint2 *d_rb = nullptr;                       // rank bitmap over the code space (index_dist.cuh)
    CodeLookup lookup() const { return CodeLookup{d_rb, d_uoff, (uint32_t)(space >> 5)}; }
};

static int index_finish(kssd_ctx *c, kssd_index *ix, uint32_t *d_sorted_codes)
{
    // d_sorted_codes: n_postings codes ascending (scratch), ix->d_gids already in (code, gid) order
    const uint64_t n = ix->n_postings;
    uint32_t nuniq = 0;
    size_t rle_tmp = 0;
    const bool big = n >= 0x7fffffffull;
    if (n && !big) {
        // unique codes and their run lengths in ONE pass over the sorted codes (run-length encode with a look-back scan inside), into
        // scratch that can hold the worst case; the exact-size arrays are filled below
        CU(c->flags.ensure(n * 4));                           // unique codes (scratch)
        CU(c->pos.ensure(n * 4));                             // run lengths (scratch)
        CU(c->misc.ensure(16));
        cub::DeviceRunLengthEncode::Encode(nullptr, rle_tmp, d_sorted_codes, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), c->misc.as<uint32_t>(), (int)n, c->stream);
        CU(c->cubtmp.ensure(rle_tmp));
        CU(cub::DeviceRunLengthEncode::Encode(c->cubtmp.p, rle_tmp, d_sorted_codes, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), c->misc.as<uint32_t>(), (int)n,
                                              c->stream));
        LAUNCHED(2);
        CU(cudaMemcpyAsync(&nuniq, c->misc.p, 4, cudaMemcpyDeviceToHost, c->stream));
        bool bad = false;
        const int vrc = validation_failed(c, &bad);               // (synchronises)
        if (vrc) return vrc;
        if (bad) return fail(KSSD_E_INVAL, "index: a code or genome id lies outside the component's space / the genome count (mismatched or corrupt input)");
    }
    if (big) {                                                // head flags + scan + scatter (three passes)
        CU(c->flags.ensure(n * 4));
        CU(c->pos.ensure(n * 4));
        const uint32_t nb = (uint32_t)((n + 255) / 256);
        head_flags_kernel<<<nb, 256, 0, c->stream>>>(d_sorted_codes, n, c->flags.as<uint32_t>());
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n, c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n, c->stream));
        LAUNCHED(3);
        uint32_t lp = 0, lf = 0;
        CU(cudaMemcpyAsync(&lp, c->pos.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&lf, c->flags.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        bool bad = false;
        const int vrc = validation_failed(c, &bad);               // (synchronises)
        if (vrc) return vrc;
        if (bad) return fail(KSSD_E_INVAL, "index: a code or genome id lies outside the component's space / the genome count (mismatched or corrupt input)");
        nuniq = lp + lf;
    }
    ix->n_unique = nuniq;
    CU(cudaMallocAsync(&ix->d_ucodes, std::max<size_t>(nuniq, 1) * 4, c->stream));
    CU(cudaMallocAsync(&ix->d_uoff, ((size_t)nuniq + 1) * 4, c->stream));
    if (big) {
        csr_scatter_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, c->stream>>>(d_sorted_codes, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n, ix->d_ucodes, ix->d_uoff);
        LAUNCHED(1);
    } else if (nuniq) {
        CU(cudaMemcpyAsync(ix->d_ucodes, c->flags.p, (size_t)nuniq * 4, cudaMemcpyDeviceToDevice, c->stream));
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->pos.as<uint32_t>(), ix->d_uoff, nuniq, c->stream);
        CU(c->cubtmp.ensure(std::max(tmp, rle_tmp)));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->pos.as<uint32_t>(), ix->d_uoff, nuniq, c->stream));
        LAUNCHED(2);
    }
    const uint32_t n32 = (uint32_t)n;
    CU(cudaMemcpyAsync(ix->d_uoff + nuniq, &n32, 4, cudaMemcpyHostToDevice, c->stream));
    // rank bitmap for O(1) query lookups
    CU(cudaMallocAsync(&ix->d_rb, (ix->space >> 5) * sizeof(uint2), c->stream));
    CU(cudaMemsetAsync(ix->d_rb, 0, (ix->space >> 5) * sizeof(uint2), c->stream));
    if (nuniq) {
        rb_build_kernel<<<(nuniq + 255) / 256, 256, 0, c->stream>>>(ix->d_ucodes, nuniq, ix->d_rb);
        LAUNCHED(1);
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return KSSD_OK;
}

extern "C" int kssd_index_build_dev(kssd_ctx_t *c, const uint32_t *combco_dev, const uint64_t *cbdcoindex_dev, int n_genomes, uint64_t n_codes,
                                    kssd_index_t **out)
{
    if (!c || !out || n_genomes <= 0 || (n_codes && (!combco_dev || !cbdcoindex_dev))) return fail(KSSD_E_INVAL, "kssd_index_build_dev: bad argument");
    if (n_codes >= 0xffffffffull) return fail(KSSD_E_INVAL, "kssd_index_build_dev: more than 2^32 postings in one component");
    CU(cudaSetDevice(c->device));
    kssd_index *ix = new kssd_index();
    ix->ctx = c; ix->n_genomes = n_genomes; ix->n_postings = n_codes;
    ix->space = 1ull << (4 * c->info.component_sz);   // the reference's dense table always spans 16^COMPONENT_SZ
    CU(cudaEventRecord(c->ev[0], c->stream));
    CU(cudaMallocAsync(&ix->d_gids, std::max<size_t>(n_codes, 1) * 4, c->stream));
    uint32_t *d_sorted = nullptr;
    if (n_codes) {
        CU(c->keys.ensure(n_codes * 4));     // gid tags (unsorted)
        CU(c->keys2.ensure(n_codes * 4));    // sorted codes
        // gid of every posting position + the check that no code lies outside 16^COMPONENT_SZ, one pass
        tag_gids_kernel<<<(uint32_t)((((n_codes + 1023) >> 10) * 32 + 255) / 256), 256, 0, c->stream>>>(cbdcoindex_dev, n_genomes, n_codes, combco_dev, ix->space,
                                                                                                      c->keys.as<uint32_t>(), c->d_flag);
        LAUNCHED(1);
        size_t tmp = 0;
        tmp = 0;
        const int bits = 4 * std::min(c->info.component_sz, c->info.k - c->info.drlevel);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, combco_dev, c->keys2.as<uint32_t>(), c->keys.as<uint32_t>(), ix->d_gids, n_codes, 0, bits,
                                        c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp, combco_dev, c->keys2.as<uint32_t>(), c->keys.as<uint32_t>(), ix->d_gids, n_codes, 0,
                                           bits, c->stream));
        LAUNCHED(5);
        d_sorted = c->keys2.as<uint32_t>();
    }
    int rc = index_finish(c, ix, d_sorted);
    if (rc) { kssd_index_free(ix); return rc; }
    CU(cudaEventRecord(c->ev[1], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventElapsedTime(&c->last_ms[2], c->ev[0], c->ev[1]));
    *out = ix;
    return KSSD_OK;
}

extern "C" int kssd_index_build_host(kssd_ctx_t *c, const uint32_t *combco, const uint64_t *cbdcoindex, int n_genomes, kssd_index_t **out)
{
    if (!c || !cbdcoindex || !out || n_genomes <= 0) return fail(KSSD_E_INVAL, "kssd_index_build_host: bad argument");
    CU(cudaSetDevice(c->device));
    const uint64_t n = cbdcoindex[n_genomes];
    if (n && !combco) return fail(KSSD_E_INVAL, "kssd_index_build_host: null combco");
    CU(c->seq.ensure(n * 4 + 8ull * (n_genomes + 1) + 64));
    uint64_t *d_index = c->seq.as<uint64_t>();
    uint32_t *d_codes = reinterpret_cast<uint32_t *>(d_index + n_genomes + 1);
    CU(cudaMemcpyAsync(d_index, cbdcoindex, 8ull * (n_genomes + 1), cudaMemcpyHostToDevice, c->stream));
    if (n) CU(cudaMemcpyAsync(d_codes, combco, n * 4, cudaMemcpyHostToDevice, c->stream));
    return kssd_index_build_dev(c, d_codes, d_index, n_genomes, n, out);
}

extern "C" int kssd_index_sizes(const kssd_index_t *ix, uint64_t *n_unique, uint64_t *n_postings, int *n_genomes)
{
    if (!ix) return fail(KSSD_E_INVAL, "kssd_index_sizes: null");
    if (n_unique) *n_unique = ix->n_unique;
    if (n_postings) *n_postings = ix->n_postings;
    if (n_genomes) *n_genomes = ix->n_genomes;
    return KSSD_OK;
}

extern "C" int kssd_index_fetch(const kssd_index_t *ix, uint32_t *ucodes, uint64_t *uoff, uint32_t *gids)
{
    if (!ix) return fail(KSSD_E_INVAL, "kssd_index_fetch: null");
    kssd_ctx *c = ix->ctx;
    CU(cudaSetDevice(c->device));
    if (ucodes && ix->n_unique) CU(cudaMemcpyAsync(ucodes, ix->d_ucodes, ix->n_unique * 4, cudaMemcpyDeviceToHost, c->stream));
    if (gids && ix->n_postings) CU(cudaMemcpyAsync(gids, ix->d_gids, ix->n_postings * 4, cudaMemcpyDeviceToHost, c->stream));
    std::vector<uint32_t> tmp;
    if (uoff) {
        tmp.resize(ix->n_unique + 1);
        CU(cudaMemcpyAsync(tmp.data(), ix->d_uoff, (ix->n_unique + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    if (uoff) for (size_t i = 0; i <= ix->n_unique; i++) uoff[i] = tmp[i];
    return KSSD_OK;
}

extern "C" int kssd_index_fetch_dense(const kssd_index_t *ix, uint64_t *dense_out)
{
    if (!ix || !dense_out) return fail(KSSD_E_INVAL, "kssd_index_fetch_dense: null");
    kssd_ctx *c = ix->ctx;
    CU(cudaSetDevice(c->device));
    // the reference's dense table is a file format: built here for the export only (zero, mark list ends, max-scan)
    uint32_t *d_dense = nullptr;
    CU(cudaMallocAsync(&d_dense, (ix->space + 1) * 4, c->stream));
    CU(cudaMemsetAsync(d_dense, 0, (ix->space + 1) * 4, c->stream));
    if (ix->n_unique) {
        dense_mark_kernel<<<(uint32_t)((ix->n_unique + 255) / 256), 256, 0, c->stream>>>(ix->d_ucodes, ix->d_uoff, (uint32_t)ix->n_unique, d_dense);
        size_t tmpb = 0;
        cub::DeviceScan::InclusiveScan(nullptr, tmpb, d_dense, d_dense, cub::Max(), ix->space + 1, c->stream);
        if (c->cubtmp.ensure(tmpb) != cudaSuccess) { cudaFreeAsync(d_dense, c->stream); return fail(KSSD_E_CUDA, "kssd_index_fetch_dense: out of memory"); }
        cub::DeviceScan::InclusiveScan(c->cubtmp.p, tmpb, d_dense, d_dense, cub::Max(), ix->space + 1, c->stream);
        LAUNCHED(3);
    }
    const uint64_t chunk = 1ull << 24;   // 128 MiB of u64 per step
    if (c->keys.ensure(chunk * 8) != cudaSuccess) { cudaFreeAsync(d_dense, c->stream); return fail(KSSD_E_CUDA, "kssd_index_fetch_dense: out of memory"); }
    for (uint64_t first = 0; first < ix->space; first += chunk) {
        const uint64_t cnt = std::min(chunk, ix->space - first);
        dense_incl64_kernel<<<(uint32_t)((cnt + 255) / 256), 256, 0, c->stream>>>(d_dense, first, cnt, c->keys.as<uint64_t>());
        LAUNCHED(1);
        cudaMemcpyAsync(dense_out + first, c->keys.p, cnt * 8, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    cudaFreeAsync(d_dense, c->stream);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return KSSD_OK;
}

__global__ void codes_from_dense_kernel(const uint32_t *__restrict__ dense, uint64_t space, uint32_t *__restrict__ codes)
{
    // every code writes itself over its posting range
    const uint64_t cidx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= space) return;
    const uint32_t s = dense[cidx], e = dense[cidx + 1];
    for (uint32_t g = s; g < e; g++) codes[g] = (uint32_t)cidx;
}

extern "C" int kssd_index_from_dense_host(kssd_ctx_t *c, const uint64_t *dense_incl, const uint32_t *gids, uint64_t n_postings, int n_genomes,
                                          kssd_index_t **out)
{
    if (!c || !dense_incl || !out || n_genomes <= 0 || (n_postings && !gids)) return fail(KSSD_E_INVAL, "kssd_index_from_dense_host: bad argument");
    if (n_postings >= 0xffffffffull) return fail(KSSD_E_INVAL, "kssd_index_from_dense_host: too many postings");
    CU(cudaSetDevice(c->device));
    kssd_index *ix = new kssd_index();
    ix->ctx = c; ix->n_genomes = n_genomes; ix->n_postings = n_postings;
    ix->space = 1ull << (4 * c->info.component_sz);   // the reference's dense table always spans 16^COMPONENT_SZ
    CU(cudaMallocAsync(&ix->d_gids, std::max<size_t>(n_postings, 1) * 4, c->stream));
    if (n_postings) CU(cudaMemcpyAsync(ix->d_gids, gids, n_postings * 4, cudaMemcpyHostToDevice, c->stream));
    validate_below(c, ix->d_gids, n_postings, (uint64_t)n_genomes);
    if (dense_incl[ix->space - 1] != n_postings) {
        kssd_index_free(ix);
        return fail(KSSD_E_MISMATCH, "kssd_index_from_dense_host: mco.index ends at %llu, mco holds %llu postings", (unsigned long long)dense_incl[ix->space - 1],
                    (unsigned long long)n_postings);
    }
    uint32_t *d_dense_tmp = nullptr;
    CU(cudaMallocAsync(&d_dense_tmp, (ix->space + 1) * 4, c->stream));
    const uint64_t chunk = 1ull << 24;
    CU(c->keys.ensure(chunk * 8));
    for (uint64_t first = 0; first < ix->space; first += chunk) {
        const uint64_t cnt = std::min(chunk, ix->space - first);
        CU(cudaMemcpyAsync(c->keys.p, dense_incl + first, cnt * 8, cudaMemcpyHostToDevice, c->stream));
        dense_from_incl64_kernel<<<(uint32_t)((cnt + 255) / 256), 256, 0, c->stream>>>(c->keys.as<uint64_t>(), first, cnt, first ? dense_incl[first - 1] : 0ull,
                                                                                       n_postings, d_dense_tmp, c->d_flag);
        LAUNCHED(1);
        CU(cudaStreamSynchronize(c->stream));
    }
    CU(c->keys2.ensure(std::max<size_t>(n_postings, 1) * 4));
    codes_from_dense_kernel<<<(uint32_t)((ix->space + 255) / 256), 256, 0, c->stream>>>(d_dense_tmp, ix->space, c->keys2.as<uint32_t>());
    LAUNCHED(1);
    int rc = index_finish(c, ix, n_postings ? c->keys2.as<uint32_t>() : nullptr);
    cudaFreeAsync(d_dense_tmp, c->stream);
    if (rc) { kssd_index_free(ix); return rc; }
    *out = ix;
    return KSSD_OK;
}

extern "C" void kssd_index_free(kssd_index_t *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    cudaStream_t st = ix->ctx->stream;
    if (ix->d_ucodes) cudaFreeAsync(ix->d_ucodes, st);
    if (ix->d_uoff) cudaFreeAsync(ix->d_uoff, st);
    if (ix->d_gids) cudaFreeAsync(ix->d_gids, st);
    if (ix->d_rb) cudaFreeAsync(ix->d_rb, st);
    delete ix;
}

// ------------------------------------------------------------------------------------------------
// Stage III
// ------------------------------------------------------------------------------------------------
struct kssd_dist {
    kssd_ctx *ctx = nullptr;
    int n_qry = 0, n_ref = 0, components_done = 0;
    uint32_t max_qry_size = 0;
    uint64_t max_count = 0;                      // bound of any cell from the DATA: sum over components of the largest query extent
    bool empty_qry = false, empty_ref = false;   // some sketch has no k-mer (its cells give NaN statistics)
    uint32_t *d_ct = nullptr, *d_qsz = nullptr, *d_rsz = nullptr;
    bool owns_ct = true;
    StatRow *d_rows = nullptr;
    uint64_t n_rows = 0;
    // sparse job (kssd_dist_create_sparse): no Q x R matrix; the components are registered and counted inside kssd_dist_stats
    bool sparse = false;
    std::vector<SparseComp> comps;
    std::vector<const kssd_index_t *> comp_ix;
    std::vector<uint64_t> comp_ncodes;
    std::vector<void *> owned;                   // device copies of host query sketches
    std::vector<uint32_t> h_qsz;                 // host copies of the sketch sizes (sub-jobs of the sparse path)
    std::shared_ptr<std::vector<uint32_t>> h_rsz;   // shared with the context's cache of reference sets (d_rsz is its device copy)
    int async_slot = -1;                         // pinned slot + events of kssd_dist_stats_async (from the context's pool)
    uint8_t *d_async = nullptr;                  // its per-job scratch: hit list, per-query tallies (several searches are in flight)
    cudaEvent_t counted = nullptr;
    // kssd_dist_stats_async: the search is in flight; its outcome lands in pinned memory behind `done`
    bool async_pending = false;
    uint64_t *h_async = nullptr;                 // pinned: total hits | n_over (low 32) , bad extent (high 32)
    cudaEvent_t done = nullptr;
    uint64_t async_cap = 0;
    kssd_stat_opts_t async_opts{};
};

static int dist_create(kssd_ctx_t *c, int n_qry, int n_ref, const uint32_t *qry_ctx_ct, const uint32_t *ref_ctx_ct, uint32_t *ct_ext, int filled,
                       kssd_dist_t **out)
{
    if (!c || !out || n_qry <= 0 || n_ref <= 0 || !qry_ctx_ct || !ref_ctx_ct) return fail(KSSD_E_INVAL, "kssd_dist_create: bad argument");
    CU(cudaSetDevice(c->device));
    kssd_dist *d = new kssd_dist();
    d->ctx = c; d->n_qry = n_qry; d->n_ref = n_ref;
    for (int i = 0; i < n_qry; i++) { d->max_qry_size = std::max(d->max_qry_size, qry_ctx_ct[i]); d->empty_qry |= qry_ctx_ct[i] == 0; }
    d->h_qsz.assign(qry_ctx_ct, qry_ctx_ct + n_qry);
    // the reference sizes: on the device already if an earlier job of this context searched the same reference set
    {
        kssd_ctx::SizeSet *hit = nullptr;
        for (auto &e : c->size_sets)                                   // (exact: one memcmp of the array, ~30 us per 100 000 references)
            if (e.refs == (uint32_t)n_ref && memcmp(e.host->data(), ref_ctx_ct, (size_t)n_ref * 4) == 0) { hit = &e; break; }
        if (!hit) {
            for (size_t i = 0; i < c->size_sets.size() && c->size_sets.size() >= 4;)      // keep a few; never drop one a job still uses
                if (c->size_sets[i].host.use_count() == 1) { cudaFreeAsync(c->size_sets[i].dev, c->stream); c->size_sets.erase(c->size_sets.begin() + i); }
                else i++;
            kssd_ctx::SizeSet e;
            e.refs = (uint32_t)n_ref;
            for (int i = 0; i < n_ref; i++) e.any_empty |= ref_ctx_ct[i] == 0;
            e.host = std::make_shared<std::vector<uint32_t>>(ref_ctx_ct, ref_ctx_ct + n_ref);
            if (cudaMallocAsync(&e.dev, (size_t)n_ref * 4, c->stream) != cudaSuccess) { delete d; return fail(KSSD_E_CUDA, "kssd_dist_create: out of device memory"); }
            CU(cudaMemcpyAsync(e.dev, e.host->data(), (size_t)n_ref * 4, cudaMemcpyHostToDevice, c->stream));
            c->size_sets.push_back(std::move(e));
            hit = &c->size_sets.back();
        }
        d->h_rsz = hit->host;
        d->d_rsz = hit->dev;
        d->empty_ref = hit->any_empty;
    }
    if (ct_ext) { d->d_ct = ct_ext; d->owns_ct = false; d->components_done = filled ? 1 : 0; }
    else if (filled < 0) d->sparse = true;       // sparse job: the matrix is allocated only if a query overflows the sparse path
    else CU(cudaMallocAsync(&d->d_ct, (size_t)n_qry * n_ref * 4, c->stream));
    CU(cudaMallocAsync(&d->d_qsz, (size_t)n_qry * 4, c->stream));
    CU(cudaMemcpyAsync(d->d_qsz, d->h_qsz.data(), (size_t)n_qry * 4, cudaMemcpyHostToDevice, c->stream));     // (the job's own copy: no wait)
    *out = d;
    return KSSD_OK;
}

extern "C" int kssd_dist_create(kssd_ctx_t *c, int n_qry, int n_ref, const uint32_t *qry_ctx_ct, const uint32_t *ref_ctx_ct, kssd_dist_t **out)
{
    return dist_create(c, n_qry, n_ref, qry_ctx_ct, ref_ctx_ct, nullptr, 0, out);
}

extern "C" int kssd_dist_create_ext(kssd_ctx_t *c, int n_qry, int n_ref, const uint32_t *qry_ctx_ct, const uint32_t *ref_ctx_ct,
                                    uint32_t *ct_dev, int already_filled, kssd_dist_t **out)
{
    if (!ct_dev) return fail(KSSD_E_INVAL, "kssd_dist_create_ext: null count buffer");
    return dist_create(c, n_qry, n_ref, qry_ctx_ct, ref_ctx_ct, ct_dev, already_filled, out);
}

extern "C" int kssd_dist_create_sparse(kssd_ctx_t *c, int n_qry, int n_ref, const uint32_t *qry_ctx_ct, const uint32_t *ref_ctx_ct, kssd_dist_t **out)
{
    return dist_create(c, n_qry, n_ref, qry_ctx_ct, ref_ctx_ct, nullptr, -1, out);
}

extern "C" int kssd_dist_sparse_add_dev(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes_dev, const uint64_t *qindex_dev,
                                        uint64_t n_qcodes)
{
    if (!d || !ref_ix || !qindex_dev || (n_qcodes && !qcodes_dev)) return fail(KSSD_E_INVAL, "kssd_dist_sparse_add_dev: null argument");
    if (!d->sparse) return fail(KSSD_E_INVAL, "kssd_dist_sparse_add_dev: not a sparse job");
    if (ref_ix->n_genomes != d->n_ref) return fail(KSSD_E_MISMATCH, "query args not match ref args: index has %d genomes, job has %d", ref_ix->n_genomes, d->n_ref);
    if (d->comps.size() >= 256) return fail(KSSD_E_INVAL, "kssd_dist_sparse_add_dev: more than 256 components");
    d->comps.push_back(SparseComp{qcodes_dev, qindex_dev, ref_ix->lookup(), ref_ix->d_gids});
    d->comp_ix.push_back(ref_ix);
    d->comp_ncodes.push_back(n_qcodes);
    return KSSD_OK;
}

extern "C" int kssd_dist_sparse_add_host(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes, const uint64_t *qindex)
{
    if (!d || !ref_ix || !qindex) return fail(KSSD_E_INVAL, "kssd_dist_sparse_add_host: null argument");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    const uint64_t n = qindex[d->n_qry];
    if (n && !qcodes) return fail(KSSD_E_INVAL, "kssd_dist_sparse_add_host: null qcodes");
    uint8_t *buf = nullptr;
    const size_t ib = 8ull * (d->n_qry + 1);
    CU(cudaMallocAsync(&buf, ib + n * 4 + 16, c->stream));
    d->owned.push_back(buf);
    CU(cudaMemcpyAsync(buf, qindex, ib, cudaMemcpyHostToDevice, c->stream));
    if (n) CU(cudaMemcpyAsync(buf + ib, qcodes, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return kssd_dist_sparse_add_dev(d, ref_ix, reinterpret_cast<const uint32_t *>(buf + ib), reinterpret_cast<const uint64_t *>(buf), n);
}

// sparse job -> ordinary job: allocate the matrix and count every registered component into it
static int dist_densify(kssd_dist *d)
{
    kssd_ctx *c = d->ctx;
    CU(cudaMallocAsync(&d->d_ct, (size_t)d->n_qry * d->n_ref * 4, c->stream));
    d->owns_ct = true;
    d->sparse = false;
    d->components_done = 0;
    for (size_t i = 0; i < d->comps.size(); i++) {
        const int rc = kssd_dist_accumulate_dev(d, d->comp_ix[i], d->comps[i].qcodes, d->comps[i].qindex, d->comp_ncodes[i]);
        if (rc) return rc;
    }
    return KSSD_OK;
}

extern "C" int kssd_dist_accumulate_dev(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes_dev, const uint64_t *qindex_dev,
                                        uint64_t n_qcodes)
{
    if (!d || !ref_ix || !qindex_dev || (n_qcodes && !qcodes_dev)) return fail(KSSD_E_INVAL, "kssd_dist_accumulate_dev: null argument");
    if (d->sparse) return fail(KSSD_E_INVAL, "kssd_dist_accumulate_dev: sparse job, register components with kssd_dist_sparse_add_*");
    if (ref_ix->n_genomes != d->n_ref) return fail(KSSD_E_MISMATCH, "query args not match ref args: index has %d genomes, job has %d", ref_ix->n_genomes, d->n_ref);
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->ev[0], c->stream));
    const char *force = getenv("KSSD_DIST_KERNEL");           // "strip" / "rows": A/B switch for profiling
    const bool rows_fit_l2 = (uint64_t)d->n_ref * 4 * c->sm_count <= (128ull << 20);   // measured fine up to 4 rows per SM in flight
    const bool use_rows = force ? (strcmp(force, "rows") == 0) : rows_fit_l2;
    if (use_rows) {
        const char *rps = getenv("KSSD_DIST_ROWS_PER_SM");
        const uint32_t grid = (uint32_t)std::min<int>(d->n_qry, c->sm_count * (rps ? atoi(rps) : 4));
        dist_count_rows_kernel<<<grid, kDistRowThreads, 0, c->stream>>>(qcodes_dev, qindex_dev, ref_ix->lookup(), ref_ix->d_gids, (uint32_t)d->n_qry,
                                                                       (uint32_t)d->n_ref, d->d_ct, d->components_done > 0);
    } else {
        // 16-bit strip counters only if the DATA cannot overflow them (the declared sketch sizes may be per shard)
        uint32_t ext = 0;
        { const int rc = max_query_extent(c, qindex_dev, d->n_qry, &ext); if (rc) return rc; }
        const bool small = d->max_qry_size < 65536u && ext < 65536u;
        const uint32_t elem = small ? 2 : 4;
        // strip width: whole row when it fits in 112 KiB (two CTAs per SM), else equal tiles
        const char *skb = getenv("KSSD_DIST_STRIP_KB");
        const uint32_t max_refs = ((skb ? (uint32_t)atoi(skb) : 112u) << 10) / elem;
        const uint32_t n_tiles = ((uint32_t)d->n_ref + max_refs - 1) / max_refs;
        uint32_t tile = ((uint32_t)d->n_ref + n_tiles - 1) / n_tiles;
        tile = (tile + 7) & ~7u;      // strips start 16-byte aligned in the output row
        const size_t smem = ((size_t)tile * elem + 15) & ~(size_t)15;
        const uint32_t grid = (uint32_t)d->n_qry * n_tiles;
        if (small) {
            CU(cudaFuncSetAttribute(dist_count_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dist_count_kernel<uint16_t><<<grid, kDistThreads, smem, c->stream>>>(qcodes_dev, qindex_dev, ref_ix->lookup(), ref_ix->d_gids, (uint32_t)d->n_ref,
                                                                               tile, n_tiles, d->d_ct, d->components_done > 0);
        } else {
            CU(cudaFuncSetAttribute(dist_count_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dist_count_kernel<uint32_t><<<grid, kDistThreads, smem, c->stream>>>(qcodes_dev, qindex_dev, ref_ix->lookup(), ref_ix->d_gids, (uint32_t)d->n_ref,
                                                                               tile, n_tiles, d->d_ct, d->components_done > 0);
        }
    }
    LAUNCHED(1);
    CU(cudaEventRecord(c->ev[1], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&c->last_ms[3], c->ev[0], c->ev[1]));
    d->components_done++;
    return KSSD_OK;
}

extern "C" int kssd_dist_accumulate_host(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes, const uint64_t *qindex)
{
    if (!d || !ref_ix || !qindex) return fail(KSSD_E_INVAL, "kssd_dist_accumulate_host: null argument");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    const uint64_t n = qindex[d->n_qry];
    if (n && !qcodes) return fail(KSSD_E_INVAL, "kssd_dist_accumulate_host: null qcodes");
    CU(c->seq.ensure(n * 4 + 8ull * (d->n_qry + 1) + 64));
    uint64_t *d_index = c->seq.as<uint64_t>();
    uint32_t *d_codes = reinterpret_cast<uint32_t *>(d_index + d->n_qry + 1);
    CU(cudaMemcpyAsync(d_index, qindex, 8ull * (d->n_qry + 1), cudaMemcpyHostToDevice, c->stream));
    if (n) CU(cudaMemcpyAsync(d_codes, qcodes, n * 4, cudaMemcpyHostToDevice, c->stream));
    return kssd_dist_accumulate_dev(d, ref_ix, d_codes, d_index, n);
}

extern "C" int kssd_dist_accumulate_peer(kssd_ctx_t *c, const kssd_index_t *ref_ix, const uint32_t *qcodes_dev, const uint64_t *qindex_dev, int n_qry,
                                         int n_ref, uint32_t *const *row_blocks, int rows_per_block, int world)
{
    if (!c || !ref_ix || !qindex_dev || !row_blocks || n_qry <= 0 || n_ref <= 0 || rows_per_block <= 0 || world <= 0 || world > kMaxPeers)
        return fail(KSSD_E_INVAL, "kssd_dist_accumulate_peer: bad argument");
    if (ref_ix->n_genomes != n_ref) return fail(KSSD_E_MISMATCH, "query args not match ref args: index has %d genomes, job has %d", ref_ix->n_genomes, n_ref);
    if ((int64_t)rows_per_block * world < n_qry) return fail(KSSD_E_INVAL, "kssd_dist_accumulate_peer: row blocks do not cover the queries");
    CU(cudaSetDevice(c->device));
    PeerRows pr{};
    for (int i = 0; i < world; i++) pr.block[i] = row_blocks[i];
    CU(cudaEventRecord(c->ev[0], c->stream));
    const uint32_t grid = (uint32_t)std::min<int>(n_qry, c->sm_count * 4);
    dist_count_peer_kernel<<<grid, kDistRowThreads, 0, c->stream>>>(qcodes_dev, qindex_dev, ref_ix->lookup(), ref_ix->d_gids, (uint32_t)n_qry, (uint32_t)n_ref,
                                                                   pr, (uint32_t)rows_per_block);
    LAUNCHED(1);
    CU(cudaEventRecord(c->ev[1], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&c->last_ms[3], c->ev[0], c->ev[1]));
    return KSSD_OK;
}

extern "C" int kssd_dev_alloc(kssd_ctx_t *c, size_t bytes, void **ptr)
{
    if (!c || !ptr || !bytes) return fail(KSSD_E_INVAL, "kssd_dev_alloc: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(ptr, bytes));
    CU(cudaMemsetAsync(*ptr, 0, bytes, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

extern "C" int kssd_dev_zero(kssd_ctx_t *c, void *ptr, size_t bytes)
{
    if (!c || !ptr) return fail(KSSD_E_INVAL, "kssd_dev_zero: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(ptr, 0, bytes, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

extern "C" void kssd_dev_free(kssd_ctx_t *c, void *ptr)
{
    if (!c || !ptr) return;
    cudaSetDevice(c->device);
    cudaFree(ptr);
}

extern "C" int kssd_ipc_export(kssd_ctx_t *c, void *ptr, uint8_t handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!c || !ptr || !handle) return fail(KSSD_E_INVAL, "kssd_ipc_export: bad argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return KSSD_OK;
}

extern "C" int kssd_ipc_open(kssd_ctx_t *c, const uint8_t handle[64], void **ptr)
{
    if (!c || !ptr || !handle) return fail(KSSD_E_INVAL, "kssd_ipc_open: bad argument");
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KSSD_OK;
}

extern "C" int kssd_ipc_close(kssd_ctx_t *c, void *ptr)
{
    if (!c || !ptr) return fail(KSSD_E_INVAL, "kssd_ipc_close: bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaIpcCloseMemHandle(ptr));
    return KSSD_OK;
}

extern "C" int kssd_dist_fetch_counts(const kssd_dist_t *d, uint32_t *ct_out)
{
    if (!d || !ct_out) return fail(KSSD_E_INVAL, "kssd_dist_fetch_counts: null");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    if (d->sparse) {                              // the matrix was never built: build it now
        const int rc = dist_densify(const_cast<kssd_dist *>(d));
        if (rc) return rc;
    }
    if (d->components_done == 0) CU(cudaMemsetAsync(d->d_ct, 0, (size_t)d->n_qry * d->n_ref * 4, c->stream));
    CU(cudaMemcpyAsync(ct_out, d->d_ct, (size_t)d->n_qry * d->n_ref * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

extern "C" const uint32_t *kssd_dist_counts_dev(const kssd_dist_t *d) { return d ? d->d_ct : nullptr; }

// -N: the rows topn_*_kernel listed (row counts in `counts`, up to N per query at q * N in tmp_rows) -> d->d_rows, query-major
static int topn_collect(kssd_ctx *c, kssd_dist *d, const StatRow *tmp_rows, int N, uint32_t *counts /* n_qry + 1 words */, uint64_t *total_out)
{
    CU(c->pos.ensure(((size_t)d->n_qry + 1) * 8));
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, counts, c->pos.as<uint64_t>(), d->n_qry + 1, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cudaMemsetAsync(counts + d->n_qry, 0, 4, c->stream));
    CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, counts, c->pos.as<uint64_t>(), d->n_qry + 1, c->stream));
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, c->pos.as<uint64_t>() + d->n_qry, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMallocAsync(&d->d_rows, std::max<uint64_t>(total, 1) * sizeof(StatRow), c->stream));
    topn_gather_kernel<<<d->n_qry, 64, 0, c->stream>>>(tmp_rows, counts, c->pos.as<uint64_t>(), N, d->d_rows);
    LAUNCHED(3);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    *total_out = total;
    return KSSD_OK;
}

extern "C" int64_t kssd_dist_stats(kssd_dist_t *d, const kssd_stat_opts_t *o)
{
    if (!d || !o) return fail(KSSD_E_INVAL, "kssd_dist_stats: null");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    if (o->n_neighbors < 0 || o->n_neighbors > 1024 || o->n_neighbors > d->n_ref)
        return fail(KSSD_E_NNEIGH, "neighborN_max %d should smaller than NREF 1024 and ref_num %d", o->n_neighbors, d->n_ref);
    StatParams S;
    S.metric = o->metric; S.correction = o->correction; S.kmerlen = o->kmerlen; S.dim_rd_len = o->dim_rd_len;
    S.skip_zero = o->skip_zero; S.dthreshold = o->dthreshold;
    S.cmprsn_num = o->cmprsn_num ? (double)o->cmprsn_num
                                 : (double)(uint32_t)((uint32_t)d->n_ref * (uint32_t)d->n_qry);   // 32-bit wrap, command_dist.c:1186
    if (d->d_rows) { cudaFreeAsync(d->d_rows, c->stream); d->d_rows = nullptr; }
    d->n_rows = 0;
    if (d->sparse) {
        // only searches that cannot print a zero-shared cell run sparse; the others need the matrix
        // (a cell with I = 0 has metric 0 -> dist 1 > D, unless its denominator is 0 too: NaN is never "> D" and prints --
        //  Jaccard X + Y - I = 0 needs both sketches empty, containment min(X, Y) = 0 needs one of them empty)
        const bool nan_cells = o->metric == 0 ? (d->empty_qry && d->empty_ref) : (d->empty_qry || d->empty_ref);
        const bool no_zero_rows = o->n_neighbors == 0 && (o->skip_zero || (o->dthreshold < 1.0 && !o->correction && !nan_cells));
        const uint32_t bw = ((uint32_t)d->n_ref + 31) / 32;
        // packed table (gid | count in one word) when the ref ids and the largest possible count leave room in 32 bits
        uint32_t gbits = 1;
        while ((1ull << gbits) - 1 < (uint64_t)d->n_ref) gbits++;                // n_ref <= 2^gbits - 1: no gid is all ones
        const uint32_t cb = 32 - gbits;
        // (the declared sizes pick the packed table; the kernel itself checks every query's real extent against the count bits)
        const bool packed = gbits <= 24 && (uint64_t)d->max_qry_size + 1 < (1ull << cb) - 1 && !getenv("KSSD_SPARSE_UNPACKED");   // env: A/B and tests
        // CTA shape (index_dist.cuh): narrow unless the largest query could touch more references than its table holds --
        // chance hits of max_qry_size random codes in an index of this density, with head-room; a narrow run in which some
        // query overflowed after all is repeated wide
        uint64_t postings = 0, spaces = 0;
        for (const kssd_index_t *ix : d->comp_ix) { postings += ix->n_postings; spaces += ix->space; }
        const double est_touch = spaces ? (double)d->max_qry_size * (double)postings / (double)spaces : 0.0;
        const char *shape = getenv("KSSD_SPARSE_SHAPE");             // "wide" / "narrow": A/B and tests
        // (measured at 10,000 x 100,000: with ~550 chance references per query the wide shape's 512 threads walk the ~18,000
        //  postings of a query faster; the narrow one pays off when a shard leaves a query a few thousand postings)
        bool narrow = shape ? strcmp(shape, "narrow") == 0 : est_touch < 200.0;
        auto smem_of = [&](bool nar) -> size_t {
            const size_t slots = nar ? SparseNarrow::kSlots : SparseWide::kSlots, tile = nar ? SparseNarrow::kTile : SparseWide::kTile;
            return ((packed ? 1ull : 2ull) * slots + 2ull * tile + 1 + bw) * 4;
        };
        // -N: a reference that shares nothing is never listed, so the best n of a query are among the cells the sparse kernel touches
        const bool topn_sparse = o->n_neighbors > 0 && !getenv("KSSD_TOPN_DENSE");
        if ((no_zero_rows || topn_sparse) && smem_of(false) <= 200u * 1024u) {
            const bool trivial = topn_sparse || S.dthreshold >= 1.0;      // (-N lists first and applies the keep rule to what it listed)
            const int nc = (int)d->comps.size();
            CU(cudaEventRecord(c->ev[0], c->stream));
            CU(c->flags.ensure((size_t)d->n_qry * 4));                 // q_cnt
            CU(c->counts.ensure((size_t)d->n_qry * 8));                // q_pos
            CU(c->pos.ensure(((size_t)d->n_qry + 1) * 8));             // q_out
            CU(c->misc.ensure(16 + sizeof(SparseComp) * 256));
            uint8_t *mb = c->misc.as<uint8_t>();
            if (nc) CU(cudaMemcpyAsync(mb + 16, d->comps.data(), sizeof(SparseComp) * nc, cudaMemcpyHostToDevice, c->stream));
            uint64_t total = 0, cap = std::max<uint64_t>(1ull << 22, (uint64_t)d->n_qry * 1024);
            uint32_t n_over = 0, bad_extent = 0;
            CU(c->ords2.ensure(((size_t)d->n_qry + 1) * 4));           // queries that overflow the table (-N: rows listed per query)
            for (int attempt = 0;; attempt++) {
                if (cap > 0xffffffffull) return fail(KSSD_E_NOMEM, "kssd_dist_stats: more than 2^32 rows pass the filter; tighten -D");
                CU(c->keys.ensure(cap * sizeof(SparseHit)));
                CU(cudaMemsetAsync(mb, 0, 16, c->stream));
                const size_t smem = smem_of(narrow);
                auto launch = [&](auto kern, int threads, int max_per_sm) -> int {
                    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const int per_sm = std::max(1, std::min(max_per_sm, (int)((227u * 1024u) / (smem + 3400))));
                    const uint32_t grid = (uint32_t)std::min<int>(d->n_qry, c->sm_count * per_sm);
                    kern<<<grid, threads, smem, c->stream>>>(reinterpret_cast<const SparseComp *>(mb + 16), nc, (uint32_t)d->n_qry, (uint32_t)d->n_ref, cb, S,
                                                             d->d_qsz, d->d_rsz, c->flags.as<uint32_t>(), c->counts.as<unsigned long long>(),
                                                             reinterpret_cast<unsigned long long *>(mb), cap, c->keys.as<SparseHit>(),
                                                             reinterpret_cast<uint32_t *>(mb + 8), c->ords2.as<uint32_t>(), reinterpret_cast<uint32_t *>(mb + 12));
                    return KSSD_OK;
                };
                int lrc;
                // (wide: the variant that evaluates the keep rule in place needs more registers -- two CTAs per SM measured faster than three)
                if (narrow) {
                    if (trivial) lrc = packed ? launch(dist_sparse_kernel<SparseNarrow, true, true>, 128, 12) : launch(dist_sparse_kernel<SparseNarrow, true, false>, 128, 12);
                    else lrc = packed ? launch(dist_sparse_kernel<SparseNarrow, false, true>, 128, 12) : launch(dist_sparse_kernel<SparseNarrow, false, false>, 128, 12);
                } else {
                    if (trivial) lrc = packed ? launch(dist_sparse_kernel<SparseWide, true, true>, 512, 3) : launch(dist_sparse_kernel<SparseWide, true, false>, 512, 3);
                    else lrc = packed ? launch(dist_sparse_kernel<SparseWide, false, true>, 512, 2) : launch(dist_sparse_kernel<SparseWide, false, false>, 512, 2);
                }
                if (lrc) return lrc;
                LAUNCHED(1);
                CU(cudaEventRecord(c->ev[2], c->stream));
                CU(cudaMemcpyAsync(&total, mb, 8, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaMemcpyAsync(&n_over, mb + 8, 4, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaMemcpyAsync(&bad_extent, mb + 12, 4, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                CU(cudaGetLastError());
                if (bad_extent) return fail(KSSD_E_INVAL, "kssd_dist_stats: a query sketch holds more codes than its declared size allows for");
                if (narrow && n_over && !shape) { narrow = false; attempt = -1; continue; }      // some query needs the big table
                if (total <= cap) break;
                if (attempt > 0) return fail(KSSD_E_NOMEM, "kssd_dist_stats: hit list overflow");
                cap = total;
            }
            if (topn_sparse && n_over == 0) {
                float count_ms = 0;
                CU(cudaEventElapsedTime(&count_ms, c->ev[0], c->ev[2]));
                c->last_ms[3] = count_ms;
                const int N = o->n_neighbors;
                StatRow *tmp_rows = nullptr;
                CU(cudaMallocAsync(&tmp_rows, (size_t)d->n_qry * N * sizeof(StatRow), c->stream));
                topn_sparse_kernel<<<d->n_qry, 256, 0, c->stream>>>(S, c->keys.as<SparseHit>(), c->counts.as<unsigned long long>(), c->flags.as<uint32_t>(), d->d_qsz,
                                                                    d->d_rsz, N, c->ords2.as<uint32_t>(), tmp_rows);
                LAUNCHED(1);
                uint64_t listed = 0;
                const int grc = topn_collect(c, d, tmp_rows, N, c->ords2.as<uint32_t>(), &listed);
                cudaFreeAsync(tmp_rows, c->stream);
                if (grc) return grc;
                d->n_rows = listed;
                CU(cudaEventRecord(c->ev[1], c->stream));
                CU(cudaStreamSynchronize(c->stream));
                CU(cudaEventElapsedTime(&c->last_ms[4], c->ev[2], c->ev[1]));
                return (int64_t)d->n_rows;
            }
            if (!topn_sparse && (uint64_t)n_over * 2 <= (uint64_t)d->n_qry) {
                float count_ms = 0;
                CU(cudaEventElapsedTime(&count_ms, c->ev[0], c->ev[2]));             // counting + listing
                c->last_ms[3] = count_ms;
                uint32_t *p_cnt = c->flags.as<uint32_t>();
                unsigned long long *p_pos = c->counts.as<unsigned long long>();
                SparseHit *p_hits = c->keys.as<SparseHit>();
                // queries whose refs did not fit the table: a small dense sub-job (n_over x R matrix), merged in print order
                kssd_dist *sub = nullptr;
                uint64_t sub_rows = 0;
                std::vector<uint32_t> over(n_over);
                std::vector<uint64_t> first(n_over + 1, 0);
                std::vector<void *> scratch;
                uint32_t *d_list = nullptr;
                auto cleanup = [&] { for (void *p : scratch) cudaFreeAsync(p, c->stream); if (sub) kssd_dist_free(sub); };
                if (n_over) {
                    // the sub-job below goes through the context's scratch buffers: keep what the sparse kernel left there
                    void *keep[3] = {nullptr, nullptr, nullptr};
                    const size_t kb[3] = {(size_t)d->n_qry * 4, (size_t)d->n_qry * 8, std::max<uint64_t>(total, 1) * sizeof(SparseHit)};
                    const void *src[3] = {p_cnt, p_pos, p_hits};
                    for (int i = 0; i < 3; i++) {
                        CU(cudaMallocAsync(&keep[i], kb[i], c->stream));
                        scratch.push_back(keep[i]);
                        CU(cudaMemcpyAsync(keep[i], src[i], kb[i], cudaMemcpyDeviceToDevice, c->stream));
                    }
                    p_cnt = static_cast<uint32_t *>(keep[0]);
                    p_pos = static_cast<unsigned long long *>(keep[1]);
                    p_hits = static_cast<SparseHit *>(keep[2]);
                    CU(cudaMemcpyAsync(over.data(), c->ords2.p, (size_t)n_over * 4, cudaMemcpyDeviceToHost, c->stream));
                    CU(cudaStreamSynchronize(c->stream));
                    std::sort(over.begin(), over.end());
                    std::vector<uint32_t> sub_qsz(n_over);
                    for (uint32_t i = 0; i < n_over; i++) sub_qsz[i] = d->h_qsz[over[i]];
                    int rc = dist_create(c, (int)n_over, d->n_ref, sub_qsz.data(), d->h_rsz->data(), nullptr, 0, &sub);
                    if (rc) return rc;
                    CU(cudaMallocAsync(&d_list, (size_t)n_over * 4, c->stream));
                    scratch.push_back(d_list);
                    CU(cudaMemcpyAsync(d_list, over.data(), (size_t)n_over * 4, cudaMemcpyHostToDevice, c->stream));
                    std::vector<uint64_t> qix((size_t)d->n_qry + 1), six((size_t)n_over + 1);
                    for (int cc = 0; cc < nc; cc++) {
                        CU(cudaMemcpyAsync(qix.data(), d->comps[cc].qindex, qix.size() * 8, cudaMemcpyDeviceToHost, c->stream));
                        CU(cudaStreamSynchronize(c->stream));
                        six[0] = 0;
                        for (uint32_t i = 0; i < n_over; i++) six[i + 1] = six[i] + (qix[over[i] + 1] - qix[over[i]]);
                        uint64_t *d_six = nullptr;
                        uint32_t *d_codes = nullptr;
                        CU(cudaMallocAsync(&d_six, six.size() * 8, c->stream));
                        scratch.push_back(d_six);
                        CU(cudaMallocAsync(&d_codes, std::max<uint64_t>(six[n_over], 1) * 4, c->stream));
                        scratch.push_back(d_codes);
                        CU(cudaMemcpyAsync(d_six, six.data(), six.size() * 8, cudaMemcpyHostToDevice, c->stream));
                        gather_query_codes_kernel<<<n_over, 256, 0, c->stream>>>(d->comps[cc].qcodes, d->comps[cc].qindex, d_list, d_six, d_codes);
                        LAUNCHED(1);
                        rc = kssd_dist_accumulate_dev(sub, d->comp_ix[cc], d_codes, d_six, six[n_over]);
                        if (rc) { cleanup(); return rc; }
                    }
                    kssd_stat_opts_t o2 = *o;
                    o2.cmprsn_num = (uint64_t)S.cmprsn_num;                        // the whole job's comparison count
                    const int64_t nr = kssd_dist_stats(sub, &o2);
                    if (nr < 0) { cleanup(); return (int)nr; }
                    sub_rows = (uint64_t)nr;
                    // per-query row counts of the sub-job -> q_cnt of the whole job
                    uint32_t *d_cnt = nullptr;
                    CU(cudaMallocAsync(&d_cnt, (size_t)n_over * 4, c->stream));
                    scratch.push_back(d_cnt);
                    CU(cudaMemsetAsync(d_cnt, 0, (size_t)n_over * 4, c->stream));
                    if (sub_rows) count_rows_by_query_kernel<<<(uint32_t)((sub_rows + 255) / 256), 256, 0, c->stream>>>(sub->d_rows, sub_rows, d_cnt);
                    std::vector<uint32_t> cnt(n_over), qc((size_t)d->n_qry);
                    CU(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)n_over * 4, cudaMemcpyDeviceToHost, c->stream));
                    CU(cudaMemcpyAsync(qc.data(), p_cnt, (size_t)d->n_qry * 4, cudaMemcpyDeviceToHost, c->stream));
                    CU(cudaStreamSynchronize(c->stream));
                    for (uint32_t i = 0; i < n_over; i++) { qc[over[i]] = cnt[i]; first[i + 1] = first[i] + cnt[i]; }
                    CU(cudaMemcpyAsync(p_cnt, qc.data(), (size_t)d->n_qry * 4, cudaMemcpyHostToDevice, c->stream));
                    CU(cudaStreamSynchronize(c->stream));                          // qc dies here
                    c->last_ms[3] = count_ms;
                }
                const uint64_t all_rows = total + sub_rows;
                if (all_rows > 0xffffffffull) { cleanup(); return fail(KSSD_E_NOMEM, "kssd_dist_stats: more than 2^32 rows pass the filter; tighten -D"); }
                CU(cudaMallocAsync(&d->d_rows, std::max<uint64_t>(all_rows, 1) * sizeof(StatRow), c->stream));
                if (all_rows) {
                    size_t tmp = 0;
                    CU(c->pos.ensure(((size_t)d->n_qry + 1) * 8));
                    cub::DeviceScan::ExclusiveSum(nullptr, tmp, p_cnt, c->pos.as<uint64_t>(), d->n_qry, c->stream);
                    CU(c->cubtmp.ensure(tmp));
                    CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, p_cnt, c->pos.as<uint64_t>(), d->n_qry, c->stream));
                    if (total)
                        stats_rows_sparse_kernel<<<(uint32_t)((total + kStatThreads - 1) / kStatThreads), kStatThreads, 0, c->stream>>>(
                            S, d->d_qsz, d->d_rsz, p_hits, total, p_pos, c->pos.as<uint64_t>(), d->d_rows);
                    if (sub_rows) {
                        uint64_t *d_first = nullptr;
                        CU(cudaMallocAsync(&d_first, first.size() * 8, c->stream));
                        scratch.push_back(d_first);
                        CU(cudaMemcpyAsync(d_first, first.data(), first.size() * 8, cudaMemcpyHostToDevice, c->stream));
                        place_sub_rows_kernel<<<(uint32_t)((sub_rows + 255) / 256), 256, 0, c->stream>>>(
                            sub->d_rows, sub_rows, d_list, d_first, c->pos.as<uint64_t>(), d->d_rows);
                    }
                    LAUNCHED(4);
                }
                d->n_rows = all_rows;
                CU(cudaEventRecord(c->ev[1], c->stream));
                if (n_over) {
                    CU(cudaStreamSynchronize(c->stream));
                    CU(cudaGetLastError());
                    CU(cudaEventElapsedTime(&c->last_ms[4], c->ev[2], c->ev[1]));
                } else c->stats_ms_pending = true;      // the rows kernel runs on; kssd_ctx_last_ms(4) / the next fetch wait for it
                cleanup();
                return (int64_t)d->n_rows;
            }
        }
        // most queries touch too many references, or zero cells are wanted: go through the matrix after all
        const int rc = dist_densify(d);
        if (rc) return rc;
    }
    if (d->components_done == 0) CU(cudaMemsetAsync(d->d_ct, 0, (size_t)d->n_qry * d->n_ref * 4, c->stream));
    CU(cudaEventRecord(c->ev[0], c->stream));
    if (o->n_neighbors > 0) {
        const int N = o->n_neighbors;
        StatRow *tmp_rows = nullptr;
        CU(cudaMallocAsync(&tmp_rows, (size_t)d->n_qry * N * sizeof(StatRow), c->stream));
        CU(c->flags.ensure(((size_t)d->n_qry + 1) * 4));
        topn_kernel<<<d->n_qry, 256, 0, c->stream>>>(S, d->d_ct, d->d_qsz, d->d_rsz, (uint32_t)d->n_ref, N, c->flags.as<uint32_t>(), tmp_rows);
        LAUNCHED(1);
        uint64_t total = 0;
        {
            const int grc = topn_collect(c, d, tmp_rows, N, c->flags.as<uint32_t>(), &total);
            if (grc) { cudaFreeAsync(tmp_rows, c->stream); return grc; }
        }
        cudaFreeAsync(tmp_rows, c->stream);
        d->n_rows = total;
    } else if (S.dthreshold >= 1.0 && !o->skip_zero) {
        const uint64_t cells = (uint64_t)d->n_qry * d->n_ref;
        CU(cudaMallocAsync(&d->d_rows, cells * sizeof(StatRow), c->stream));
        stats_rows_dense_kernel<<<(uint32_t)((cells + kStatThreads - 1) / kStatThreads), kStatThreads, 0, c->stream>>>(
            S, d->d_ct, d->d_qsz, d->d_rsz, (uint32_t)d->n_ref, cells, d->d_rows);
        LAUNCHED(1);
        d->n_rows = cells;
    } else {
        const uint32_t cpr = ((uint32_t)d->n_ref + kStatRefsPerBlock - 1) / kStatRefsPerBlock;
        const uint64_t nchunks = (uint64_t)cpr * d->n_qry;
        const uint64_t cells = (uint64_t)d->n_qry * d->n_ref;
        const uint32_t grid = (uint32_t)((nchunks * 32 + kStatThreads - 1) / kStatThreads);
        const bool trivial = S.dthreshold >= 1.0;
        CU(c->flags.ensure(nchunks * 4));            // chunk_cnt
        CU(c->counts.ensure(nchunks * 4));           // chunk_pos
        CU(c->pos.ensure((nchunks + 1) * 8));        // chunk_out
        CU(c->misc.ensure(8));
        uint64_t total = 0;
        uint64_t cap = std::min<uint64_t>(cells, std::max<uint64_t>(1ull << 24, (uint64_t)d->n_qry * 4096));
        for (int attempt = 0;; attempt++) {
            if (cap > 0xffffffffull) return fail(KSSD_E_NOMEM, "kssd_dist_stats: more than 2^32 rows pass the filter; tighten -D / -N");
            CU(c->keys.ensure(cap * sizeof(uint2)));
            CU(cudaMemsetAsync(c->misc.p, 0, 8, c->stream));
            if (trivial) stats_list_kernel<true><<<grid, kStatThreads, 0, c->stream>>>(S, d->d_ct, d->d_qsz, d->d_rsz, (uint32_t)d->n_ref, cpr, nchunks, c->flags.as<uint32_t>(),
                                                                                     c->counts.as<uint32_t>(), c->misc.as<unsigned long long>(), cap, c->keys.as<uint2>());
            else stats_list_kernel<false><<<grid, kStatThreads, 0, c->stream>>>(S, d->d_ct, d->d_qsz, d->d_rsz, (uint32_t)d->n_ref, cpr, nchunks, c->flags.as<uint32_t>(),
                                                                              c->counts.as<uint32_t>(), c->misc.as<unsigned long long>(), cap, c->keys.as<uint2>());
            LAUNCHED(1);
            CU(cudaMemcpyAsync(&total, c->misc.p, 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            if (total <= cap) break;
            if (attempt) return fail(KSSD_E_NOMEM, "kssd_dist_stats: pair list overflow");
            cap = total;
        }
        CU(cudaMallocAsync(&d->d_rows, std::max<uint64_t>(total, 1) * sizeof(StatRow), c->stream));
        if (total) {
            size_t tmp = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->flags.as<uint32_t>(), c->pos.as<uint64_t>(), nchunks, c->stream);
            CU(c->cubtmp.ensure(tmp));
            CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->flags.as<uint32_t>(), c->pos.as<uint64_t>(), nchunks, c->stream));
            stats_rows_kernel<<<(uint32_t)((total + kStatThreads - 1) / kStatThreads), kStatThreads, 0, c->stream>>>(
                S, d->d_ct, d->d_qsz, d->d_rsz, (uint32_t)d->n_ref, cpr, c->counts.as<uint32_t>(), c->pos.as<uint64_t>(), c->keys.as<uint2>(), total, d->d_rows);
            LAUNCHED(3);
        }
        d->n_rows = total;
    }
    CU(cudaEventRecord(c->ev[1], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&c->last_ms[4], c->ev[0], c->ev[1]));
    return (int64_t)d->n_rows;
}

// Sparse search without a host round trip: count + list kernel, scan of the per-query row counts and the rows kernel are all
// queued at once -- the rows kernel covers the hit list's capacity and reads the hit count on the device -- and the outcome
// (hits, overflowing queries) is copied to pinned memory behind an event.  A host that feeds batch after batch never waits for
// the GPU, and the GPU never waits for a host that was descheduled for a moment (one process per GPU, eight on a box).
extern "C" int kssd_dist_stats_async(kssd_dist_t *d, const kssd_stat_opts_t *o)
{
    if (!d || !o) return fail(KSSD_E_INVAL, "kssd_dist_stats_async: null");
    if (!d->sparse) return fail(KSSD_E_INVAL, "kssd_dist_stats_async: sparse jobs only (kssd_dist_create_sparse)");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    const bool nan_cells = o->metric == 0 ? (d->empty_qry && d->empty_ref) : (d->empty_qry || d->empty_ref);
    const bool no_zero_rows = o->n_neighbors == 0 && (o->skip_zero || (o->dthreshold < 1.0 && !o->correction && !nan_cells));
    d->async_opts = *o;
    d->async_pending = false;
    if (!no_zero_rows) return KSSD_OK;                       // needs the matrix: kssd_dist_stats_wait runs the ordinary path
    StatParams S;
    S.metric = o->metric; S.correction = o->correction; S.kmerlen = o->kmerlen; S.dim_rd_len = o->dim_rd_len;
    S.skip_zero = o->skip_zero; S.dthreshold = o->dthreshold;
    S.cmprsn_num = o->cmprsn_num ? (double)o->cmprsn_num : (double)(uint32_t)((uint32_t)d->n_ref * (uint32_t)d->n_qry);
    const uint32_t bw = ((uint32_t)d->n_ref + 31) / 32;
    uint32_t gbits = 1;
    while ((1ull << gbits) - 1 < (uint64_t)d->n_ref) gbits++;
    const uint32_t cb = 32 - gbits;
    const bool packed = gbits <= 24 && (uint64_t)d->max_qry_size + 1 < (1ull << cb) - 1 && !getenv("KSSD_SPARSE_UNPACKED");
    uint64_t postings = 0, spaces = 0;
    for (const kssd_index_t *ix : d->comp_ix) { postings += ix->n_postings; spaces += ix->space; }
    const double est_touch = spaces ? (double)d->max_qry_size * (double)postings / (double)spaces : 0.0;
    const char *shape = getenv("KSSD_SPARSE_SHAPE");
    const bool narrow = shape ? strcmp(shape, "narrow") == 0 : est_touch < 200.0;
    const size_t slots = narrow ? SparseNarrow::kSlots : SparseWide::kSlots, tile = narrow ? SparseNarrow::kTile : SparseWide::kTile;
    const size_t smem = ((packed ? 1ull : 2ull) * slots + 2ull * tile + 1 + bw) * 4;
    if (smem > 200u * 1024u) return KSSD_OK;
    const bool trivial = S.dthreshold >= 1.0;
    const int nc = (int)d->comps.size();
    if (d->async_slot < 0) {                                 // a pinned 16-byte slot and two events from the context's pool
        if (!c->async_page) {
            CU(cudaMallocHost(&c->async_page, 4096));
            for (int i = 255; i >= 0; i--) c->async_free.push_back(i);
            c->async_events.assign(256, nullptr);
            c->async_counted.assign(256, nullptr);
            CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        }
        if (c->async_free.empty()) return KSSD_OK;           // 256 searches in flight: this one takes the ordinary path in _wait
        d->async_slot = c->async_free.back();
        c->async_free.pop_back();
        if (!c->async_events[d->async_slot]) {
            CU(cudaEventCreateWithFlags(&c->async_events[d->async_slot], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->async_counted[d->async_slot], cudaEventDisableTiming));
        }
        d->h_async = c->async_page + 2 * d->async_slot;
        d->done = c->async_events[d->async_slot];
        d->counted = c->async_counted[d->async_slot];
    }
    const uint64_t cap = std::max<uint64_t>(1ull << 20, (uint64_t)d->n_qry * 1024);
    d->async_cap = cap;
    // per-job scratch (stream-ordered): misc | q_cnt u32[Q] | q_pos u64[Q] | q_out u64[Q+1] | over_list u32[Q] | hits[cap] | scan temp
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_misc = 0, o_cnt = up(16 + sizeof(SparseComp) * std::max(nc, 1)), o_pos = o_cnt + up((size_t)d->n_qry * 4),
                 o_out = o_pos + up((size_t)d->n_qry * 8), o_over = o_out + up(((size_t)d->n_qry + 1) * 8), o_hits = o_over + up((size_t)d->n_qry * 4),
                 o_tmp = o_hits + up(cap * sizeof(SparseHit));
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t *)nullptr, (uint64_t *)nullptr, d->n_qry, c->stream2);
    const size_t total_b = o_tmp + up(tmp);
    if (d->d_async) { cudaFreeAsync(d->d_async, c->stream); d->d_async = nullptr; }
    CU(cudaMallocAsync(&d->d_async, total_b, c->stream));
    uint8_t *mb = d->d_async + o_misc;
    uint32_t *q_cnt = reinterpret_cast<uint32_t *>(d->d_async + o_cnt), *over_list = reinterpret_cast<uint32_t *>(d->d_async + o_over);
    unsigned long long *q_pos = reinterpret_cast<unsigned long long *>(d->d_async + o_pos);
    uint64_t *q_out = reinterpret_cast<uint64_t *>(d->d_async + o_out);
    SparseHit *hits = reinterpret_cast<SparseHit *>(d->d_async + o_hits);
    if (d->d_rows) { cudaFreeAsync(d->d_rows, c->stream); d->d_rows = nullptr; }
    CU(cudaMallocAsync(&d->d_rows, cap * sizeof(StatRow), c->stream));
    if (nc) CU(cudaMemcpyAsync(mb + 16, d->comps.data(), sizeof(SparseComp) * nc, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(mb, 0, 16, c->stream));
    auto launch = [&](auto kern, int threads, int max_per_sm) -> int {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = std::max(1, std::min(max_per_sm, (int)((227u * 1024u) / (smem + 3400))));
        const uint32_t grid = (uint32_t)std::min<int>(d->n_qry, c->sm_count * per_sm);
        kern<<<grid, threads, smem, c->stream>>>(reinterpret_cast<const SparseComp *>(mb + 16), nc, (uint32_t)d->n_qry, (uint32_t)d->n_ref, cb, S, d->d_qsz,
                                                 d->d_rsz, q_cnt, q_pos, reinterpret_cast<unsigned long long *>(mb),
                                                 cap, hits, reinterpret_cast<uint32_t *>(mb + 8), over_list,
                                                 reinterpret_cast<uint32_t *>(mb + 12));
        return KSSD_OK;
    };
    int lrc;
    if (narrow) {
        if (trivial) lrc = packed ? launch(dist_sparse_kernel<SparseNarrow, true, true>, 128, 12) : launch(dist_sparse_kernel<SparseNarrow, true, false>, 128, 12);
        else lrc = packed ? launch(dist_sparse_kernel<SparseNarrow, false, true>, 128, 12) : launch(dist_sparse_kernel<SparseNarrow, false, false>, 128, 12);
    } else {
        if (trivial) lrc = packed ? launch(dist_sparse_kernel<SparseWide, true, true>, 512, 3) : launch(dist_sparse_kernel<SparseWide, true, false>, 512, 3);
        else lrc = packed ? launch(dist_sparse_kernel<SparseWide, false, true>, 512, 2) : launch(dist_sparse_kernel<SparseWide, false, false>, 512, 2);
    }
    if (lrc) return lrc;
    // the rows pass (print order by a scan of the per-query tallies, then one thread per hit) goes to the side stream: it is
    // memory bound, the count kernel is latency bound, and the next search's count kernel starts right behind this one
    CU(cudaEventRecord(d->counted, c->stream));
    CU(cudaStreamWaitEvent(c->stream2, d->counted, 0));
    CU(cub::DeviceScan::ExclusiveSum(d->d_async + o_tmp, tmp, q_cnt, q_out, d->n_qry, c->stream2));
    stats_rows_sparse_dev_kernel<<<(uint32_t)((cap + kStatThreads - 1) / kStatThreads), kStatThreads, 0, c->stream2>>>(
        S, d->d_qsz, d->d_rsz, hits, reinterpret_cast<const unsigned long long *>(mb), cap, q_pos, q_out, d->d_rows);
    LAUNCHED(4);
    CU(cudaMemcpyAsync(d->h_async, mb, 16, cudaMemcpyDeviceToHost, c->stream2));
    CU(cudaEventRecord(d->done, c->stream2));
    d->async_pending = true;
    return KSSD_OK;
}

// waits for kssd_dist_stats_async and returns the number of rows; anything the fast path could not finish (a query that
// overflowed its table, more hits than the list holds, options that print zero cells) is redone by kssd_dist_stats here
extern "C" int64_t kssd_dist_stats_wait(kssd_dist_t *d)
{
    if (!d) return fail(KSSD_E_INVAL, "kssd_dist_stats_wait: null");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    if (d->async_pending) {
        d->async_pending = false;
        CU(cudaEventSynchronize(d->done));
        CU(cudaGetLastError());
        const uint64_t total = d->h_async[0];
        const uint32_t n_over = (uint32_t)d->h_async[1], bad_extent = (uint32_t)(d->h_async[1] >> 32);
        if (bad_extent) return fail(KSSD_E_INVAL, "kssd_dist_stats: a query sketch holds more codes than its declared size allows for");
        if (n_over == 0 && total <= d->async_cap) { d->n_rows = total; return (int64_t)total; }
    }
    return kssd_dist_stats(d, &d->async_opts);
}

extern "C" int kssd_dist_fetch_stats(const kssd_dist_t *d, kssd_stat_row_t *rows_out)
{
    static_assert(sizeof(kssd_stat_row_t) == sizeof(StatRow), "row layout");
    if (!d || !rows_out) return fail(KSSD_E_INVAL, "kssd_dist_fetch_stats: null");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    if (d->n_rows) CU(cudaMemcpyAsync(rows_out, d->d_rows, d->n_rows * sizeof(StatRow), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

extern "C" void kssd_dist_free(kssd_dist_t *d)
{
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaStream_t st = d->ctx->stream;
    if (d->async_slot >= 0) {
        if (d->async_pending) cudaEventSynchronize(d->done);   // the slot must not be reused while a copy into it is queued
        cudaStreamWaitEvent(st, d->done, 0);                   // the rows pass ran on the side stream: the frees below come after it
        d->ctx->async_free.push_back(d->async_slot);
    }
    if (d->owns_ct && d->d_ct) cudaFreeAsync(d->d_ct, st);
    for (void *p : d->owned) cudaFreeAsync(p, st);
    cudaFreeAsync(d->d_qsz, st);
    d->h_rsz.reset();                                        // (d_rsz belongs to the context's cache of reference sets)
    if (d->d_rows) cudaFreeAsync(d->d_rows, st);
    if (d->d_async) cudaFreeAsync(d->d_async, st);
    delete d;
}


// ---- distance.out text (reference dist_print_nobin / output_ctrl, command_dist.c:1188-1195, 1267-1285) ----
namespace {
const char *const kDistHeader[2][3] = {{"Jaccard\tMashD", "P-value(J)\tFDR(J)", "Jaccard_CI\tMashD_CI"},
                                       {"ContainmentM\tAafD", "P-value(C)\tFDR(C)", "ContainmentM_CI\tAafD_CI"}};

void format_rows_range(const kssd_stat_row_t *rows, size_t lo, size_t hi, const char *qn, const char *rn, size_t stride, int outfields,
                       std::string &out)
{
    char line[1024];
    out.reserve((hi - lo) * 160);
    for (size_t i = lo; i < hi; i++) {
        const kssd_stat_row_t &r = rows[i];
        int n = snprintf(line, sizeof line, "%s\t%s\t%u-%u|%u|%u\t%.6lf\t%.6lf", qn + (size_t)r.qry * stride, rn + (size_t)r.ref * stride, r.shared,
                         r.rs_u, r.ref_size, r.qry_size, r.metric, r.dist);
        if (n < 0) n = 0;
        if ((size_t)n >= sizeof line) n = sizeof line - 1;
        if (outfields >= 1) n += snprintf(line + n, sizeof line - n, "\t%E\t%E", r.pvalue, r.fdr);
        if (outfields >= 2 && (size_t)n < sizeof line)
            n += snprintf(line + n, sizeof line - n, "\t[%.6lf,%.6lf]\t[%.6lf,%.6lf]", r.ci_metric_lo, r.ci_metric_hi, r.ci_dist_lo, r.ci_dist_hi);
        if ((size_t)n >= sizeof line - 1) n = sizeof line - 2;
        line[n++] = '\n';
        out.append(line, (size_t)n);
    }
}
}  // namespace

extern "C" int kssd_format_distance_rows(const kssd_stat_row_t *rows, size_t n_rows, const char *qry_names, const char *ref_names,
                                         size_t name_stride, int metric, int outfields, int with_header, int n_threads, char **text_out,
                                         size_t *text_len)
{
    if (!text_out || !text_len || (n_rows && (!rows || !qry_names || !ref_names)) || metric < 0 || metric > 1 || outfields < 0 || outfields > 2)
        return fail(KSSD_E_INVAL, "kssd_format_distance_rows: bad argument");
    unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(1, n_rows / 4096));
    std::vector<std::string> parts(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) {
        const size_t lo = n_rows * t / nt, hi = n_rows * (t + 1) / nt;
        if (t + 1 == nt) format_rows_range(rows, lo, hi, qry_names, ref_names, name_stride, outfields, parts[t]);
        else th.emplace_back(format_rows_range, rows, lo, hi, qry_names, ref_names, name_stride, outfields, std::ref(parts[t]));
    }
    for (auto &x : th) x.join();
    std::string head;
    if (with_header) {
        head = "Qry\tRef\tShared_k|Ref_s|Qry_s";
        for (int i = 0; i <= outfields; i++) { head += "\t"; head += kDistHeader[metric][i]; }
        head += "\n";
    }
    size_t total = head.size();
    for (auto &p : parts) total += p.size();
    char *buf = (char *)malloc(total + 1);
    if (!buf) return fail(KSSD_E_NOMEM, "kssd_format_distance_rows: out of host memory");
    size_t off = 0;
    memcpy(buf, head.data(), head.size());
    off += head.size();
    for (auto &p : parts) { memcpy(buf + off, p.data(), p.size()); off += p.size(); }
    buf[off] = 0;
    *text_out = buf;
    *text_len = off;
    return KSSD_OK;
}

// distance.out straight from rows on the device: lengths, scan, lines (dist_text.cuh); the host only adds the header.
// If the integer formatter handed a value back, the rows are fetched and formatted by kssd_format_distance_rows instead.
// `pinned`: the text goes to the context's pinned buffer (*text_out points into it) instead of a malloc'd one.
static int format_text_dev(kssd_ctx *c, const StatRow *d_rows, uint64_t n, int n_qry, int n_ref, const char *qry_names, const char *ref_names,
                           size_t name_stride, int metric, int outfields, int with_header, bool pinned, char **text_out, size_t *text_len)
{
    std::string head;
    if (with_header) {
        head = "Qry\tRef\tShared_k|Ref_s|Qry_s";
        for (int i = 0; i <= outfields; i++) { head += "\t"; head += kDistHeader[metric][i]; }
        head += "\n";
    }
    uint64_t total = 0;
    uint32_t handed[2] = {0, 0};
    StreamScratch scr(c->stream);
    char *d_text = nullptr;
    if (n && (n_qry <= 0 || n_ref <= 0)) return fail(KSSD_E_INVAL, "distance.out text: rows without names");
    if (n) {
        const size_t qb = (size_t)n_qry * name_stride, rb = (size_t)n_ref * name_stride;
        char *d_qn = static_cast<char *>(scr.alloc(qb)), *d_rn = static_cast<char *>(scr.alloc(rb));
        uint16_t *d_ql = static_cast<uint16_t *>(scr.alloc((size_t)n_qry * 2)), *d_rl = static_cast<uint16_t *>(scr.alloc((size_t)n_ref * 2));
        uint32_t *d_len = static_cast<uint32_t *>(scr.alloc((n + 1) * 4)), *d_handed = static_cast<uint32_t *>(scr.alloc(8));
        uint64_t *d_off = static_cast<uint64_t *>(scr.alloc((n + 1) * 8));
        if (!d_qn || !d_rn || !d_ql || !d_rl || !d_len || !d_handed || !d_off) return fail(KSSD_E_NOMEM, "distance.out text: out of device memory");
        CU(cudaMemcpyAsync(d_qn, qry_names, qb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_rn, ref_names, rb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemsetAsync(d_handed, 0, 8, c->stream));
        CU(cudaMemsetAsync(d_len + n, 0, 4, c->stream));
        CU(cudaEventRecord(c->ev[4], c->stream));
        name_len_kernel<<<(uint32_t)((n_qry + 255) / 256), 256, 0, c->stream>>>(d_qn, name_stride, (uint32_t)n_qry, d_ql);
        name_len_kernel<<<(uint32_t)((n_ref + 255) / 256), 256, 0, c->stream>>>(d_rn, name_stride, (uint32_t)n_ref, d_rl);
        const uint32_t nb = (uint32_t)((n + kTextThreads - 1) / kTextThreads);
        dist_text_len_kernel<<<nb, kTextThreads, 0, c->stream>>>(d_rows, n, (uint32_t)n_qry, (uint32_t)n_ref, d_ql, d_rl, outfields, d_len, d_handed);
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_len, d_off, n + 1, c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, d_len, d_off, n + 1, c->stream));
        CU(cudaMemcpyAsync(&total, d_off + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(handed, d_handed, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        LAUNCHED(5);
        if (handed[1]) return fail(KSSD_E_INVAL, "distance.out text: a row names a query or reference outside the name lists");
        if (handed[0] || getenv("KSSD_TEXT_ON_HOST")) {      // (env: tests of this branch)
            std::vector<kssd_stat_row_t> rows(n);
            CU(cudaMemcpyAsync(rows.data(), d_rows, n * sizeof(StatRow), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            char *t = nullptr;
            size_t tl = 0;
            const int rc = kssd_format_distance_rows(rows.data(), n, qry_names, ref_names, name_stride, metric, outfields, with_header, 0, &t, &tl);
            if (rc || !pinned) { *text_out = t; *text_len = tl; return rc; }
            if (tl + 1 > c->h_text_cap) {
                if (c->h_text) cudaFreeHost(c->h_text);
                c->h_text = nullptr; c->h_text_cap = 0;
                if (cudaMallocHost(&c->h_text, tl + 1) != cudaSuccess) { free(t); return fail(KSSD_E_NOMEM, "distance.out text: out of pinned host memory"); }
                c->h_text_cap = tl + 1;
            }
            memcpy(c->h_text, t, tl + 1);
            free(t);
            *text_out = c->h_text; *text_len = tl;
            return KSSD_OK;
        }
        d_text = static_cast<char *>(scr.alloc(std::max<uint64_t>(total, 1)));
        if (!d_text) return fail(KSSD_E_NOMEM, "distance.out text: out of device memory");
        dist_text_write_kernel<<<nb, kTextThreads, 0, c->stream>>>(d_rows, n, d_qn, d_rn, name_stride, d_ql, d_rl, outfields, d_off, d_text);
        LAUNCHED(1);
        CU(cudaEventRecord(c->ev[5], c->stream));
    }
    const size_t need = head.size() + total + 1;
    char *buf = nullptr;
    if (pinned) {
        if (need > c->h_text_cap) {
            if (c->h_text) cudaFreeHost(c->h_text);
            c->h_text = nullptr; c->h_text_cap = 0;
            const size_t cap = need + need / 4;
            if (cudaMallocHost(&c->h_text, cap) != cudaSuccess) return fail(KSSD_E_NOMEM, "distance.out text: out of pinned host memory");
            c->h_text_cap = cap;
        }
        buf = c->h_text;
    } else buf = (char *)malloc(need);
    if (!buf) return fail(KSSD_E_NOMEM, "distance.out text: out of host memory");
    memcpy(buf, head.data(), head.size());
    if (total) {
        const cudaError_t e = cudaMemcpyAsync(buf + head.size(), d_text, total, cudaMemcpyDeviceToHost, c->stream);
        const cudaError_t e2 = e == cudaSuccess ? cudaStreamSynchronize(c->stream) : e;
        if (e2 != cudaSuccess) { if (!pinned) free(buf); return fail(KSSD_E_CUDA, "distance.out text: %s", cudaGetErrorString(e2)); }
        cudaEventElapsedTime(&c->last_ms[5], c->ev[4], c->ev[5]);
    }
    buf[head.size() + total] = 0;
    *text_out = buf;
    *text_len = head.size() + total;
    return KSSD_OK;
}

extern "C" int kssd_dist_format_text(kssd_dist_t *d, const char *qry_names, const char *ref_names, size_t name_stride, int metric, int outfields,
                                     int with_header, char **text_out, size_t *text_len)
{
    if (!d || !text_out || !text_len || metric < 0 || metric > 1 || outfields < 0 || outfields > 2 || name_stride == 0 || name_stride > 65535)
        return fail(KSSD_E_INVAL, "kssd_dist_format_text: bad argument");
    if (d->async_pending) return fail(KSSD_E_INVAL, "kssd_dist_format_text: call kssd_dist_stats_wait first");
    if (d->n_rows && (!qry_names || !ref_names || !d->d_rows)) return fail(KSSD_E_INVAL, "kssd_dist_format_text: no rows (kssd_dist_stats first) or null names");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    return format_text_dev(c, d->d_rows, d->n_rows, d->n_qry, d->n_ref, qry_names, ref_names, name_stride, metric, outfields, with_header, false, text_out, text_len);
}

extern "C" int kssd_dist_text(kssd_dist_t *d, const char *qry_names, const char *ref_names, size_t name_stride, int metric, int outfields, int with_header,
                              const char **text, size_t *text_len)
{
    if (!d || !text || !text_len || metric < 0 || metric > 1 || outfields < 0 || outfields > 2 || name_stride == 0 || name_stride > 65535)
        return fail(KSSD_E_INVAL, "kssd_dist_text: bad argument");
    if (d->async_pending) return fail(KSSD_E_INVAL, "kssd_dist_text: call kssd_dist_stats_wait first");
    if (d->n_rows && (!qry_names || !ref_names || !d->d_rows)) return fail(KSSD_E_INVAL, "kssd_dist_text: no rows (kssd_dist_stats first) or null names");
    kssd_ctx *c = d->ctx;
    CU(cudaSetDevice(c->device));
    char *t = nullptr;
    const int rc = format_text_dev(c, d->d_rows, d->n_rows, d->n_qry, d->n_ref, qry_names, ref_names, name_stride, metric, outfields, with_header, true, &t, text_len);
    *text = t;
    return rc;
}

extern "C" int kssd_format_distance_rows_gpu(kssd_ctx_t *c, const kssd_stat_row_t *rows, size_t n_rows, int n_qry, int n_ref, const char *qry_names,
                                             const char *ref_names, size_t name_stride, int metric, int outfields, int with_header, char **text_out,
                                             size_t *text_len)
{
    if (!c || !text_out || !text_len || metric < 0 || metric > 1 || outfields < 0 || outfields > 2 || name_stride == 0 || name_stride > 65535 || n_qry < 0 ||
        n_ref < 0 || (n_rows && (!rows || !qry_names || !ref_names)))
        return fail(KSSD_E_INVAL, "kssd_format_distance_rows_gpu: bad argument");
    CU(cudaSetDevice(c->device));
    StreamScratch scr(c->stream);
    StatRow *d_rows = nullptr;
    if (n_rows) {
        d_rows = static_cast<StatRow *>(scr.alloc(n_rows * sizeof(StatRow)));
        if (!d_rows) return fail(KSSD_E_NOMEM, "kssd_format_distance_rows_gpu: out of device memory");
        CU(cudaMemcpyAsync(d_rows, rows, n_rows * sizeof(StatRow), cudaMemcpyHostToDevice, c->stream));
    }
    return format_text_dev(c, d_rows, n_rows, n_qry, n_ref, qry_names, ref_names, name_stride, metric, outfields, with_header, false, text_out, text_len);
}

// The integer formatter of fmt_exact.cuh against snprintf on the host: n values per family (any bit pattern, metric-like decimals,
// dyadic fractions where exact ties live, e^-x), returns the number of differing strings; *handed_back counts the values
// the formatter declined.  No GPU involved: the CPU test suite calls it.
extern "C" int64_t kssd_format_selftest(uint64_t n, uint64_t seed, uint64_t *handed_back)
{
    uint64_t st = seed * 0x9e3779b97f4a7c15ull + 0x2545f4914f6cdd1dull, handed = 0;
    auto next = [&]() -> uint64_t { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st * 0x2545f4914f6cdd1dull; };
    int64_t bad = 0;
    auto check = [&](double x) {
        char a[64], b[64];
        int l = fmt::put_f6(a, x);
        if (l < 0) handed++;
        else { a[l] = 0; snprintf(b, sizeof b, "%.6lf", x); bad += strcmp(a, b) != 0; }
        l = fmt::put_e6(a, x);
        if (l < 0) handed++;
        else { a[l] = 0; snprintf(b, sizeof b, "%E", x); bad += strcmp(a, b) != 0; }
    };
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t bits = next();
        double x;
        memcpy(&x, &bits, 8);
        check(x);
        check((double)(next() % 2000001) / 1000000.0 - 0.5);
        check(ldexp((double)(next() >> 11), -(int)(next() % 120)));
        check(((double)(next() % 20000000) + 0.5) / 8.0);
        check(exp(-(double)(next() % 700000) / 1000.0));
    }
    const double special[] = {0.0, -0.0, 1.0, 0.5, 1e-7, 5e-7, 0.0000015, 0.0000025, 9.9999995, 999999.95, 9999999.5, 99999995.0, 1e22, 1e23, 5e-324,
                              2.2250738585072014e-308, 1.7976931348623157e308, NAN, -NAN, INFINITY, -INFINITY, 1e12, 1099511627775.5, 8388608.5, 0.1};
    for (double x : special) check(x);
    if (handed_back) *handed_back = handed;
    return bad;
}

extern "C" void kssd_host_free(void *p) { free(p); }


// ------------------------------------------------------------------------------------------------
// kssd set (reference command_set.c): union / uniq union, intersect / subtract
// ------------------------------------------------------------------------------------------------
static int set_code_words(kssd_ctx *c, uint64_t *n_words)
{
    const int bits = 4 * std::min(c->P.k - c->P.L, c->info.component_sz > 0 ? c->info.component_sz : 7);
    *n_words = (1ull << bits) / 32;
    return bits;
}

// members of (seen & ~twice) ascending into out (device, capacity cap); returns the count through *n_out
static int set_emit(kssd_ctx *c, const uint32_t *seen, const uint32_t *twice, uint64_t n_words, uint32_t *out, uint64_t cap, uint64_t *n_out)
{
    const uint32_t nblk = (uint32_t)((n_words + kSetWordsPerBlock - 1) / kSetWordsPerBlock);
    CU(c->flags.ensure((size_t)nblk * 4));
    CU(c->pos.ensure((size_t)nblk * 4));
    set_count_kernel<<<nblk, kSetThreads, 0, c->stream>>>(seen, twice, n_words, c->flags.as<uint32_t>());
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), nblk, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), nblk, c->stream));
    uint32_t lastoff = 0, lastcnt = 0;
    CU(cudaMemcpyAsync(&lastoff, c->pos.as<uint32_t>() + (nblk - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&lastcnt, c->flags.as<uint32_t>() + (nblk - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const uint64_t total = (uint64_t)lastoff + lastcnt;
    if (total > cap) return fail(KSSD_E_INVAL, "kssd_set_union: output buffer too small (%llu > %llu)", (unsigned long long)total, (unsigned long long)cap);
    if (total) set_fill_kernel<<<nblk, kSetThreads, 0, c->stream>>>(seen, twice, n_words, c->pos.as<uint32_t>(), out);
    LAUNCHED(4);
    *n_out = total;
    return KSSD_OK;
}

extern "C" int kssd_set_union_dev(kssd_ctx_t *c, const uint32_t *combco_dev, uint64_t n_codes, int uniq, uint32_t *pan_dev, uint64_t pan_cap,
                                  uint64_t *n_out)
{
    if (!c || !n_out || (n_codes && (!combco_dev || !pan_dev))) return fail(KSSD_E_INVAL, "kssd_set_union_dev: null argument");
    CU(cudaSetDevice(c->device));
    uint64_t n_words;
    set_code_words(c, &n_words);
    CU(c->keys.ensure(n_words * 4));
    CU(cudaMemsetAsync(c->keys.p, 0, n_words * 4, c->stream));
    uint32_t *twice = nullptr;
    if (uniq) {
        CU(c->keys2.ensure(n_words * 4));
        CU(cudaMemsetAsync(c->keys2.p, 0, n_words * 4, c->stream));
        twice = c->keys2.as<uint32_t>();
    }
    if (n_codes) {
        validate_below(c, combco_dev, n_codes, n_words * 32);
        bool bad = false;
        const int vrc = validation_failed(c, &bad);
        if (vrc) return vrc;
        if (bad) return fail(KSSD_E_INVAL, "kssd_set_union: a code lies outside the component's code space (mismatched or corrupt sketch)");
        set_mark_kernel<<<(uint32_t)((n_codes + 255) / 256), 256, 0, c->stream>>>(combco_dev, n_codes, c->keys.as<uint32_t>(), twice);
        LAUNCHED(1);
    }
    const int rc = set_emit(c, c->keys.as<uint32_t>(), twice, n_words, pan_dev, pan_cap, n_out);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return KSSD_OK;
}

extern "C" int kssd_set_union_host(kssd_ctx_t *c, const uint32_t *combco, uint64_t n_codes, int uniq, uint32_t *pan_out, uint64_t *n_out)
{
    if (!c || !n_out || (n_codes && (!combco || !pan_out))) return fail(KSSD_E_INVAL, "kssd_set_union_host: null argument");
    CU(cudaSetDevice(c->device));
    CU(c->seq.ensure(std::max<uint64_t>(n_codes, 1) * 8));
    uint32_t *d_in = c->seq.as<uint32_t>(), *d_out = d_in + n_codes;
    if (n_codes) CU(cudaMemcpyAsync(d_in, combco, n_codes * 4, cudaMemcpyHostToDevice, c->stream));
    const int rc = kssd_set_union_dev(c, d_in, n_codes, uniq, d_out, n_codes, n_out);
    if (rc) return rc;
    if (*n_out) CU(cudaMemcpyAsync(pan_out, d_out, *n_out * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}

extern "C" int kssd_set_operate_dev(kssd_ctx_t *c, const uint32_t *combco_dev, const uint64_t *index_dev, int n_genomes, uint64_t n_codes,
                                    const uint32_t *pan_dev, uint64_t n_pan, int intersect, uint32_t *combco_out_dev, uint64_t *index_out_dev)
{
    if (!c || !index_dev || !index_out_dev || n_genomes < 0 || (n_codes && (!combco_dev || !combco_out_dev)) || (n_pan && !pan_dev))
        return fail(KSSD_E_INVAL, "kssd_set_operate_dev: bad argument");
    if (n_codes >= 0xffffffffull) return fail(KSSD_E_INVAL, "kssd_set_operate_dev: more than 2^32 codes in one component");
    CU(cudaSetDevice(c->device));
    uint64_t n_words;
    set_code_words(c, &n_words);
    CU(c->keys.ensure(n_words * 4));
    CU(cudaMemsetAsync(c->keys.p, 0, n_words * 4, c->stream));
    {
        validate_below(c, pan_dev, n_pan, n_words * 32);
        validate_below(c, combco_dev, n_codes, n_words * 32);
        bool bad = false;
        const int vrc = validation_failed(c, &bad);
        if (vrc) return vrc;
        if (bad) return fail(KSSD_E_INVAL, "kssd_set_operate: a code lies outside the component's code space (mismatched or corrupt sketch / pan file)");
    }
    if (n_pan) set_mark_kernel<<<(uint32_t)((n_pan + 255) / 256), 256, 0, c->stream>>>(pan_dev, n_pan, c->keys.as<uint32_t>(), nullptr);
    const uint64_t nn = std::max<uint64_t>(n_codes, 1);
    CU(c->flags.ensure(nn * 4));
    CU(c->pos.ensure(nn * 4));
    if (n_codes) {
        const uint32_t nb = (uint32_t)((n_codes + 255) / 256);
        set_flag_kernel<<<nb, 256, 0, c->stream>>>(combco_dev, n_codes, c->keys.as<uint32_t>(), intersect ? 1 : 0, c->flags.as<uint32_t>());
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n_codes, c->stream);
        CU(c->cubtmp.ensure(tmp));
        CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), n_codes, c->stream));
        set_scatter_kernel<<<nb, 256, 0, c->stream>>>(combco_dev, n_codes, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(), combco_out_dev);
    }
    set_index_kernel<<<(n_genomes + 1 + 255) / 256, 256, 0, c->stream>>>(index_dev, n_genomes, n_codes, c->flags.as<uint32_t>(), c->pos.as<uint32_t>(),
                                                                        index_out_dev);
    LAUNCHED(6);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    return KSSD_OK;
}

extern "C" int kssd_set_group_host(kssd_ctx_t *c, const uint32_t *combco, const uint64_t *index, int n_genomes, const uint32_t *member_gids,
                                   const uint64_t *group_index, int n_groups, uint32_t *codes_out, uint64_t *index_out)
{
    if (!c || !index || !group_index || !index_out || n_genomes < 0 || n_groups < 0) return fail(KSSD_E_INVAL, "kssd_set_group_host: bad argument");
    CU(cudaSetDevice(c->device));
    const uint64_t n_codes = index[n_genomes], M = group_index[n_groups];
    if (M && !member_gids) return fail(KSSD_E_INVAL, "kssd_set_group_host: null member list");
    if (M >= 0xffffffffull) return fail(KSSD_E_INVAL, "kssd_set_group_host: too many group members");
    // where every member's codes come from and go to (group-major, members in the given order)
    std::vector<uint64_t> m_src(M + 1, 0), m_dst(M + 1, 0);
    std::vector<uint32_t> m_group(std::max<uint64_t>(M, 1), 0);
    for (int g = 0; g < n_groups; g++) {
        if (group_index[g + 1] < group_index[g]) return fail(KSSD_E_INVAL, "kssd_set_group_host: group index not ascending");
        for (uint64_t m = group_index[g]; m < group_index[g + 1]; m++) {
            const uint32_t gid = member_gids[m];
            if (gid >= (uint32_t)n_genomes) return fail(KSSD_E_INVAL, "kssd_set_group_host: member %u of group %d is not a genome of the sketch (%d genomes)", gid, g, n_genomes);
            if (index[gid + 1] < index[gid] || index[gid + 1] > n_codes) return fail(KSSD_E_INVAL, "kssd_set_group_host: combco index not ascending");
            m_src[m] = index[gid];
            m_dst[m + 1] = m_dst[m] + (index[gid + 1] - index[gid]);
            m_group[m] = (uint32_t)g;
        }
    }
    const uint64_t T = m_dst[M];
    if (T >= 0xffffffffull) return fail(KSSD_E_INVAL, "kssd_set_group_host: more than 2^32 member codes in one component");
    for (int g = 0; g <= n_groups; g++) index_out[g] = 0;
    if (T == 0) return KSSD_OK;
    if (!combco || !codes_out) return fail(KSSD_E_INVAL, "kssd_set_group_host: null code buffer");
    StreamScratch scr(c->stream);
    auto galloc = [&](auto *&ptr, size_t bytes) -> bool { ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(scr.alloc(bytes)); return ptr != nullptr; };
#define GALLOC(ptr, bytes) do { if (!galloc(ptr, bytes)) return fail(KSSD_E_NOMEM, "kssd_set_group_host: out of device memory"); } while (0)
    uint32_t *d_combco = nullptr, *d_mgroup = nullptr, *d_pos = nullptr, *d_pos2 = nullptr, *d_flags = nullptr, *d_excl = nullptr, *d_hpos = nullptr,
             *d_hcode = nullptr, *d_hpos2 = nullptr, *d_hcode2 = nullptr;
    uint64_t *d_msrc = nullptr, *d_mdst = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    GALLOC(d_combco, n_codes * 4); GALLOC(d_msrc, (M + 1) * 8); GALLOC(d_mdst, (M + 1) * 8); GALLOC(d_mgroup, M * 4);
    GALLOC(d_keys, T * 8); GALLOC(d_keys2, T * 8); GALLOC(d_pos, T * 4); GALLOC(d_pos2, T * 4);
    GALLOC(d_flags, T * 4); GALLOC(d_excl, T * 4);
    CU(cudaMemcpyAsync(d_combco, combco, n_codes * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_msrc, m_src.data(), (M + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_mdst, m_dst.data(), (M + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_mgroup, m_group.data(), M * 4, cudaMemcpyHostToDevice, c->stream));
    group_gather_kernel<<<(uint32_t)((M * 32 + 255) / 256), 256, 0, c->stream>>>(d_combco, d_msrc, d_mdst, d_mgroup, (uint32_t)M, d_keys, d_pos);
    int gbits = 1;
    while ((1ll << gbits) < n_groups) gbits++;
    size_t tmp = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_keys, d_keys2, d_pos, d_pos2, T, 0, 32 + gbits, c->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_flags, d_excl, T, c->stream);
    CU(c->cubtmp.ensure(std::max(tmp, tmp2)));
    CU(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp, d_keys, d_keys2, d_pos, d_pos2, T, 0, 32 + gbits, c->stream));      // stable: the earliest position leads its run
    const uint32_t nb = (uint32_t)((T + 255) / 256);
    group_heads_kernel<<<nb, 256, 0, c->stream>>>(d_keys2, T, d_flags);
    CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp2, d_flags, d_excl, T, c->stream));
    uint32_t last_e = 0, last_f = 0;
    CU(cudaMemcpyAsync(&last_e, d_excl + (T - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&last_f, d_flags + (T - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const uint64_t U = (uint64_t)last_e + last_f;                           // distinct (group, code) pairs
    GALLOC(d_hpos, U * 4); GALLOC(d_hcode, U * 4); GALLOC(d_hpos2, U * 4); GALLOC(d_hcode2, U * 4);
    group_compact_kernel<<<nb, 256, 0, c->stream>>>(d_keys2, d_pos2, d_flags, d_excl, T, d_hpos, d_hcode);
    tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, d_hpos, d_hpos2, d_hcode, d_hcode2, U, 0, 32, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp, d_hpos, d_hpos2, d_hcode, d_hcode2, U, 0, 32, c->stream));          // back to the order of first occurrence
    LAUNCHED(12);
    std::vector<uint32_t> hpos(U);
    CU(cudaMemcpyAsync(hpos.data(), d_hpos2, U * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(codes_out, d_hcode2, U * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    for (int g = 0; g <= n_groups; g++)                                      // a group's codes sit where its members' positions do
        index_out[g] = (uint64_t)(std::lower_bound(hpos.begin(), hpos.end(), (uint32_t)std::min<uint64_t>(m_dst[group_index[g]], 0xffffffffull)) - hpos.begin());
    index_out[n_groups] = U;
    return KSSD_OK;
#undef GALLOC
}

extern "C" int kssd_set_operate_host(kssd_ctx_t *c, const uint32_t *combco, const uint64_t *index, int n_genomes, const uint32_t *pan, uint64_t n_pan,
                                     int intersect, uint32_t *combco_out, uint64_t *index_out)
{
    if (!c || !index || !index_out || n_genomes < 0) return fail(KSSD_E_INVAL, "kssd_set_operate_host: bad argument");
    const uint64_t n = index[n_genomes];
    if ((n && (!combco || !combco_out)) || (n_pan && !pan)) return fail(KSSD_E_INVAL, "kssd_set_operate_host: null argument");
    CU(cudaSetDevice(c->device));
    const size_t ib = 8ull * (n_genomes + 1);
    CU(c->seq.ensure(2 * ib + (2 * n + n_pan) * 4 + 64));
    uint64_t *d_ix = c->seq.as<uint64_t>(), *d_ox = d_ix + n_genomes + 1;
    uint32_t *d_in = reinterpret_cast<uint32_t *>(d_ox + n_genomes + 1), *d_out = d_in + n, *d_pan = d_out + n;
    CU(cudaMemcpyAsync(d_ix, index, ib, cudaMemcpyHostToDevice, c->stream));
    if (n) CU(cudaMemcpyAsync(d_in, combco, n * 4, cudaMemcpyHostToDevice, c->stream));
    if (n_pan) CU(cudaMemcpyAsync(d_pan, pan, n_pan * 4, cudaMemcpyHostToDevice, c->stream));
    const int rc = kssd_set_operate_dev(c, d_in, d_ix, n_genomes, n, d_pan, n_pan, intersect, d_out, d_ox);
    if (rc) return rc;
    CU(cudaMemcpyAsync(index_out, d_ox, ib, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const uint64_t kept = index_out[n_genomes];
    if (kept) CU(cudaMemcpyAsync(combco_out, d_out, kept * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return KSSD_OK;
}


// ------------------------------------------------------------------------------------------------
// kssd composite (reference get_species_abundance, command_composite.c:389-547)
// ------------------------------------------------------------------------------------------------
extern "C" int kssd_composite_host(kssd_ctx_t *c, int n_comp, const kssd_index_t *const *ref_ix, const uint32_t *const *qcodes,
                                   const uint64_t *const *qindex, const uint16_t *const *qabund, int n_qry, int min_kmers,
                                   kssd_comp_row_t **rows_out, uint64_t *n_rows)
{
    static_assert(sizeof(kssd_comp_row_t) == sizeof(CompRow), "row layout");
    if (!c || n_comp <= 0 || !ref_ix || !qcodes || !qindex || !qabund || n_qry <= 0 || !rows_out || !n_rows)
        return fail(KSSD_E_INVAL, "kssd_composite_host: bad argument");
    if (n_qry >= (1 << 20)) return fail(KSSD_E_INVAL, "kssd_composite_host: at most 2^20 - 1 queries per call");
    for (int cc = 0; cc < n_comp; cc++) {
        if (!ref_ix[cc] || !qindex[cc]) return fail(KSSD_E_INVAL, "kssd_composite_host: null component %d", cc);
        if (ref_ix[cc]->n_genomes != ref_ix[0]->n_genomes || ref_ix[cc]->n_genomes >= (1 << 24))
            return fail(KSSD_E_MISMATCH, "kssd_composite_host: component %d has %d references", cc, ref_ix[cc]->n_genomes);
    }
    CU(cudaSetDevice(c->device));
    if (min_kmers < 1) min_kmers = 6;                                   // MIN_KM_S
    *rows_out = nullptr;
    *n_rows = 0;
    StreamScratch scratch(c->stream);
    auto dalloc = [&](size_t bytes) -> void * { return scratch.alloc(bytes); };
    struct Comp { uint32_t *codes; uint16_t *ab; uint64_t *index, *off; uint64_t n, pairs; };
    std::vector<Comp> comps(n_comp);
    uint64_t P = 0;
    size_t tmp = 0;
    for (int cc = 0; cc < n_comp; cc++) {
        Comp &C = comps[cc];
        C.n = qindex[cc][n_qry];
        if (C.n && (!qcodes[cc] || !qabund[cc])) { return fail(KSSD_E_INVAL, "kssd_composite_host: null codes in component %d", cc); }
        C.codes = (uint32_t *)dalloc(C.n * 4); C.ab = (uint16_t *)dalloc(C.n * 2); C.index = (uint64_t *)dalloc(8ull * (n_qry + 1));
        C.off = (uint64_t *)dalloc(C.n * 8);
        uint32_t *len = (uint32_t *)dalloc(C.n * 4);
        if (!C.codes || !C.ab || !C.index || !C.off || !len) { return fail(KSSD_E_NOMEM, "kssd_composite_host: out of device memory"); }
        CU(cudaMemcpyAsync(C.index, qindex[cc], 8ull * (n_qry + 1), cudaMemcpyHostToDevice, c->stream));
        C.pairs = 0;
        if (C.n) {
            CU(cudaMemcpyAsync(C.codes, qcodes[cc], C.n * 4, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(C.ab, qabund[cc], C.n * 2, cudaMemcpyHostToDevice, c->stream));
            comp_count_kernel<<<(uint32_t)((C.n + 255) / 256), 256, 0, c->stream>>>(C.codes, C.n, ref_ix[cc]->lookup(), len);
            cub::DeviceScan::ExclusiveSum(nullptr, tmp, len, C.off, C.n, c->stream);
            CU(c->cubtmp.ensure(tmp));
            CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, len, C.off, C.n, c->stream));
            uint64_t lo = 0;
            uint32_t ll = 0;
            CU(cudaMemcpyAsync(&lo, C.off + (C.n - 1), 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaMemcpyAsync(&ll, len + (C.n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            C.pairs = lo + ll;
            LAUNCHED(3);
        }
        P += C.pairs;
    }
    if (P == 0) { CU(cudaStreamSynchronize(c->stream)); return KSSD_OK; }
    if (P >= 0x7fffffffull) { return fail(KSSD_E_NOMEM, "kssd_composite_host: more than 2^31 shared k-mer pairs in one call; split the queries"); }
    uint64_t *keys = (uint64_t *)dalloc(P * 8), *sorted = (uint64_t *)dalloc(P * 8), *groups = (uint64_t *)dalloc(P * 8);
    uint64_t *run_key = (uint64_t *)dalloc(P * 8), *run_off = (uint64_t *)dalloc(P * 8);
    uint32_t *run_len = (uint32_t *)dalloc(P * 4), *d_meta = (uint32_t *)dalloc(16);
    if (!keys || !sorted || !groups || !run_key || !run_off || !run_len || !d_meta) { return fail(KSSD_E_NOMEM, "kssd_composite_host: out of device memory"); }
    uint64_t base = 0;
    for (int cc = 0; cc < n_comp; cc++) {
        const Comp &C = comps[cc];
        if (C.n) comp_emit_kernel<<<(uint32_t)((C.n + 255) / 256), 256, 0, c->stream>>>(C.codes, C.ab, C.index, n_qry, C.n, ref_ix[cc]->lookup(),
                                                                                       ref_ix[cc]->d_gids, C.off, keys + base);
        base += C.pairs;
    }
    int qbits = 1;
    while ((1ll << qbits) < n_qry) qbits++;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, sorted, P, 0, kCompQryShift + qbits, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortKeys(c->cubtmp.p, tmp, keys, sorted, P, 0, kCompQryShift + qbits, c->stream));
    comp_group_kernel<<<(uint32_t)((P + 255) / 256), 256, 0, c->stream>>>(sorted, P, groups);
    cub::DeviceRunLengthEncode::Encode(nullptr, tmp, groups, run_key, run_len, d_meta, (int)P, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cub::DeviceRunLengthEncode::Encode(c->cubtmp.p, tmp, groups, run_key, run_len, d_meta, (int)P, c->stream));
    uint32_t n_runs = 0;
    CU(cudaMemcpyAsync(&n_runs, d_meta, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, run_len, run_off, n_runs, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, run_len, run_off, n_runs, c->stream));
    CompRow *rows = (CompRow *)dalloc((size_t)n_runs * sizeof(CompRow)), *rows_sorted = (CompRow *)dalloc((size_t)n_runs * sizeof(CompRow));
    uint64_t *okey = (uint64_t *)dalloc((size_t)n_runs * 8), *okey2 = (uint64_t *)dalloc((size_t)n_runs * 8);
    uint32_t *perm = (uint32_t *)dalloc((size_t)n_runs * 4), *perm2 = (uint32_t *)dalloc((size_t)n_runs * 4);
    if (!rows || !rows_sorted || !okey || !okey2 || !perm || !perm2) { return fail(KSSD_E_NOMEM, "kssd_composite_host: out of device memory"); }
    CU(cudaMemsetAsync(d_meta + 1, 0, 4, c->stream));
    const uint32_t nb = (n_runs + 255) / 256;
    comp_rows_kernel<<<nb, 256, 0, c->stream>>>(sorted, run_key, run_len, run_off, n_runs, (uint32_t)min_kmers, rows, okey, d_meta + 1);
    comp_iota_kernel<<<nb, 256, 0, c->stream>>>(perm, n_runs);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, okey, okey2, perm, perm2, n_runs, 0, 64, c->stream);
    CU(c->cubtmp.ensure(tmp));
    CU(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp, okey, okey2, perm, perm2, n_runs, 0, 64, c->stream));
    comp_gather_kernel<<<nb, 256, 0, c->stream>>>(rows, perm2, n_runs, rows_sorted);
    LAUNCHED(20);
    uint32_t kept = 0;
    CU(cudaMemcpyAsync(&kept, d_meta + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    if (kept) {
        kssd_comp_row_t *h = (kssd_comp_row_t *)malloc((size_t)kept * sizeof(kssd_comp_row_t));
        if (!h) { return fail(KSSD_E_NOMEM, "kssd_composite_host: out of host memory"); }
        CU(cudaMemcpyAsync(h, rows_sorted, (size_t)kept * sizeof(CompRow), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        *rows_out = h;
        *n_rows = kept;
    }
    return KSSD_OK;
}
