// index_dist.cuh -- Stage II (inverted index) and Stage III (shared counts + statistics) kernels.
#pragma once
#include "kssd_device.cuh"

namespace kssd {

// ------------------------------------------------------------------------------------------------
// Stage II helpers.  combco2mco (reference co2mco.c:42-55) appends genome j to the list of every
// code of genome j, j ascending: postings = gids ordered by (code, gid).  Here: tag each code with
// its gid, stable LSD radix sort by code (gid order is preserved inside a code), run heads -> CSR.
// ------------------------------------------------------------------------------------------------
// gid of every posting position: mark the first position of each non-empty genome with its id, then an inclusive
// max-scan spreads it (ids ascend with position).  Replaces a per-element binary search.
__global__ void mark_genome_starts_kernel(const uint64_t *__restrict__ index, int n_genomes, uint32_t *__restrict__ gid)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_genomes) return;
    const uint64_t a = index[j], b = index[j + 1];
    if (a < b) gid[a] = (uint32_t)j;
}

// The same in one pass and without the scan: a warp takes 1024 consecutive posting positions, finds the genome of the first one by a
// binary search of the (L2-resident) index and walks on from there; every lane also checks its codes against the component's code
// space (a code beyond it has no place in the index: flag, KSSD_E_INVAL).  4 B read + 4 B written per posting.
__global__ void tag_gids_kernel(const uint64_t *__restrict__ index, int n_genomes, uint64_t n, const uint32_t *__restrict__ codes, uint64_t space,
                                uint32_t *__restrict__ gid, uint32_t *__restrict__ flag)
{
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t base = w << 10;
    if (base >= n) return;
    uint32_t lo = 0, hi = (uint32_t)n_genomes;                // index[lo] <= base < index[hi] for a well-formed index
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (index[mid] <= base) lo = mid; else hi = mid;
    }
    uint32_t g = lo;
    bool bad = false;
#pragma unroll 4
    for (int j = 0; j < 32; j++) {
        const uint64_t i = base + 32u * j + lane;
        if (i >= n) break;
        while (g + 1 < (uint32_t)n_genomes && index[g + 1] <= i) g++;
        gid[i] = g;
        bad |= codes[i] >= space;
    }
    if (bad) atomicOr(flag, 1u);
}

// run heads -> CSR by head flags + scan + scatter: three passes; kept for components beyond 2^31 postings (the run-length encode that
// replaced it counts with an int)
__global__ void head_flags_kernel(const uint32_t *__restrict__ codes, uint64_t n, uint32_t *__restrict__ flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || codes[i] != codes[i - 1]) ? 1u : 0u;
}

__global__ void csr_scatter_kernel(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ flags,
                                   const uint32_t *__restrict__ pos, uint64_t n, uint32_t *__restrict__ ucodes,
                                   uint32_t *__restrict__ uoff)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) { ucodes[pos[i]] = codes[i]; uoff[pos[i]] = (uint32_t)i; }
}

// Code -> posting list on the device: a rank bitmap over the component's code space.
//   rb[w] = { bits of the codes 32w .. 32w+31 that occur, number of occurring codes below 32w }
// 64 MiB for the 2^28 codes of a component -- it stays in the 126 MB L2 -- next to uoff[U+1], the list starts of the U
// occurring codes.  An absent code costs one 8-byte L2 read and nothing else; an occurring one adds one HBM read of two
// adjacent offsets.  (The reference's dense 2 GiB mco.index -- one offset per code of the space -- is a file format:
// it is produced on export and consumed on import, never looked up.)  Codes beyond the space have no list.
struct CodeLookup { const uint2 *rb; const uint32_t *uoff; uint32_t n_words; };

__device__ __forceinline__ void code_lookup(const CodeLookup &L, uint32_t c, uint32_t &s0, uint32_t &s1)
{
    s0 = 0; s1 = 0;
    const uint32_t w = c >> 5;
    if (w >= L.n_words) return;
    const uint2 e = __ldg(&L.rb[w]);
    const uint32_t bit = 1u << (c & 31);
    if (e.x & bit) {
        const uint32_t i = e.y + __popc(e.x & (bit - 1u));
        s0 = __ldg(&L.uoff[i]);
        s1 = __ldg(&L.uoff[i + 1]);
    }
}

// ucodes ascending: set the bit of every occurring code; the first code of a word leaves its rank (words without
// a bit are never asked for theirs)
__global__ void rb_build_kernel(const uint32_t *__restrict__ ucodes, uint32_t nuniq, uint2 *__restrict__ rb)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nuniq) return;
    const uint32_t c = ucodes[u], w = c >> 5;
    atomicOr(&rb[w].x, 1u << (c & 31));
    if (u == 0 || (ucodes[u - 1] >> 5) != w) rb[w].y = u;
}

// Dense EXCLUSIVE start table over the whole code space of one component (export / import of mco.index only):
//   dense[c] = #postings with code < c,  c in [0, space]   (space = 16^COMPONENT_SZ)
// Built as: zero, dense[ucode + 1] = end offset of that code's postings, inclusive max-scan (offsets ascend with
// the code) -- two streaming passes over the table instead of one short uncoalesced range per code.
__global__ void dense_mark_kernel(const uint32_t *__restrict__ ucodes, const uint32_t *__restrict__ uoff, uint32_t nuniq,
                                  uint32_t *__restrict__ dense)
{
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < nuniq) dense[(uint64_t)ucodes[u] + 1] = uoff[u + 1];
}

// mco.index.<c> as the reference writes it (co2mco.c:57-61): inclusive u64 prefix, chunked
__global__ void dense_incl64_kernel(const uint32_t *__restrict__ dense, uint64_t first, uint64_t count, uint64_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = dense[first + i + 1];
}

// import of a reference-written mco.index chunk: exclusive starts, and the checks that make it safe to use as offsets
// (non-decreasing, never beyond the posting count); `prev` = the entry before the chunk
__global__ void dense_from_incl64_kernel(const uint64_t *__restrict__ incl, uint64_t first, uint64_t count, uint64_t prev, uint64_t n_postings,
                                         uint32_t *__restrict__ dense, uint32_t *__restrict__ flag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const uint64_t v = incl[i], before = i ? incl[i - 1] : prev;
        if (v < before || v > n_postings) atomicOr(flag, 1u);
        dense[first + i + 1] = (uint32_t)(v > n_postings ? n_postings : v);
    }
    if (first == 0 && i == 0) dense[0] = 0;
}

// ------------------------------------------------------------------------------------------------
// Stage III.  Reference hot loop (command_dist.c:774-784): for every code of query k, walk its
// posting list and ++ct[k*R + gid].  Here one CTA owns one (query, reference-tile) strip of the
// count matrix in shared memory: postings are scattered with shared-memory atomics, and the strip
// is written to HBM exactly once, coalesced -- no global atomics and no separate memset pass.
// ------------------------------------------------------------------------------------------------
#ifndef KSSD_DIST_THREADS
#define KSSD_DIST_THREADS 256
#endif
constexpr int kDistThreads = KSSD_DIST_THREADS;

template <typename CT>   // uint16_t when every query sketch is < 65536 codes, else uint32_t
__global__ void __launch_bounds__(kDistThreads) dist_count_kernel(const uint32_t *__restrict__ qcodes, const uint64_t *__restrict__ qindex,
                                                                     const CodeLookup L, const uint32_t *__restrict__ mco,
                                                                     uint32_t n_ref, uint32_t tile_refs, uint32_t n_tiles,
                                                                     uint32_t *__restrict__ ct, int accumulate)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t q = blockIdx.x / n_tiles;
    const uint32_t t = blockIdx.x - q * n_tiles;
    const uint32_t r0 = t * tile_refs;                      // tile_refs is a multiple of 8
    const uint32_t r1 = min(r0 + tile_refs, n_ref);
    const uint32_t width = r1 - r0;
    // zero the strip, 16 bytes per store
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        const uint32_t vecs = (width * (uint32_t)sizeof(CT) + 15) / 16;
        for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint64_t qs = qindex[q], qe = qindex[q + 1];
    uint32_t *t32 = reinterpret_cast<uint32_t *>(smem_raw);
    for (uint64_t i = qs + threadIdx.x; i < qe; i += blockDim.x) {
        const uint32_t c = __ldg(&qcodes[i]);
        uint32_t s0, s1;
        code_lookup(L, c, s0, s1);
        // postings are fetched eight at a time before any atomic, so a thread keeps eight loads in flight
        for (uint32_t g = s0; g < s1; g += 8) {
            uint32_t r[8];
#pragma unroll
            for (int u = 0; u < 8; u++) r[u] = (g + u < s1) ? __ldg(&mco[g + u]) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t o = r[u] - r0;                    // 0xffffffff - r0 >= width: padding never lands
                if (o < width) {
                    if (sizeof(CT) == 4) atomicAdd(&t32[o], 1u);
                    // 16-bit counters packed two per word: a count never exceeds the query sketch size (< 65536 here)
                    else atomicAdd(&t32[o >> 1], (o & 1u) ? 0x10000u : 1u);
                }
            }
        }
    }
    __syncthreads();
    // write the strip once; 16-byte stores when the row slice is 16-byte aligned
    uint32_t *row = ct + (uint64_t)q * n_ref + r0;
    const bool vec_ok = ((n_ref & 3u) == 0) && !accumulate;
    const uint32_t w4 = vec_ok ? (width & ~3u) : 0;
    if (sizeof(CT) == 2) {
        const uint2 *src = reinterpret_cast<const uint2 *>(smem_raw);      // four 16-bit counters
        uint4 *dst = reinterpret_cast<uint4 *>(row);
        for (uint32_t i = threadIdx.x; i < w4 / 4; i += blockDim.x) {
            const uint2 v = src[i];
            dst[i] = make_uint4(v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16);
        }
    } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(smem_raw);
        uint4 *dst = reinterpret_cast<uint4 *>(row);
        for (uint32_t i = threadIdx.x; i < w4 / 4; i += blockDim.x) dst[i] = src[i];
    }
    const CT *tile = reinterpret_cast<const CT *>(smem_raw);
    for (uint32_t i = w4 + threadIdx.x; i < width; i += blockDim.x) {
        if (accumulate) row[i] += (uint32_t)tile[i];
        else row[i] = (uint32_t)tile[i];
    }
}

// L2-resident variant: one persistent CTA per SM owns one query ROW at a time.  The row (R x 4 B) is zeroed with
// 16-byte stores -- it lands in the 126 MB L2, 148 rows in flight are ~59 MB at R = 100k -- and the postings are
// added with global reductions (RED.ADD) that hit those L2 lines; the row reaches HBM once, by write-back.  No
// shared-memory strip, no reference tiles (every posting list is walked once), no copy-out phase.
constexpr int kDistRowThreads = 512;

__global__ void __launch_bounds__(kDistRowThreads) dist_count_rows_kernel(const uint32_t *__restrict__ qcodes, const uint64_t *__restrict__ qindex,
                                                                              const CodeLookup L, const uint32_t *__restrict__ mco,
                                                                              uint32_t n_qry, uint32_t n_ref, uint32_t *__restrict__ ct, int accumulate)
{
    for (uint32_t q = blockIdx.x; q < n_qry; q += gridDim.x) {
        uint32_t *row = ct + (uint64_t)q * n_ref;
        if (!accumulate) {
            // head up to a 16-byte boundary, body as uint4, tail
            const uint32_t mis = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(row) & 15)) & 15) / 4;
            const uint32_t head = min(mis, n_ref);
            if (threadIdx.x < head) row[threadIdx.x] = 0;
            uint4 *v = reinterpret_cast<uint4 *>(row + head);
            const uint32_t nv = (n_ref - head) / 4;
            for (uint32_t i = threadIdx.x; i < nv; i += blockDim.x) v[i] = make_uint4(0, 0, 0, 0);
            for (uint32_t i = head + nv * 4 + threadIdx.x; i < n_ref; i += blockDim.x) row[i] = 0;
            __syncthreads();
        }
        // a WARP walks 32 query codes at a time: every lane looks one code up, then the warp reads each posting
        // list with consecutive lanes (coalesced; lists average ~15 gids) -- eight lists in flight before any RED
        const uint64_t qs = qindex[q], qe = qindex[q + 1];
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (uint64_t base = qs + 32ull * wid; base < qe; base += 32ull * nw) {
            const uint64_t i = base + lane;
            uint32_t s0 = 0, s1 = 0;
            if (i < qe) {
                const uint32_t c = __ldg(&qcodes[i]);
                code_lookup(L, c, s0, s1);
            }
            const uint32_t cnt = (uint32_t)min((uint64_t)32, qe - base);
            for (uint32_t j0 = 0; j0 < cnt; j0 += 8) {
                uint32_t r[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t a = __shfl_sync(kFull, s0, (j0 + u) & 31), b = __shfl_sync(kFull, s1, (j0 + u) & 31);
                    r[u] = (j0 + u < cnt && a + lane < b) ? __ldg(&mco[a + lane]) : 0xffffffffu;
                }
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (r[u] != 0xffffffffu) atomicAdd(&row[r[u]], 1u);
#pragma unroll
                for (int u = 0; u < 8; u++) {          // lists longer than a warp (rare)
                    const uint32_t a = __shfl_sync(kFull, s0, (j0 + u) & 31), b = __shfl_sync(kFull, s1, (j0 + u) & 31);
                    if (j0 + u < cnt)
                        for (uint32_t g = a + 32 + lane; g < b; g += 32) atomicAdd(&row[__ldg(&mco[g])], 1u);
                }
            }
        }
    }
}

// Peer variant of the row kernel (reference index sharded by code range across GPUs): no zeroing here -- the owner of
// each query block did that -- and the row of query q lives in the memory of rank q / rows_per_block, reached through
// a peer mapping.  The RED.ADDs of remote rows travel over NVLink as they are issued, overlapped with the posting
// walk; nothing dense is ever exchanged.
constexpr int kMaxPeers = 16;
struct PeerRows { uint32_t *block[kMaxPeers]; };

__global__ void __launch_bounds__(kDistRowThreads) dist_count_peer_kernel(const uint32_t *__restrict__ qcodes, const uint64_t *__restrict__ qindex,
                                                                              const CodeLookup L, const uint32_t *__restrict__ mco,
                                                                              uint32_t n_qry, uint32_t n_ref, PeerRows rows, uint32_t rows_per_block)
{
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (uint32_t q = blockIdx.x; q < n_qry; q += gridDim.x) {
        const uint32_t owner = q / rows_per_block;
        uint32_t *row = rows.block[owner] + (uint64_t)(q - owner * rows_per_block) * n_ref;
        const uint64_t qs = qindex[q], qe = qindex[q + 1];
        for (uint64_t base = qs + 32ull * wid; base < qe; base += 32ull * nw) {
            const uint64_t i = base + lane;
            uint32_t s0 = 0, s1 = 0;
            if (i < qe) {
                const uint32_t c = __ldg(&qcodes[i]);
                code_lookup(L, c, s0, s1);
            }
            // most codes of a query fall outside this rank's code range: skip the empty lists without shuffling them
            uint32_t live = __ballot_sync(kFull, s1 > s0);
            while (live) {
                uint32_t r[4];
                int src[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    src[u] = live ? __ffs(live) - 1 : -1;
                    if (live) live &= live - 1;
                    r[u] = 0xffffffffu;
                    if (src[u] >= 0) {
                        const uint32_t a = __shfl_sync(kFull, s0, src[u]), b = __shfl_sync(kFull, s1, src[u]);
                        if (a + lane < b) r[u] = __ldg(&mco[a + lane]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (r[u] != 0xffffffffu) atomicAdd(&row[r[u]], 1u);
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (src[u] >= 0) {              // lists longer than a warp (rare)
                        const uint32_t a = __shfl_sync(kFull, s0, src[u]), b = __shfl_sync(kFull, s1, src[u]);
                        for (uint32_t g = a + 32 + lane; g < b; g += 32) atomicAdd(&row[__ldg(&mco[g])], 1u);
                    }
            }
        }
    }
}

// ---- statistics (reference output_ctrl, command_dist.c:1251-1287), double precision ----
struct StatParams {
    int metric, correction, kmerlen, dim_rd_len, skip_zero;
    double dthreshold;
    double cmprsn_num;    // (double)(llong)(uint32)(ref_num*qry_num), command_dist.c:1186
};

struct StatRow {   // mirrors kssd_stat_row_t
    uint32_t qry, ref, shared, rs_u, ref_size, qry_size;
    double metric, dist, pvalue, fdr, ci_m_lo, ci_m_hi, ci_d_lo, ci_d_hi;
};

__device__ __forceinline__ double get_matric(int metric_kind, double y)
{
    return metric_kind == 0 ? 1.0 / (2.0 * y) + 0.5 : 1.0 / y;
}

// (unsigned int)rs as gcc/x86-64 evaluates it at command_dist.c:1269 (cvttsd2si to 64 bits, low half kept):
// NaN and out-of-range doubles give 0x8000000000000000 -> 0
__device__ __forceinline__ uint32_t x86_double_to_u32(double v)
{
    if (!(fabs(v) < 9223372036854775808.0)) return 0u;
    return (uint32_t)(unsigned long long)(long long)v;
}

// keep / suppress decision only (first pass): no statistics, one log at most
__device__ __forceinline__ bool stat_keep(const StatParams &S, uint32_t X, uint32_t Y, uint32_t I)
{
    if (S.skip_zero && I == 0) return false;
    if (S.dthreshold >= 1.0) return true;          // dist is clamped to <= 1 (NaN never compares greater)
    double rs = 0.0;
    if (S.correction) {
        const uint32_t xo = X - I, yo = Y - I;
        const double base = 1.0 - 1.0 / pow(4.0, (double)(S.kmerlen - S.dim_rd_len));
        const double px = 1.0 - pow(base, (double)xo);
        const double py = 1.0 - pow(base, (double)yo);
        rs = px * py * (double)(xo + yo) / (px + py - 2.0 * px * py);
    }
    const uint32_t tmp = S.metric == 0 ? X + Y - I : (X < Y ? X : Y);
    const double m = ((double)I - rs) / (double)tmp;
    double dist = log(get_matric(S.metric, m)) / (double)S.kmerlen;
    if (dist > 1.0) dist = 1.0;
    return !(dist > S.dthreshold);
}

// returns false when the row is suppressed
__device__ __forceinline__ bool stat_row(const StatParams &S, uint32_t X, uint32_t Y, uint32_t I, StatRow &r)
{
    if (S.skip_zero && I == 0) return false;
    double rs = 0.0;
    if (S.correction) {
        const uint32_t xo = X - I, yo = Y - I;
        const double base = 1.0 - 1.0 / pow(4.0, (double)(S.kmerlen - S.dim_rd_len));
        const double px = 1.0 - pow(base, (double)xo);
        const double py = 1.0 - pow(base, (double)yo);
        rs = px * py * (double)(xo + yo) / (px + py - 2.0 * px * py);
    }
    const uint32_t tmp = S.metric == 0 ? X + Y - I : (X < Y ? X : Y);
    const double m = ((double)I - rs) / (double)tmp;
    double dist = log(get_matric(S.metric, m)) / (double)S.kmerlen;
    if (dist > 1.0) dist = 1.0;
    if (dist > S.dthreshold) return false;
    if (S.skip_zero && I == 0) return false;
    const double sd = sqrt(m * (1.0 - m) / (double)tmp);
    const double pv = 0.5 * erfc(m / sd * 0.70710678118654757);   // pow(0.5,0.5) rounded to double
    const double c1 = m - 1.96 * sd, c2 = m + 1.96 * sd;
    r.shared = I; r.rs_u = x86_double_to_u32(rs); r.ref_size = X; r.qry_size = Y;
    r.metric = m; r.dist = dist; r.pvalue = pv; r.fdr = pv * S.cmprsn_num;
    r.ci_m_lo = c1; r.ci_m_hi = c2;
    r.ci_d_lo = log(get_matric(S.metric, c2)) / (double)S.kmerlen;
    r.ci_d_hi = log(get_matric(S.metric, c1)) / (double)S.kmerlen;
    return true;
}

// Statistics.  One WARP owns one chunk of 1024 refs of one query row.
//   list pass  : read the chunk once (eight 16-byte loads per lane in flight), decide which cells are printed, reserve
//                room in the pair list with one atomicAdd per chunk and write the (query, ref) pairs, ref ascending;
//   (scan of the per-chunk counts gives every chunk its place in print order: query-major, refs ascending, exactly
//    as dist_print_nobin emits rows, command_dist.c:1228-1242)
//   rows pass  : one thread per printed row evaluates output_ctrl densely -- no lane idles through another lane's
//                erfc/log -- and stores the row at its final position.
// TRIVIAL = the keep decision needs no arithmetic (-D >= 1: dist is clamped to <= 1, NaN never compares greater).
constexpr int kStatThreads = 256;
constexpr int kStatRefsPerBlock = 1024;

template <bool TRIVIAL>
__device__ __forceinline__ bool stat_keep_t(const StatParams &S, uint32_t X, uint32_t Y, uint32_t I)
{
    if (TRIVIAL) return !(S.skip_zero && I == 0);
    return stat_keep(S, X, Y, I);
}

template <bool TRIVIAL>
__global__ void __launch_bounds__(kStatThreads) stats_list_kernel(const StatParams S, const uint32_t *__restrict__ ct,
                                                                   const uint32_t *__restrict__ qsz, const uint32_t *__restrict__ rsz,
                                                                   uint32_t n_ref, uint32_t chunks_per_row, uint64_t n_chunks,
                                                                   uint32_t *__restrict__ chunk_cnt, uint32_t *__restrict__ chunk_pos,
                                                                   unsigned long long *__restrict__ cursor, uint64_t cap, uint2 *__restrict__ pairs)
{
    uint64_t wb = ((uint64_t)blockIdx.x * kStatThreads + threadIdx.x) >> 5;     // this warp's chunk
    const bool live = wb < n_chunks;                 // dead warps of the last CTA still take part in the barriers
    if (!live) wb = n_chunks - 1;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = (uint32_t)(wb / chunks_per_row);
    const uint32_t b = (uint32_t)(wb - (uint64_t)q * chunks_per_row);
    const uint32_t Y = qsz[q];
    const uint32_t *row = ct + (uint64_t)q * n_ref;
    const uint32_t r0 = live ? b * kStatRefsPerBlock : n_ref;     // a dead warp sees an empty chunk
    // lane l holds refs r0 + 128*i + 4*l + {0..3}, i = 0..7 (vector path) -- ascending in (i, lane, component)
    uint32_t m[8];
    uint32_t kept = 0;
    const bool vec = (n_ref & 3u) == 0 && r0 + kStatRefsPerBlock <= n_ref;
    if (vec) {
        const uint4 *v = reinterpret_cast<const uint4 *>(row + r0);
        uint4 x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __ldg(&v[lane + 32 * i]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t r = r0 + 128 * i + 4 * lane;
            if (TRIVIAL) {
                m[i] = S.skip_zero ? ((x[i].x != 0) | ((x[i].y != 0) << 1) | ((x[i].z != 0) << 2) | ((x[i].w != 0) << 3)) : 15u;
            } else {
                m[i] = (uint32_t)stat_keep(S, rsz[r], Y, x[i].x) | ((uint32_t)stat_keep(S, rsz[r + 1], Y, x[i].y) << 1) |
                       ((uint32_t)stat_keep(S, rsz[r + 2], Y, x[i].z) << 2) | ((uint32_t)stat_keep(S, rsz[r + 3], Y, x[i].w) << 3);
            }
            kept += __popc(m[i]);
        }
    } else {
        // ragged tail / unaligned rows: lane l holds refs r0 + 32*j + l, bit j of m[j >> 2 ... ] -- simple scalar layout
#pragma unroll
        for (int i = 0; i < 8; i++) {
            m[i] = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t r = r0 + 32 * (4 * i + c) + lane;
                if (r < n_ref && stat_keep_t<TRIVIAL>(S, rsz[r], Y, row[r])) m[i] |= 1u << c;
            }
            kept += __popc(m[i]);
        }
    }
    const uint32_t total = __reduce_add_sync(kFull, kept);
    // one atomicAdd per CTA (eight chunks) reserves the room; warps take their share in chunk order
    __shared__ uint32_t wtot[kStatThreads / 32];
    __shared__ unsigned long long cta_base;
    const uint32_t wid = threadIdx.x >> 5;
    if (lane == 0) wtot[wid] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t sum = 0;
        for (int w = 0; w < kStatThreads / 32; w++) sum += wtot[w];
        cta_base = sum ? atomicAdd(cursor, (unsigned long long)sum) : 0ull;
    }
    __syncthreads();
    unsigned long long base = cta_base;
    for (uint32_t w = 0; w < wid; w++) base += wtot[w];
    if (lane == 0 && live) {
        chunk_cnt[wb] = total;
        chunk_pos[wb] = (uint32_t)base;
    }
    if (total == 0 || base + total > cap) return;
    uint64_t o = base;
    if (vec) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t cnt = __popc(m[i]);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            uint64_t w = o + incl - cnt;
            const uint32_t r = r0 + 128 * i + 4 * lane;
            if (m[i] & 1u) pairs[w++] = make_uint2(q, r);
            if (m[i] & 2u) pairs[w++] = make_uint2(q, r + 1);
            if (m[i] & 4u) pairs[w++] = make_uint2(q, r + 2);
            if (m[i] & 8u) pairs[w++] = make_uint2(q, r + 3);
            o += __shfl_sync(kFull, incl, 31);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const bool keep = (m[i] >> c) & 1u;
                const uint32_t bal = __ballot_sync(kFull, keep);
                if (keep) pairs[o + __popc(bal & ((1u << lane) - 1u))] = make_uint2(q, r0 + 32 * (4 * i + c) + lane);
                o += __popc(bal);
            }
    }
}

// one thread per listed pair evaluates the statistics and stores the row at its print-order position
// (chunk of the pair -> where that chunk's rows start + rank of the pair inside its chunk)
__global__ void __launch_bounds__(kStatThreads) stats_rows_kernel(const StatParams S, const uint32_t *__restrict__ ct,
                                                                   const uint32_t *__restrict__ qsz, const uint32_t *__restrict__ rsz,
                                                                   uint32_t n_ref, uint32_t chunks_per_row, const uint32_t *__restrict__ chunk_pos,
                                                                   const uint64_t *__restrict__ chunk_out, const uint2 *__restrict__ pairs,
                                                                   uint64_t n_pairs, StatRow *__restrict__ rows)
{
    const uint64_t i = (uint64_t)blockIdx.x * kStatThreads + threadIdx.x;
    if (i >= n_pairs) return;
    const uint2 p = pairs[i];
    const uint64_t wb = (uint64_t)p.x * chunks_per_row + p.y / kStatRefsPerBlock;
    StatRow out;
    stat_row(S, rsz[p.y], qsz[p.x], ct[(uint64_t)p.x * n_ref + p.y], out);   // kept by construction
    out.qry = p.x; out.ref = p.y;
    rows[chunk_out[wb] + (i - chunk_pos[wb])] = out;
}

// every cell is printed (-D >= 1, no skip_zero): rows land at q * R + r directly
__global__ void __launch_bounds__(kStatThreads) stats_rows_dense_kernel(const StatParams S, const uint32_t *__restrict__ ct,
                                                                         const uint32_t *__restrict__ qsz, const uint32_t *__restrict__ rsz,
                                                                         uint32_t n_ref, uint64_t n_cells, StatRow *__restrict__ rows)
{
    const uint64_t i = (uint64_t)blockIdx.x * kStatThreads + threadIdx.x;
    if (i >= n_cells) return;
    const uint32_t q = (uint32_t)(i / n_ref), r = (uint32_t)(i - (uint64_t)q * n_ref);
    StatRow out;
    stat_row(S, rsz[r], qsz[q], ct[i], out);
    out.qry = q; out.ref = r;
    rows[i] = out;
}

// ---- fused sparse Stage III: counts + filter + listing without the Q x R matrix ----
// For searches that can never print a zero-shared cell (skip_zero, or -D < 1 without --correction: I = 0 gives
// dist = 1 > D) only the cells some posting touches matter -- P increments instead of Q x R cells.  One CTA owns one
// query at a time: the gids of its postings are counted in a shared-memory hash table (CAS insert + ATOMS add), a
// shared bitmap of the touched refs gives the ascending-ref order dist_print_nobin needs without a sort, the keep
// rule of output_ctrl is applied in place, and the (query, ref, shared) hits are appended to one list; the rows pass
// then evaluates the statistics densely, one thread per hit.  All components of the index add into the same table.
struct SparseComp { const uint32_t *qcodes; const uint64_t *qindex; CodeLookup L; const uint32_t *mco; };
struct SparseHit { uint32_t q, r, shared; };
// Two shapes of the per-query CTA.  Wide: 512 threads, 8192 hash slots, 2048 query codes per pass -- queries that touch
// thousands of references.  Narrow: 128 threads, 2048 slots, 512 codes per pass -- the usual search (a query touches its
// relatives and a few chance hits), and every shard of a reference set split over GPUs: four times as many queries in
// flight per SM hide the dependent lookup -> posting chain, and the per-query fixed work (clearing and walking the table)
// is a quarter.
template <int THREADS_, uint32_t SLOTS_, uint32_t TILE_, int HBITS_>
struct SparseCfg {
    static constexpr int kThreads = THREADS_;
    static constexpr uint32_t kSlots = SLOTS_, kTile = TILE_;
    static constexpr uint32_t kMaxDistinct = SLOTS_ / 4 * 3;   // refs one query may touch before the dense path takes over
    static constexpr int kHashShift = 32 - HBITS_;
};
using SparseWide = SparseCfg<512, 8192, 2048, 13>;
using SparseNarrow = SparseCfg<128, 2048, 512, 11>;
constexpr uint32_t kSparseEmpty = 0xffffffffu;

// PACKED: one 32-bit word per slot, gid in the high bits and the count in the low `cb` bits (the host picks it when
// R and the largest query sketch leave room) -- the table is 32 KiB instead of 64 and a third CTA fits on the SM.
template <class C, bool PACKED>
__device__ __forceinline__ void sparse_insert(uint32_t *keys, uint32_t *vals, uint32_t cb, uint32_t *bitmap, uint32_t *distinct, uint32_t g)
{
    uint32_t h = (g * 0x9E3779B1u) >> C::kHashShift;
    for (uint32_t probes = 0; probes < C::kSlots; probes++) {
        // a ref shared with the query is hit once per shared code: most inserts find their key already there
        const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&keys[h]);
        if (PACKED) {
            if ((cur >> cb) == g) { atomicAdd(&keys[h], 1u); return; }          // (the empty pattern never decodes to a gid)
        } else {
            if (cur == g) { atomicAdd(&vals[h], 1u); return; }
        }
        const uint32_t old = atomicCAS(&keys[h], kSparseEmpty, PACKED ? ((g << cb) | 1u) : g);
        if (old == kSparseEmpty) {
            atomicAdd(distinct, 1u);
            atomicOr(&bitmap[g >> 5], 1u << (g & 31));
            if (!PACKED) atomicAdd(&vals[h], 1u);
            return;
        }
        if (PACKED) {
            if ((old >> cb) == g) { atomicAdd(&keys[h], 1u); return; }
        } else {
            if (old == g) { atomicAdd(&vals[h], 1u); return; }
        }
        h = (h + 1) & (C::kSlots - 1);
    }
    atomicAdd(distinct, C::kSlots);                        // table full: poison the tally, the query goes dense
}


// block-wide exclusive prefix of two values per thread (x, y); returns the prefixes, *total = the block sums (2 barriers)
template <class C>
__device__ __forceinline__ uint2 sparse_block_scan(uint2 v, uint2 *wsum, uint2 *total)
{
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint2 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tx = __shfl_up_sync(kFull, incl.x, o), ty = __shfl_up_sync(kFull, incl.y, o);
        if (lane >= (uint32_t)o) { incl.x += tx; incl.y += ty; }
    }
    __syncthreads();                                          // wsum may still be read from the previous scan
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint2 off = make_uint2(incl.x - v.x, incl.y - v.y), tot = make_uint2(0, 0);
#pragma unroll
    for (uint32_t w = 0; w < C::kThreads / 32; w++) {
        const uint2 s = wsum[w];
        off.x += w < wid ? s.x : 0u;
        off.y += w < wid ? s.y : 0u;
        tot.x += s.x;
        tot.y += s.y;
    }
    *total = tot;
    return off;
}

// The posting walk of one query (C::kThreads threads).  The dense row kernel keeps its warp-per-list walk: with the
// row zeroing and RED traffic in the way this one measured 1.55 ms against 1.43 ms there.
// Per tile of query codes: (A) every thread looks its codes up -- all the random reads of the tile are in flight at
// once -- and leaves (list start, prefix) descriptors of the non-empty lists in shared memory; (B) the tile's postings,
// numbered through a prefix sum of the list lengths, are split evenly over the warps, 32 consecutive postings per step
// (coalesced pieces of two or three lists), four steps of gid loads in flight ahead of emit(gid).
template <class C, typename F>
__device__ __forceinline__ void walk_query_postings(const uint32_t *__restrict__ qcodes, const uint64_t *__restrict__ qindex,
                                                    const CodeLookup L, const uint32_t *__restrict__ mco, uint32_t q,
                                                    uint32_t *lstart, uint32_t *lpre, uint2 *wsum, F emit)
{
    const uint64_t qs = qindex[q], qe = qindex[q + 1];
    for (uint64_t t0 = qs; t0 < qe; t0 += C::kTile) {
        const uint32_t nt = (uint32_t)min((uint64_t)C::kTile, qe - t0);
        constexpr uint32_t kPer = C::kTile / C::kThreads;           // consecutive codes per thread
        uint32_t st[kPer], len[kPer], sum = 0, live = 0;
#pragma unroll
        for (uint32_t j = 0; j < kPer; j++) {
            const uint32_t i = threadIdx.x * kPer + j;
            uint32_t s0 = 0, s1 = 0;
            if (i < nt) {
                const uint32_t c = __ldg(&qcodes[t0 + i]);
                code_lookup(L, c, s0, s1);
            }
            st[j] = s0;
            len[j] = s1 - s0;
            sum += len[j];
            live += len[j] != 0;
        }
        // only the non-empty lists get a descriptor (a query code no reference holds has an empty list)
        uint2 tot;
        const uint2 pre = sparse_block_scan<C>(make_uint2(live, sum), wsum, &tot);
        const uint32_t T = tot.y, nl = tot.x;
        {
            uint32_t slot = pre.x, p = pre.y;
#pragma unroll
            for (uint32_t j = 0; j < kPer; j++)
                if (len[j]) {
                    lstart[slot] = st[j];
                    lpre[slot] = p;
                    slot++;
                    p += len[j];
                }
        }
        if (threadIdx.x == 0) lpre[nl] = T;
        __syncthreads();
        // every WARP takes an equal run of the tile's postings, 32 consecutive ones per step: lane l finds the
        // list of posting pb + l by walking forward from the list of the step's first posting (a step spans two or
        // three lists), so the gid loads of a step fall into a few contiguous pieces
        constexpr uint32_t nw = C::kThreads / 32;
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const uint32_t wchunk = ((T + nw - 1) / nw + 31) & ~31u;
        const uint32_t pw0 = min(wid * wchunk, T), pw1 = min(pw0 + wchunk, T);
        if (pw0 < pw1) {
            uint32_t lo = 0, hi = nl;                                     // list holding posting pw0 (warp-uniform search)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (lpre[mid] <= pw0) lo = mid; else hi = mid;
            }
            uint32_t j0 = lo;
            auto fetch = [&](uint32_t pb, uint32_t &jl) -> uint32_t {
                const uint32_t p = pb + lane;
                uint32_t j = j0;
                uint32_t g = 0xffffffffu;
                if (p < pw1) {
                    while (lpre[j + 1] <= p) j++;
                    g = __ldg(&mco[lstart[j] + (p - lpre[j])]);
                }
                jl = j;
                return g;
            };
            // four steps of gid loads stay in flight ahead of the inserts (a first touch of a list is a DRAM miss)
            constexpr int kAhead = 4;
            uint32_t g[kAhead], jl;
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                g[u] = 0xffffffffu;
                if (pw0 + 32u * u < pw1) {
                    g[u] = fetch(pw0 + 32u * u, jl);
                    j0 = __shfl_sync(kFull, jl, 31);                      // lane 31 is live in every step but the last
                }
            }
            for (uint32_t pb = pw0; pb < pw1; pb += 32u * kAhead) {
#pragma unroll
                for (int u = 0; u < kAhead; u++) {
                    const uint32_t cur = g[u];
                    g[u] = 0xffffffffu;
                    const uint32_t nxt = pb + 32u * (u + kAhead);
                    if (nxt < pw1) {
                        g[u] = fetch(nxt, jl);
                        j0 = __shfl_sync(kFull, jl, 31);
                    }
                    if (cur != 0xffffffffu) emit(cur);
                }
            }
        }
        __syncthreads();                                                  // lstart / lpre are rewritten by the next tile
    }
}

template <class C, bool TRIVIAL, bool PACKED>
__global__ void __launch_bounds__(C::kThreads) dist_sparse_kernel(const SparseComp *__restrict__ comps, int n_comp, uint32_t n_qry, uint32_t n_ref, uint32_t cb,
                                                                     const StatParams S, const uint32_t *__restrict__ qsz,
                                                                     const uint32_t *__restrict__ rsz, uint32_t *__restrict__ q_cnt,
                                                                     unsigned long long *__restrict__ q_pos, unsigned long long *__restrict__ cursor,
                                                                     uint64_t cap, SparseHit *__restrict__ hits, uint32_t *__restrict__ over_n,
                                                                     uint32_t *__restrict__ over_list, uint32_t *__restrict__ err_flag)
{
    extern __shared__ __align__(16) uint32_t sparse_sm[];
    uint32_t *keys = sparse_sm, *vals = keys + C::kSlots, *lstart = vals + (PACKED ? 0 : C::kSlots), *lpre = lstart + C::kTile,
             *bitmap = lpre + C::kTile + 1;                  // PACKED: no vals array, the descriptors start right after the keys
    // touched refs in the words before a word, inside its thread's run of words: kept in the list descriptors' space, which is free once
    // the walk is over (when it fits: R <= 16 (2 kTile + 1) references; else the emission sums the run's words itself)
    uint16_t *wpfx = reinterpret_cast<uint16_t *>(lstart);
    const bool use_pfx = ((n_ref + 31) / 32) * 2 <= (2 * C::kTile + 1) * 4;
    const uint32_t cmask = PACKED ? ((1u << cb) - 1u) : 0u;
    __shared__ uint32_t distinct, tbase[C::kThreads];
    __shared__ uint2 wsum[C::kThreads / 32];
    __shared__ unsigned long long base_s;
    const uint32_t bw = (n_ref + 31) / 32;
    for (uint32_t q = blockIdx.x; q < n_qry; q += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < C::kSlots; i += C::kThreads) {
            keys[i] = kSparseEmpty;
            if (!PACKED) vals[i] = 0;
        }
        for (uint32_t i = threadIdx.x; i < bw; i += C::kThreads) bitmap[i] = 0;
        if (threadIdx.x == 0) {
            distinct = 0;
            if (PACKED) {      // a count shares its word with the ref id: the query must be as small as the host was told
                uint64_t codes = 0;
                for (int cc = 0; cc < n_comp; cc++) codes += comps[cc].qindex[q + 1] - comps[cc].qindex[q];
                if (codes + 1 >= (1ull << cb) - 1) atomicOr(err_flag, 1u);
            }
        }
        // ---- walk (walk_query_postings): every gid of every posting list of the query's codes goes into the table
        for (int cc = 0; cc < n_comp; cc++) {
            const SparseComp Cm = comps[cc];
            walk_query_postings<C>(Cm.qcodes, Cm.qindex, Cm.L, Cm.mco, q, lstart, lpre, wsum,
                                   [&](uint32_t g) { sparse_insert<C, PACKED>(keys, vals, cb, bitmap, &distinct, g); });
        }
        __syncthreads();
        if (distinct > C::kMaxDistinct) {                 // uniform: every thread reads the same shared word
            if (threadIdx.x == 0) { over_list[atomicAdd(over_n, 1u)] = q; q_cnt[q] = 0; q_pos[q] = 0; }   // handled densely by the host
            __syncthreads();
            continue;
        }
        // ---- emission in ascending ref order without a sort: a hit's place is the number of touched refs below it,
        // read off the bitmap; the threads go over the TABLE slots (hashing spreads the hits evenly over them)
        const uint32_t Y = qsz[q];
        constexpr uint32_t kSlotsPer = C::kSlots / C::kThreads;
        if (!TRIVIAL) {                                       // output_ctrl's keep rule, applied per touched cell
            for (uint32_t i = 0; i < kSlotsPer; i++) {
                const uint32_t sidx = threadIdx.x + i * C::kThreads, e = keys[sidx];
                const uint32_t r = PACKED ? e >> cb : e, cnt_r = PACKED ? e & cmask : vals[sidx];
                if (e != kSparseEmpty && !stat_keep(S, rsz[r], Y, cnt_r)) {
                    atomicAnd(&bitmap[r >> 5], ~(1u << (r & 31)));
                    keys[sidx] = kSparseEmpty;
                }
            }
            __syncthreads();
        }
        uint32_t per_log = 0;                                 // consecutive bitmap words per thread: a power of two
        while (((uint32_t)C::kThreads << per_log) < bw) per_log++;
        const uint32_t w0 = min(threadIdx.x << per_log, bw), w1 = min(w0 + (1u << per_log), bw);
        uint32_t cnt = 0;
        for (uint32_t w = w0; w < w1; w++) { if (use_pfx) wpfx[w] = (uint16_t)cnt; cnt += __popc(bitmap[w]); }      // (a run is < 2^16 refs)
        uint2 tot2;
        const uint32_t off = sparse_block_scan<C>(make_uint2(cnt, 0u), wsum, &tot2).x;
        const uint32_t total = tot2.x;
        tbase[threadIdx.x] = off;
        if (threadIdx.x == 0) {
            base_s = total ? atomicAdd(cursor, (unsigned long long)total) : 0ull;
            q_cnt[q] = total;
            q_pos[q] = base_s;
        }
        __syncthreads();
        const unsigned long long base = base_s;
        if (total && base + total <= cap) {
            for (uint32_t i = 0; i < kSlotsPer; i++) {
                const uint32_t sidx = threadIdx.x + i * C::kThreads, e = keys[sidx];
                if (e == kSparseEmpty) continue;
                const uint32_t r = PACKED ? e >> cb : e;
                const uint32_t w = r >> 5, owner = w >> per_log;
                uint32_t rank = tbase[owner] + __popc(bitmap[w] & ((1u << (r & 31)) - 1u));
                if (use_pfx) rank += wpfx[w];
                else for (uint32_t x = owner << per_log; x < w; x++) rank += __popc(bitmap[x]);
                SparseHit h;
                h.q = q; h.r = r; h.shared = PACKED ? e & cmask : vals[sidx];
                hits[base + rank] = h;
            }
        }
        __syncthreads();                                      // the table is cleared for the next query
    }
}

// ---- queries that overflowed the sparse table go through a small dense sub-job; these kernels move data in and out ----
// codes of the listed queries, compacted: out[sub_index[i] ..] = qcodes[qindex[list[i]] ..]
__global__ void gather_query_codes_kernel(const uint32_t *__restrict__ qcodes, const uint64_t *__restrict__ qindex, const uint32_t *__restrict__ list,
                                          const uint64_t *__restrict__ sub_index, uint32_t *__restrict__ out)
{
    const uint32_t i = blockIdx.x;
    const uint64_t a = qindex[list[i]], n = qindex[list[i] + 1] - a, o = sub_index[i];
    for (uint64_t j = threadIdx.x; j < n; j += blockDim.x) out[o + j] = qcodes[a + j];
}

__global__ void count_rows_by_query_kernel(const StatRow *__restrict__ rows, uint64_t n, uint32_t *__restrict__ cnt)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) atomicAdd(&cnt[rows[j].qry], 1u);
}

// rows of the sub-job (query-major, local query numbers) -> their place in the print order of the whole job
__global__ void place_sub_rows_kernel(const StatRow *__restrict__ sub_rows, uint64_t n, const uint32_t *__restrict__ list,
                                      const uint64_t *__restrict__ first, const uint64_t *__restrict__ q_out, StatRow *__restrict__ rows)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    StatRow r = sub_rows[j];
    const uint32_t lq = r.qry, q = list[lq];
    r.qry = q;
    rows[q_out[q] + (j - first[lq])] = r;
}

// same, launched before the host knows how many hits there are: the grid covers the hit list's capacity and the count is
// read from where the count kernel left it (kssd_dist_stats_async)
__global__ void __launch_bounds__(kStatThreads) stats_rows_sparse_dev_kernel(const StatParams S, const uint32_t *__restrict__ qsz,
                                                                              const uint32_t *__restrict__ rsz, const SparseHit *__restrict__ hits,
                                                                              const unsigned long long *__restrict__ n_hits_dev, uint64_t cap,
                                                                              const unsigned long long *__restrict__ q_pos, const uint64_t *__restrict__ q_out,
                                                                              StatRow *__restrict__ rows)
{
    const uint64_t i = (uint64_t)blockIdx.x * kStatThreads + threadIdx.x;
    const uint64_t n_hits = *n_hits_dev;
    if (i >= n_hits || n_hits > cap) return;
    const SparseHit h = hits[i];
    StatRow out;
    stat_row(S, rsz[h.r], qsz[h.q], h.shared, out);
    out.qry = h.q; out.ref = h.r;
    rows[q_out[h.q] + (i - q_pos[h.q])] = out;
}

__global__ void __launch_bounds__(kStatThreads) stats_rows_sparse_kernel(const StatParams S, const uint32_t *__restrict__ qsz,
                                                                          const uint32_t *__restrict__ rsz, const SparseHit *__restrict__ hits,
                                                                          uint64_t n_hits, const unsigned long long *__restrict__ q_pos,
                                                                          const uint64_t *__restrict__ q_out, StatRow *__restrict__ rows)
{
    const uint64_t i = (uint64_t)blockIdx.x * kStatThreads + threadIdx.x;
    if (i >= n_hits) return;
    const SparseHit h = hits[i];
    StatRow out;
    stat_row(S, rsz[h.r], qsz[h.q], h.shared, out);          // kept by construction
    out.qry = h.q; out.ref = h.r;
    rows[q_out[h.q] + (i - q_pos[h.q])] = out;
}

// -N: best n refs per query by the raw metric with the reference's insertion rule
// (command_dist.c:1212-1227): strictly greater than everything it passes, so ties keep the lower
// rid first and zero-metric refs are never listed.  One CTA per query; n <= 1024.
__global__ void __launch_bounds__(256) topn_kernel(const StatParams S, const uint32_t *__restrict__ ct, const uint32_t *__restrict__ qsz,
                                                   const uint32_t *__restrict__ rsz, uint32_t n_ref, int nmax,
                                                   uint32_t *__restrict__ row_counts, StatRow *__restrict__ rows)
{
    const uint32_t q = blockIdx.x;
    const uint32_t Y = qsz[q];
    __shared__ double best_m[256];
    __shared__ int best_r[256];
    __shared__ double last_m;
    __shared__ int last_r;
    __shared__ uint32_t nout;
    if (threadIdx.x == 0) { last_m = 1e300; last_r = -1; nout = 0; }
    __syncthreads();
    for (int it = 0; it < nmax; it++) {
        // next element in (metric desc, rid asc) order after (last_m, last_r), metric > 0
        double bm = 0.0;
        int br = -1;
        for (uint32_t r = threadIdx.x; r < n_ref; r += blockDim.x) {
            const uint32_t X = rsz[r], I = ct[(uint64_t)q * n_ref + r];
            const double m = S.metric == 1 ? (double)I / (double)(X < Y ? X : Y) : (double)I / (double)(X + Y - I);
            const bool after = (m < last_m) || (m == last_m && (int)r > last_r);
            if (!(m > 0.0) || !after) continue;
            if (br < 0 || m > bm || (m == bm && (int)r < br)) { bm = m; br = (int)r; }
        }
        best_m[threadIdx.x] = bm;
        best_r[threadIdx.x] = br;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            int r = -1;
            for (int i = 0; i < (int)blockDim.x; i++)
                if (best_r[i] >= 0 && (r < 0 || best_m[i] > m || (best_m[i] == m && best_r[i] < r))) { m = best_m[i]; r = best_r[i]; }
            last_m = m;
            last_r = r;
            if (r >= 0) {
                StatRow row;
                if (stat_row(S, rsz[r], Y, ct[(uint64_t)q * n_ref + r], row)) {
                    row.qry = q; row.ref = (uint32_t)r;
                    rows[(uint64_t)q * nmax + nout] = row;
                    nout++;
                }
            }
        }
        __syncthreads();
        if (last_r < 0) break;
    }
    if (threadIdx.x == 0) row_counts[q] = nout;
}


// -N on the sparse job: a reference that shares nothing with the query has metric 0 and is never listed, so the best n are among the
// cells the sparse count kernel touched -- the query's run of the hit list (ascending reference order) instead of a row of R cells.
// Same insertion rule as topn_kernel.
__global__ void __launch_bounds__(256) topn_sparse_kernel(const StatParams S, const SparseHit *__restrict__ hits, const unsigned long long *__restrict__ q_pos,
                                                          const uint32_t *__restrict__ q_cnt, const uint32_t *__restrict__ qsz, const uint32_t *__restrict__ rsz,
                                                          int nmax, uint32_t *__restrict__ row_counts, StatRow *__restrict__ rows)
{
    const uint32_t q = blockIdx.x;
    const uint32_t Y = qsz[q], cnt = q_cnt[q];
    const SparseHit *hq = hits + q_pos[q];
    __shared__ double best_m[256];
    __shared__ int best_i[256];
    __shared__ double last_m;
    __shared__ int last_r;
    __shared__ uint32_t nout;
    if (threadIdx.x == 0) { last_m = 1e300; last_r = -1; nout = 0; }
    __syncthreads();
    for (int it = 0; it < nmax; it++) {
        double bm = 0.0;
        int bi = -1, br = -1;
        for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
            const uint32_t r = hq[i].r, X = rsz[r], I = hq[i].shared;
            const double m = S.metric == 1 ? (double)I / (double)(X < Y ? X : Y) : (double)I / (double)(X + Y - I);
            const bool after = (m < last_m) || (m == last_m && (int)r > last_r);
            if (!(m > 0.0) || !after) continue;
            if (bi < 0 || m > bm || (m == bm && (int)r < br)) { bm = m; bi = (int)i; br = (int)r; }
        }
        best_m[threadIdx.x] = bm;
        best_i[threadIdx.x] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            int r = -1, sel = -1;
            for (int t = 0; t < (int)blockDim.x; t++) {
                if (best_i[t] < 0) continue;
                const int rt = (int)hq[best_i[t]].r;
                if (sel < 0 || best_m[t] > m || (best_m[t] == m && rt < r)) { m = best_m[t]; r = rt; sel = best_i[t]; }
            }
            last_m = m;
            last_r = r;
            if (sel >= 0) {
                StatRow row;
                if (stat_row(S, rsz[r], Y, hq[sel].shared, row)) {
                    row.qry = q; row.ref = (uint32_t)r;
                    rows[(uint64_t)q * nmax + nout] = row;
                    nout++;
                }
            }
        }
        __syncthreads();
        if (last_r < 0) break;
    }
    if (threadIdx.x == 0) row_counts[q] = nout;
}

// the listed rows of every query (up to nmax each, at q * nmax) -> one query-major list
__global__ void topn_gather_kernel(const StatRow *__restrict__ tmp, const uint32_t *__restrict__ row_counts, const uint64_t *__restrict__ row_off, int nmax,
                                   StatRow *__restrict__ rows)
{
    const uint32_t q = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < row_counts[q]; i += blockDim.x) rows[row_off[q] + i] = tmp[(uint64_t)q * nmax + i];
}

}  // namespace kssd
