// sketch_fastq.cuh -- Stage I for FASTQ text: fastq2co (reference iseq2comem.c:277-356) and the abundance
// variant mt_shortreads2koc (iseq2comem.c:554-615).
//
// Record semantics reproduced:
//   * the file is cut in lines at '\n' (fgets); record i = lines 4i..4i+3, bases come from line 4i+1,
//     qualities from line 4i+3; every record starts a fresh run (base = 1);
//   * fastq2co: a base counts iff it is ACGTacgt AND (signed char)qual[pos] >= Q (raw ASCII, no -33); anything
//     else breaks the run; record 0 is always processed, record i >= 1 only if its four lines are all
//     newline-terminated (the reference notices EOF while reading the record and stops: "last record is dropped
//     when the file has no trailing newline"); code 0 is kept;
//   * mt_shortreads2koc: quality ignored, a record is processed iff its four lines exist (the last may be
//     unterminated), occurrence counts saturate at 65535 (done in rle_kernel).
// Lines that do not fit the reference's fgets buffers (20000 / 4096 bytes) make the reference mis-frame records;
// such genomes are flagged (status bit 1) instead of imitated.
//
// B200 mapping: a line index (positions of every '\n', two streaming passes at HBM speed) turns the text into
// independent records; one THREAD walks one read with a rolling forward 2k-mer, 16-byte aligned loads, and the
// same shared-memory prefilter / exact sampled-set lookup as the FASTA kernel.  Reads are short, so there is no
// cross-thread state at all.
#pragma once
#include "sketch_scan.cuh"

namespace kssd {

constexpr int kNlBlock = 256;            // threads per block of the line-index passes
constexpr int kNlBytesPerThread = 64;     // four 16-byte chunks in a row: the scans and the block bookkeeping are paid per 64 B
constexpr int kNlBytesPerBlock = kNlBlock * kNlBytesPerThread;

// 4 bits: bit b = byte b of w equals the byte replicated in `pat` (exact zero-byte test of w ^ pat, then a
// multiply that gathers the four 0x80 flags into the top nibble)
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t pat)
{
    const uint32_t x = w ^ pat;
    const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    const uint32_t z = ~(t | x) & 0x80808080u;
    return ((z >> 7) * 0x10204080u) >> 28;
}

__device__ __forceinline__ uint32_t eq_mask16(const uint4 &v, uint32_t pat)
{
    return eq_mask4(v.x, pat) | (eq_mask4(v.y, pat) << 4) | (eq_mask4(v.z, pat) << 8) | (eq_mask4(v.w, pat) << 12);
}

__device__ __forceinline__ uint32_t nl_mask16(const uint8_t *seq, uint64_t a, uint64_t gs, uint64_t ge)
{
    // bit i = byte a+i is '\n' and lies inside [gs, ge); a is 16-byte aligned
    if (a + 16 <= gs || a >= ge) return 0;
    uint32_t m = eq_mask16(ldg_stream(reinterpret_cast<const uint4 *>(seq + a)), 0x0a0a0a0au);
    if (a < gs) m &= ~((1u << (gs - a)) - 1u);
    if (a + 16 > ge) m &= (1u << (ge - a)) - 1u;
    return m;
}

// Header starts of FASTA-formatted reads (reference reads2mco, iseq2comem.c:127-157): a '>' opens a record unless it
// lies inside a header line, i.e. unless another '>' precedes it on the same line ('\n' to '\n').  bit i = byte a+i
// opens a record.  The backward walk stops at the first '\n' or '>' -- one line at most.
__device__ __forceinline__ uint32_t hdr_mask16(const uint8_t *seq, uint64_t a, uint64_t gs, uint64_t ge)
{
    if (a + 16 <= gs || a >= ge) return 0;
    uint32_t gt = eq_mask16(ldg_stream(reinterpret_cast<const uint4 *>(seq + a)), 0x3e3e3e3eu);
    if (a < gs) gt &= ~((1u << (gs - a)) - 1u);
    if (a + 16 > ge) gt &= (1u << (ge - a)) - 1u;
    uint32_t m = 0;
    while (gt) {
        const int i = __ffs(gt) - 1;
        gt &= gt - 1;
        bool opens = true;
        for (uint64_t p = a + i; p > gs;) {
            const uint8_t b = seq[--p];
            if (b == '\n') break;
            if (b == '>') { opens = false; break; }
        }
        m |= (uint32_t)opens << i;
    }
    return m;
}

template <bool HDR>
__device__ __forceinline__ uint32_t line_mask16(const uint8_t *seq, uint64_t a, uint64_t gs, uint64_t ge)
{
    return HDR ? hdr_mask16(seq, a, gs, ge) : nl_mask16(seq, a, gs, ge);
}

// 64-bit mask of the thread's 64 bytes (bit i = byte a+i is a line end / a header start inside [gs, ge))
template <bool HDR>
__device__ __forceinline__ uint64_t line_mask64(const uint8_t *seq, uint64_t a, uint64_t gs, uint64_t ge)
{
    const uint32_t m0 = line_mask16<HDR>(seq, a, gs, ge), m1 = line_mask16<HDR>(seq, a + 16, gs, ge);
    const uint32_t m2 = line_mask16<HDR>(seq, a + 32, gs, ge), m3 = line_mask16<HDR>(seq, a + 48, gs, ge);
    return (uint64_t)(m0 | (m1 << 16)) | ((uint64_t)(m2 | (m3 << 16)) << 32);
}

template <bool HDR = false>
__global__ void __launch_bounds__(kNlBlock) nl_count_kernel(const uint8_t *__restrict__ seq, uint64_t gs, uint64_t ge, uint64_t a0,
                                                             uint32_t *__restrict__ block_counts)
{
    const uint64_t a = a0 + (uint64_t)blockIdx.x * kNlBytesPerBlock + (uint64_t)kNlBytesPerThread * threadIdx.x;
    uint32_t c = __popcll(line_mask64<HDR>(seq, a, gs, ge));
    c = __reduce_add_sync(kFull, c);
    __shared__ uint32_t red[kNlBlock / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int i = 0; i < kNlBlock / 32; i++) s += red[i];
        block_counts[blockIdx.x] = s;
    }
}

template <bool HDR = false>
__global__ void __launch_bounds__(kNlBlock) nl_fill_kernel(const uint8_t *__restrict__ seq, uint64_t gs, uint64_t ge, uint64_t a0,
                                                            const uint32_t *__restrict__ block_offsets, uint64_t *__restrict__ nlpos)
{
    const uint64_t a = a0 + (uint64_t)blockIdx.x * kNlBytesPerBlock + (uint64_t)kNlBytesPerThread * threadIdx.x;
    uint64_t m = line_mask64<HDR>(seq, a, gs, ge);
    const uint32_t c = __popcll(m);
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    __shared__ uint32_t wsum[kNlBlock / 32];
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint32_t base = block_offsets[blockIdx.x];
#pragma unroll
    for (uint32_t w = 0; w < kNlBlock / 32; w++) base += w < wid ? wsum[w] : 0u;
    uint32_t o = base + incl - c;
    while (m) {
        const int i = __ffsll((long long)m) - 1;
        m &= m - 1;
        nlpos[o++] = a + i;
    }
}

// ---- single-pass line index (decoupled look-back): ONE read of the text gives the position of every line end ----
// The two passes above read the text twice (count, then fill) with a scan between them.  Here a block takes a ticket (so
// that blocks start in text order), finds its line ends, publishes their number, adds up what the blocks before it
// published (aggregates, or an inclusive prefix as soon as one is there) and writes its positions -- 32-bit offsets
// from a0, which is all a file below 4 GiB needs.  state[b]: bits 62-63 = 0 nothing yet / 1 aggregate / 2 inclusive
// prefix, low 32 bits = the count.  A block covers 128 KiB (512 bytes per thread in eight coalesced steps, the masks stay in registers
// between the count and the write).  If the index does not fit `cap` entries the overflow flag is raised and the host
// falls back to the two-pass index.
constexpr unsigned long long kNlAgg = 1ull << 62, kNlPrefix = 2ull << 62;
constexpr int kNlxPerThread = 512;                            // bytes per thread of the single-pass index: 8 masks of 64 bits
constexpr int kNlxBytesPerBlock = kNlBlock * kNlxPerThread;   // 128 KiB per block: ~10^4 blocks per GB, so that a block's look-back
                                                              // (one L2 round trip per 32 predecessors) keeps up with the text rate
struct NlIndexOut { unsigned long long n_nl; uint32_t overflow, ticket, highbit, pad; };   // highbit: some byte of the file is >= 0x80

__global__ void __launch_bounds__(kNlBlock, 4) nl_index_kernel(const uint8_t *__restrict__ seq, uint64_t gs, uint64_t ge, uint64_t a0, uint32_t nblk,
                                                             unsigned long long *__restrict__ state, NlIndexOut *__restrict__ out,
                                                             uint32_t *__restrict__ nlpos32, uint32_t cap, uint32_t *__restrict__ any_overflow, int dbg_no_lookback)
{
    __shared__ uint32_t s_blk, s_prefix, wsum[kNlBlock / 32];
    if (threadIdx.x == 0) s_blk = atomicAdd(&out->ticket, 1u);
    __syncthreads();
    const uint32_t b = s_blk;
    // a warp owns 16 contiguous KiB of the block's 128; in step k its lanes read 64 adjacent bytes each (2 KiB per warp and step,
    // coalesced); line ends are numbered in text order: warp, then step, then lane
    constexpr int kSteps = kNlxPerThread / 64;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t a = a0 + (uint64_t)b * kNlxBytesPerBlock + (uint64_t)wid * (32 * kNlxPerThread) + 64ull * lane;
    // line ends of the thread's 16 groups of 32 bytes, one word per group in WORD-MAJOR bit order: bit 8j + k <-> byte j of word k of the
    // group, i.e. text position 4k + j.  That order costs one shift and a third of an OR per word to build (the exact 0x80-per-matching-
    // byte flags of the eight words, shifted by 7 - k and ORed) where the text-order gather cost a multiply, a shift and an OR; counting
    // does not care, and the few words that hold a line end are put into text order when the positions are written.
    uint32_t W[2 * kSteps];
    uint32_t pre[kSteps];                                    // line ends of the warp before this lane's 64 bytes of step k
    uint32_t wtot = 0;
    const uint64_t blk0 = a0 + (uint64_t)b * kNlxBytesPerBlock;
    uint32_t hb = 0;                                          // OR of every byte: a quality byte can only fail -Q <= 0 with its high bit set
    auto zflags = [](uint32_t w) -> uint32_t {                // 0x80 in every byte of w that equals '\n' (exact)
        const uint32_t x = w ^ 0x0a0a0a0au;
        return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
    };
    auto group_word = [&](const uint4 &lo, const uint4 &hi) -> uint32_t {
        return (zflags(lo.x) >> 7) | (zflags(lo.y) >> 6) | (zflags(lo.z) >> 5) | (zflags(lo.w) >> 4) | (zflags(hi.x) >> 3) | (zflags(hi.y) >> 2) |
               (zflags(hi.z) >> 1) | zflags(hi.w);
    };
    if (blk0 >= gs && blk0 + kNlxBytesPerBlock <= ge) {       // a block inside the file: no guards, two steps of loads in flight at a time
#pragma unroll
        for (int k = 0; k < kSteps; k += 2) {
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = ldg_stream(reinterpret_cast<const uint4 *>(seq + a + 2048ull * (k + (j >> 2)) + 16ull * (j & 3)));
#pragma unroll
            for (int j = 0; j < 8; j++) hb |= v[j].x | v[j].y | v[j].z | v[j].w;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                W[2 * (k + h)] = group_word(v[4 * h], v[4 * h + 1]);
                W[2 * (k + h) + 1] = group_word(v[4 * h + 2], v[4 * h + 3]);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < kSteps; k++) {                    // (first and last block of a file only: text-order masks, re-ordered bit by bit)
            const uint64_t m = line_mask64<false>(seq, a + 2048ull * k, gs, ge);
            uint32_t w2[2] = {0u, 0u};
            for (uint64_t t = m; t; t &= t - 1) {
                const int i = __ffsll((long long)t) - 1, g = i >> 5, q = i & 31;
                w2[g] |= 1u << (8 * (q & 3) + (q >> 2));
            }
            W[2 * k] = w2[0];
            W[2 * k + 1] = w2[1];
        }
        for (int k = 0; k < kSteps; k++)
            for (int j = 0; j < 64; j++) {
                const uint64_t p = a + 2048ull * k + j;
                if (p >= gs && p < ge) hb |= seq[p];
            }
    }
    if (__any_sync(kFull, (hb & 0x80808080u) != 0) && lane == 0) out->highbit = 1u;
#pragma unroll
    for (int k = 0; k < kSteps; k++) {
        const uint32_t c = __popc(W[2 * k]) + __popc(W[2 * k + 1]);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        pre[k] = wtot + incl - c;
        wtot += __shfl_sync(kFull, incl, 31);
    }
    if (lane == 0) wsum[wid] = wtot;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (uint32_t w = 0; w < kNlBlock / 32; w++) { base += w < wid ? wsum[w] : 0u; total += wsum[w]; }
    if (wid == 0) {
        volatile unsigned long long *st = state;
        if (lane == 0) { st[b] = (b == 0 ? kNlPrefix : kNlAgg) | total; __threadfence(); }
        uint32_t excl = 0;
        if (b > 0 && !dbg_no_lookback) {
            int64_t j = (int64_t)b - 1;                      // lane l looks at block j - l
            for (;;) {
                const int64_t mine = j - lane;
                unsigned long long v = kNlPrefix;            // before the first block: an empty prefix
                if (mine >= 0) { do { v = st[mine]; } while ((v >> 62) == 0); }
                const uint32_t pm = __ballot_sync(kFull, (v >> 62) == 2);
                const uint32_t upto = pm ? (uint32_t)__ffs(pm) : 32u;             // lanes 0 .. upto-1 count (the first prefix included)
                uint32_t val = lane < upto ? (uint32_t)v : 0u;
                val = __reduce_add_sync(kFull, val);
                excl += val;
                if (pm) break;
                j -= 32;
            }
            if (lane == 0) { __threadfence(); st[b] = kNlPrefix | (unsigned long long)(excl + total); }
        }
        if (lane == 0) {
            s_prefix = excl;
            if (b == nblk - 1) out->n_nl = (unsigned long long)excl + total;
            if ((unsigned long long)excl + total > cap) { out->overflow = 1u; *any_overflow = 1u; }
        }
    }
    __syncthreads();
    const uint32_t o0 = s_prefix + base;
#pragma unroll
    for (int k = 0; k < kSteps; k++) {
        uint32_t o = o0 + pre[k];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t w = W[2 * k + h];
            if (!w) continue;
            uint32_t T = 0;                                   // the group's line ends in text order
            for (; w; w &= w - 1) {
                const int bb = __ffs(w) - 1;
                T |= 1u << (4 * (bb & 7) + (bb >> 3));
            }
            for (; T; T &= T - 1) {
                const int i = __ffs(T) - 1;
                if (o < cap) nlpos32[o] = (uint32_t)(a + 2048ull * k + 32u * h + i - a0);
                o++;
            }
        }
    }
}

struct FastqArgs {
    const uint8_t *seq;
    uint64_t seq_bytes;          // readable bytes of the batch buffer
    uint64_t gs, ge;             // genome extent
    const uint64_t *nlpos;       // positions of the newline-terminated lines' '\n' (ascending)
    const uint32_t *nlpos32;     // or (single-pass index): 32-bit offsets from pos_base, the counts read from *idx on the device
    uint64_t pos_base;
    const NlIndexOut *idx;
    int warp_walk;               // 1: sketch_fastq3_kernel walks this file unless a quality byte can fail (then this kernel does)
    uint64_t n_nl;               // newline-terminated lines
    uint64_t n_lines;            // n_nl + 1 if an unterminated tail line exists
    uint64_t n_records;          // ceil(n_lines / 4)
    uint32_t gid;
    int abund;                   // 0 = fastq2co rules, 1 = mt_shortreads2koc rules
    int Q;
    uint32_t line_cap;           // 19999 / 4095: longest line the reference's fgets keeps in one piece
    uint64_t *out_keys;
    uint64_t *out_ords;
    uint32_t out_cap;
    uint32_t *out_count;
    int32_t *gstatus;
};

__device__ __forceinline__ uint32_t byte_at(const uint8_t *seq, uint64_t p, uint64_t &chunk_addr, uint4 &chunk)
{
    const uint64_t a = p & ~15ull;
    if (a != chunk_addr) { chunk = __ldg(reinterpret_cast<const uint4 *>(seq + a)); chunk_addr = a; }
    const uint32_t i = (uint32_t)(p - a);
    const uint32_t w = i < 8 ? (i < 4 ? chunk.x : chunk.y) : (i < 12 ? chunk.z : chunk.w);
    return (w >> (8 * (i & 3))) & 0xffu;
}

constexpr int kFastqThreads = 1024;

// exact test of one forward 2k-mer ending at byte offset `ord` of its genome (the ~1/2000 that pass both filters)
__device__ __forceinline__ void fastq_resolve(const SketchParams &P, const FastqArgs &A, uint64_t kmer, uint64_t ord)
{
    const uint64_t rc = revcomp2(kmer, P.TL);
    const uint64_t u = kmer < rc ? kmer : rc;
    const uint32_t inner = (uint32_t)(u >> (2 * P.out)) & P.innermask;
    uint32_t h = mix32(inner) & P.ht_mask, pfv = 0;
    bool found = false;
    for (;;) {
        const uint2 e = __ldg(&P.ht[h]);
        if (e.x == inner) { found = true; pfv = e.y; break; }
        if (e.x == kHtEmpty) break;
        h = (h + 1) & P.ht_mask;
    }
    if (!found) return;
    const uint64_t dr = (((u & P.undomask) + ((u & P.outmask) << (4 * P.s))) >> (4 * P.L)) + pfv;
    const uint32_t o = atomicAdd(A.out_count, 1u);
    if (o < A.out_cap) {
        A.out_keys[o] = ((dr & P.comp_mask) << 56) | ((uint64_t)A.gid << 28) | (dr >> P.comp_code_bits);
        A.out_ords[o] = ord;
    }
}

// bit i of the result = byte i of the 16 (four words, byte 0 first) is non-zero
__device__ __forceinline__ uint32_t nonzero_bytes16(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    auto nib = [](uint32_t b) -> uint32_t {
        const uint32_t z = (((b & 0x7f7f7f7fu) + 0x7f7f7f7fu) | b) & 0x80808080u;     // 0x80 per non-zero byte, exact
        return (((z >> 7) * 0x00204081u) >> 21) & 0xfu;
    };
    return nib(b0) | (nib(b1) << 4) | (nib(b2) << 8) | (nib(b3) << 12);
}

__global__ void __launch_bounds__(kFastqThreads, 1) sketch_fastq_kernel(const SketchParams P, const FastqArgs A)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *pf = reinterpret_cast<uint32_t *>(smem_raw);
    if (A.nlpos32 && A.warp_walk && (A.abund || A.Q <= -128 || !A.idx->highbit)) return;      // sketch_fastq3_kernel has walked this file
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.prefilter);
        uint4 *dst = reinterpret_cast<uint4 *>(pf);
        for (uint32_t i = threadIdx.x; i < (kPfWords + kPf2Words) / 4; i += blockDim.x) dst[i] = __ldg(&src[i]);
    }
    __syncthreads();
    const int TL = P.TL;
    // the line index: 64-bit positions and host-known counts (two-pass index), or 32-bit offsets with the counts left on
    // the device by nl_index_kernel (no host round trip between the index and this walk)
    const bool idx32 = A.nlpos32 != nullptr;
    if (idx32 && A.idx->overflow) return;                                  // the host redoes this file with the two-pass index
    const uint64_t n_nl = idx32 ? (uint64_t)A.idx->n_nl : A.n_nl;
    const uint64_t n_lines = idx32 ? n_nl + (A.seq[A.ge - 1] != '\n' ? 1u : 0u) : A.n_lines;
    const uint64_t n_records = idx32 ? (n_lines + 3) / 4 : A.n_records;
    auto NLP = [&](uint64_t i) -> uint64_t { return idx32 ? A.pos_base + A.nlpos32[i] : A.nlpos[i]; };
    // -Q <= 0: a quality byte fails only with its high bit set; the single-pass index has looked at every byte of the file, and
    // when none has it the quality lines are not read at all
    const bool q_can_fail = A.Q > 0 || !idx32 || A.idx->highbit != 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_records; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t l_seq = 4 * r + 1, l_q = 4 * r + 3;
        if (l_seq >= n_lines) continue;                                    // no sequence line at all
        bool process;
        if (A.abund) process = 4 * r + 4 <= n_lines;                       // four lines exist (iseq2comem.c:567)
        else process = (r == 0) || (4 * r + 4 <= n_nl);                    // read without touching EOF (:300-307)
        if (!process) continue;
        const uint64_t s0 = NLP(l_seq - 1) + 1;                            // l_seq >= 1: previous line is terminated
        const uint64_t s1 = l_seq < n_nl ? NLP(l_seq) : A.ge;              // end of bases ('\n' position or text end)
        uint64_t q0 = 0, qlen = 0;                                          // quality line incl. its '\n'
        if (l_q < n_lines) {
            q0 = NLP(l_q - 1) + 1;
            qlen = l_q < n_nl ? NLP(l_q) + 1 - q0 : A.ge - q0;
        }
        {   // every line of the record must fit the reference's fgets buffer, or its framing (and ours) is off
            const uint64_t h0 = r == 0 ? A.gs : NLP(4 * r - 1) + 1;
            const uint64_t hlen = NLP(4 * r) - h0;
            const uint64_t plen = (l_seq + 1 < n_nl) ? NLP(l_seq + 1) - (NLP(l_seq) + 1) : 0;
            if (s1 - s0 > A.line_cap || hlen > A.line_cap || plen > A.line_cap || qlen > (uint64_t)A.line_cap + 1) {
                atomicOr(&A.gstatus[A.gid], 2);
                continue;
            }
        }
        const bool use_q = !A.abund && A.Q > -128;
        if (TL >= 16) {
            // ---- vectorised walk: 16 aligned bytes at a time (one LDG.128 per chunk, the next one requested before this
            // one is processed), same classify / pack / probe code as the FASTA kernel.  With 2k >= 16 no k-mer can both
            // start after an invalid byte of a chunk and end inside that chunk, so the chunk's valid k-mer ends are the
            // positions before its first invalid byte (given enough run before it).
            uint32_t hist0 = 0, hist1 = 0, run = 0;        // last 32 bases (newest low), valid bases since the last break
            const uint32_t qrep = (uint32_t)(A.Q & 0xff) * 0x01010101u;
            const uint64_t last16 = (A.seq_bytes - 1) & ~15ull;                       // last readable aligned chunk
            auto ld16 = [&](uint64_t addr) -> uint4 { return __ldg(reinterpret_cast<const uint4 *>(A.seq + (addr < last16 ? addr : last16))); };
            const uint64_t a0 = s0 & ~15ull;
            const bool have_q = use_q && qlen > 0 && q_can_fail;
            // quality bytes of the chunk at a: [q0 + (a - s0), +16) -- unaligned; the two aligned chunks covering it
            const uint64_t qa0 = have_q ? ((q0 - (s0 - a0)) & ~15ull) : 0;
            const uint32_t qsh = have_q ? (uint32_t)((q0 - (s0 - a0)) & 15) : 0;     // same for every chunk of the read
            uint4 cnext = ld16(a0), qx = make_uint4(0, 0, 0, 0), qnext = make_uint4(0, 0, 0, 0);
            if (have_q) { qx = ld16(qa0); qnext = ld16(qa0 + 16); }
            for (uint64_t a = a0; a < s1; a += 16) {
                const uint4 c = cnext;
                if (a + 16 < s1) cnext = ld16(a + 16);
                const bool edge = a < s0 || s1 - a < 16;
                uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0, t0, t1, t2, t3, m0, m1, m2, m3;
                classify4(c.x, d0, t0, m0);
                classify4(c.y, d1, t1, m1);
                classify4(c.z, d2, t2, m2);
                classify4(c.w, d3, t3, m3);
                const uint32_t codes = prmt(prmt(m3, m2, 0x0073u), prmt(m1, m0, 0x0073u), 0x5410u);
                // invalid: anything but ACGTacgt (line ends included: '\r' breaks a read, iseq2comem.c:311-319)
                uint32_t inv = 0;
                if (((d0 | d1 | d2 | d3) | ((t0 | t1 | t2 | t3) & 0x04040404u)) != 0u || edge) {
                    const int lo = a < s0 ? (int)(s0 - a) : 0;
                    const int hi = s1 - a < 16 ? (int)(s1 - a) : 16;
                    inv = nonzero_bytes16(d0 | (t0 & 0x04040404u), d1 | (t1 & 0x04040404u), d2 | (t2 & 0x04040404u), d3 | (t3 & 0x04040404u));
                    inv |= ~(((1u << hi) - 1u) & ~((1u << lo) - 1u)) & 0xffffu;
                }
                if (have_q) {
                    const uint4 qy = qnext;
                    const uint64_t qa = qa0 + (a - a0);
                    if (a + 16 < s1) qnext = ld16(qa + 32);
                    // Q <= 0: only bytes with the high bit set (negative as signed char) can fail -- none in ASCII text
                    if (A.Q > 0 || ((qx.x | qx.y | qx.z | qx.w | qy.x | qy.y | qy.z | qy.w) & 0x80808080u) != 0u) {
                        const uint32_t w[8] = {qx.x, qx.y, qx.z, qx.w, qy.x, qy.y, qy.z, qy.w};
                        const uint32_t ws = qsh >> 2, bs = 8 * (qsh & 3);
                        uint32_t f[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            uint32_t l = 0, h = 0;
#pragma unroll
                            for (int j = 0; j < 4; j++)
                                if (ws == (uint32_t)j) { l = w[i + j]; h = w[i + j + 1 < 8 ? i + j + 1 : 7]; }
                            f[i] = __vcmpges4(__funnelshift_r(l, h, bs), qrep);          // 0xff / 0x00 per byte
                        }
                        auto nib = [](uint32_t v) -> uint32_t { return ((v & 0x08040201u) * 0x01010101u) >> 24; };
                        uint32_t pass = nib(f[0]) | (nib(f[1]) << 4) | (nib(f[2]) << 8) | (nib(f[3]) << 12);
                        const int64_t qn = (int64_t)qlen - ((int64_t)a - (int64_t)s0);   // bytes of this chunk inside the quality line
                        const uint32_t inq = qn >= 16 ? 0xffffu : (qn <= 0 ? 0u : ((1u << qn) - 1u));
                        pass = (pass & inq) | (A.Q <= 0 ? (~inq & 0xffffu) : 0u);       // beyond the line a quality counts as 0
                        inv |= ~pass & 0xffffu;
                        if (A.Q > 127) inv = 0xffffu;                                   // no signed byte reaches Q
                    }
                    qx = qy;
                } else if (use_q && A.Q > 0) {
                    inv = 0xffffu;                                                       // no quality line: every quality counts as 0
                }
                const int fi = inv ? __ffs(inv) - 1 : 16;                             // k-mers may end at bytes [0, fi)
                const int need = TL - 1 - (int)run;                                   // ... and at byte >= need
                if (fi > need) {
                    uint32_t ends = (fi >= 16 ? 0xffffu : ((1u << fi) - 1u)) & ~(need > 0 ? ((1u << need) - 1u) : 0u);
                    // W = history : this chunk's 16 codes (byte 15 newest, in the low bits)
                    const uint32_t X0 = __funnelshift_r(codes, hist0, 2 * P.out);
                    const uint32_t X1 = __funnelshift_r(hist0, hist1, 2 * P.out);
                    auto tsh = [&](int e) -> uint32_t {
                        return e < 0 ? (X0 << 2) : (e < 16 ? __funnelshift_r(X0, X1, 2 * e) : (X1 >> (2 * e - 32)));
                    };
                    uint32_t cand = 0;
#pragma unroll
                    for (int d = 15; d >= 0; d--) {
                        const uint32_t word = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(pf) + (tsh(d - 1) & (kPfWordMask << 2)));
                        cand = __funnelshift_l(__funnelshift_l(0u, word, tsh(d + 8)), cand, 1);
                    }
                    cand &= __brev(ends) >> 16;                                       // bit d <-> byte 15 - d
                    while (cand) {
                        const int d = __ffs(cand) - 1;
                        cand &= cand - 1;
                        const uint32_t klo = __funnelshift_r(codes, hist0, 2 * d), khi = __funnelshift_r(hist0, hist1, 2 * d);
                        if (!pf2_probe(pf, __funnelshift_r(klo, khi, 2 * P.out) & P.innermask)) continue;
                        fastq_resolve(P, A, (((uint64_t)khi << 32) | klo) & P.tupmask, a + (15 - d) - A.gs);
                    }
                }
                // state after the chunk
                if (inv == 0) run = min(run + 16u, 64u);
                else run = (uint32_t)__clz(inv << 16);                                // valid bytes after the last invalid one
                hist1 = hist0;
                hist0 = codes;
            }
        } else {
            // ---- scalar walk (2k < 16: a k-mer can start and end inside one chunk) ----
            uint64_t fwd = 0, ca = ~0ull, qa = ~0ull;
            uint4 cc = make_uint4(0, 0, 0, 0), qc = make_uint4(0, 0, 0, 0);
            uint32_t run = 0;
            for (uint64_t p = s0; p < s1; p++) {
                const uint32_t b = byte_at(A.seq, p, ca, cc);
                const uint32_t l = b | 0x20u;
                bool ok = (l == 'a' || l == 'c' || l == 'g' || l == 't');
                if (ok && use_q) {
                    const uint64_t off = p - s0;
                    const int q = off < qlen ? (int)(int8_t)byte_at(A.seq, q0 + off, qa, qc) : 0;
                    ok = q >= A.Q;
                }
                if (!ok) { run = 0; continue; }
                const uint32_t t = (b >> 1) & 3u;
                fwd = (fwd << 2) | (t ^ (t >> 1));
                if (++run < (uint32_t)TL) continue;
                if (!(pf_probe(pf, (uint32_t)(fwd >> (2 * P.out))) >> 31)) continue;
                fastq_resolve(P, A, fwd & P.tupmask, p - A.gs);
            }
        }
    }
}

// ---- --byread (reference reads2mco, iseq2comem.c:78-186): every occurrence kept, stream order, per-record index ----
// occurrence keys from the scan: (comp << 56) | (gid << 28) | id with the byte offset alongside; re-keyed as
// (comp << 56) | (gid << 36) | offset so that one radix sort yields stream order per (component, file).
__global__ void byread_rekey_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ ords, uint32_t n,
                                    uint64_t *__restrict__ keys2, uint32_t *__restrict__ ids)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i];
    const uint64_t gid = (k >> 28) & 0xfffffffull;
    keys2[i] = (k & 0xff00000000000000ull) | (gid << 36) | (ords[i] & 0xfffffffffull);
    ids[i] = (uint32_t)(k & 0xfffffffull);
}

// sorted occurrence i -> its record number (header starts at or before it inside its file) and the (comp, file) tally
__global__ void byread_assign_kernel(const uint64_t *__restrict__ keys2, uint32_t n, const uint64_t *__restrict__ goff,
                                     const uint64_t *__restrict__ hpos, const uint64_t *__restrict__ hdr_off, int n_genomes,
                                     uint64_t *__restrict__ read_of, uint32_t *__restrict__ per_cg)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys2[i];
    const uint32_t comp = (uint32_t)(k >> 56), gid = (uint32_t)((k >> 36) & 0xfffffu);
    const uint64_t abs = goff[gid] + (k & 0xfffffffffull);
    uint64_t lo = hdr_off[gid], hi = hdr_off[gid + 1];
    const uint64_t base = lo;
    while (lo < hi) {                                    // first header start > abs
        const uint64_t mid = (lo + hi) >> 1;
        if (hpos[mid] <= abs) lo = mid + 1; else hi = mid;
    }
    read_of[i] = lo - base;
    atomicAdd(&per_cg[(size_t)comp * n_genomes + gid], 1u);
}

// combco.index.<c> of one file as reads2mco writes it (:175-180): out[r] = occurrences of records 0..r, r = 0..n_reads
__global__ void byread_index_kernel(const uint64_t *__restrict__ read_of, uint64_t seg_n, uint64_t n_reads, uint64_t *__restrict__ out)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    uint64_t lo = 0, hi = seg_n;
    while (lo < hi) {                                    // first occurrence with record > r
        const uint64_t mid = (lo + hi) >> 1;
        if (read_of[mid] <= r) lo = mid + 1; else hi = mid;
    }
    out[r] = lo;
}

}  // namespace kssd
