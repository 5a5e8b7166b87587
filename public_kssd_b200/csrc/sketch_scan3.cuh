// sketch_scan3.cuh -- Stage I scan, third formulation: lazy validation + group prefilter.
//
// Replaces the byte loop of fasta2co / uniq_fasta2co / reads2mco (reference iseq2comem.c:205-270, :633-700, :78-186).
// The second formulation (sketch_scan32.cuh) was bound by the integer-ALU pipe at ~400 ALU instructions per KiB of text:
// exact per-byte classification (118) and one shared-memory probe per base (172) took most of them.  This one removes
// both from the steady state:
//
//  * LAZY VALIDATION.  In a clean iteration a byte is taken at face value from three of its bits: bit 3 set -> a skipped
//    byte (true for '\n' 0x0A and '\r' 0x0D, false for every letter of ACGTacgt), else a base whose code is bits 1-2
//    (A0 C1 T2 G3 -- "raw" codes).  The only exact test per byte is the cheap one that protects the stream STATE: a
//    byte with bit 3 set whose high nibble is not zero ('>', 'N', 'n', most IUPAC letters, digits 8-9 ...) makes the
//    iteration dirty and sends it to the exact general path, which is also where header lines live.  Everything else
//    that is not a base or a line end (R, S, W, B, D, V, digits 0-7, blanks, control bytes 0x08-0x0F ...) is carried
//    along as a fake base or a fake skip.  A genuine k-mer -- 2k bases with only line ends between them -- reads the
//    same either way, so no occurrence is lost; a k-mer that covers a fake is rejected at the very end, when the
//    ~1/2000 positions that survive both filters and the exact sampled-set lookup are verified byte by byte against
//    the text (verify_window).  Per 4-byte word: 3 LOP3 + 3 IMAD (codes, skip flags, dirty test), no shifts.
//
//  * GROUP PREFILTER.  Every window of 2s >= 12 bases (the central 2s-mer that decides sampling) contains exactly one
//    10-base block that starts at a multiple of three bases, so ONE probe of a 2^20-bit shared-memory bitmap, holding
//    the blocks at offsets 0, 1, 2 of every member of S u RC(S) (24 576 entries, 2.3 % full), stands for three windows:
//    12 probes per 32 bases instead of 32.  Lanes with a block hit park in a per-warp queue; 32 lanes at a time, each
//    block hit is settled by ONE 8-byte read of an L2-resident table indexed by the block (8 MiB): per offset the
//    sixteen possible values of the window's remaining bases, i.e. exact membership of the three windows (for 2s = 14 the
//    remaining 8 bits are folded to 4).  For 2s < 12 (subk <= 5) the same code runs with one probe per window and the
//    bitmap itself is exact.
//
//  * Bases are packed OLDEST-LOWEST with raw codes: the multiply that gathers a word's four codes needs no prior shift
//    in that order, and S u RC(S) is closed under the change of representation at table-build time.  The exact
//    resolver converts the survivors back (one group reversal, one XOR).
//
// Stream structure (spans pulled by warps from a ticket, ownership of k-mers by first base, run-out past the span end,
// the general path for dirty iterations) is unchanged from sketch_scan.cuh; the general path here rolls its k-mer in
// the same oldest-lowest representation so that it shares the filters and the resolver.
#pragma once
#include <type_traits>
#include "sketch_scan32.cuh"

namespace kssd {

constexpr uint32_t kPf3Words = 1u << 15;        // first level: 2^20 bits on a 10-base block (shared memory)
constexpr uint32_t kGtabEntries = 1u << 20;     // second level: per block, 3 offsets x 16 values of the rest of the window (global, L2)

// window (4s bits, scan representation) at block offset r: the bits outside the 10-base block, folded to 4
__host__ __device__ __forceinline__ uint32_t gtab_ext(uint32_t win, int r)
{
    const uint32_t x = (win & ((1u << (2 * r)) - 1u)) | ((win >> (2 * r + 20)) << (2 * r));
    return (x ^ (x >> 4)) & 15u;
}

// reverse the order of the low nb 2-bit groups of x (1 <= nb <= 32)
__host__ __device__ __forceinline__ uint64_t rev_groups64(uint64_t x, int nb)
{
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    x = ((x >> 8) & 0x00ff00ff00ff00ffull) | ((x & 0x00ff00ff00ff00ffull) << 8);
    x = ((x >> 16) & 0x0000ffff0000ffffull) | ((x & 0x0000ffff0000ffffull) << 16);
    x = (x >> 32) | (x << 32);
    return nb == 32 ? x : (x >> (64 - 2 * nb));
}

// reference representation (newest base lowest, A0 C1 G2 T3) <-> scan representation (oldest lowest, A0 C1 T2 G3)
__host__ __device__ __forceinline__ uint64_t to_scan_repr(uint64_t x, int nb)
{
    const uint64_t r = rev_groups64(x, nb);
    return r ^ ((r >> 1) & 0x5555555555555555ull);
}

// Parked lanes of the clean path: Y (history + own bases, 128 bits), the skip-flag mask of the lane's 32 bytes, the
// windows the stream position allows, the block hits, the lane's byte offset.
struct LaneQ3 {
    uint32_t y[4][kQueueCap];
    uint32_t flags[kQueueCap], wmask[kQueueCap], cand[kQueueCap], off[kQueueCap];
};
// candidates on their way to the exact resolver: the k-mer (scan representation), the lane's offset in its genome (own-base
// number in the top five bits), the lane's skip flags, the genome.  The queue outlives spans: entries carry all they need.
struct WarpQ3 { uint32_t lo[kQueueCap], hi[kQueueCap], ordlo[kQueueCap], ordhi[kQueueCap], f[kQueueCap], gid[kQueueCap]; };
#ifdef KSSD_SCAN_TMA
constexpr size_t kScan3SmemBytes = (size_t)kPf3Words * 4 + (size_t)kScanWarps * (sizeof(WarpQ3) + sizeof(LaneQ3) + 56 + 2072);   // + SpanCtx + TmaRing
#else
constexpr size_t kScan3SmemBytes = (size_t)kPf3Words * 4 + (size_t)kScanWarps * (sizeof(WarpQ3) + sizeof(LaneQ3) + 56);   // + SpanCtx
#endif

__device__ __forceinline__ bool pf3_probe(const uint32_t *__restrict__ pf, uint32_t v)
{
    return (__funnelshift_l(0u, pf[v & 0x7fffu], v >> 15) >> 31) != 0u;
}
// second level for one window whose block sits at offset 0 (the general path): exact for 2s = 12
__device__ __forceinline__ bool gtab_probe0(const unsigned long long *__restrict__ gtab, uint32_t win)
{
    return (__ldg(&gtab[win & 0xfffffu]) >> gtab_ext(win, 0)) & 1ull;
}

// The text itself decides: walking back from the window's last base, 2k letters of ACGTacgt with nothing but '\n' and
// '\r' between them, all inside the genome.  Runs for the few positions that passed every filter and the exact lookup.
__device__ __forceinline__ bool verify_window(const uint8_t *__restrict__ seq, uint64_t gs, uint64_t p, int TL)
{
    // The 32 bytes ending at p as nine aligned words, requested together (one latency) and classified four bytes at a
    // time with the exact table of the second formulation (classify4): bit i of `letter` / `bad` describes byte p - i.
    // The window is the first 2k letters walking back; it is genuine iff every byte up to its 2k-th letter is a letter
    // or a line end.  What 32 bytes cannot decide (line ends inside a window of k > 10, very short lines) takes the byte loop.
    int cnt = 0;
    if (p - gs >= 31) {
        const uint64_t first = p - 31, base = first & ~3ull;
        const uint32_t sh = (uint32_t)(first - base);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(seq + base);
        uint32_t w[9];
#pragma unroll
        for (int k = 0; k < 9; k++) w[k] = (k < 8 || sh) ? __ldg(wp + k) : 0u;       // the ninth word exists iff the range is unaligned
        uint64_t clean = 0, skip = 0;                              // bit b <-> byte base + b
#pragma unroll
        for (int k = 0; k < 9; k++) {
            uint32_t d = 0, t3, m;
            classify4(w[k], d, t3, m);
            const uint32_t z = ~((((d & 0x7f7f7f7fu) + 0x7f7f7f7fu) | d)) & 0x80808080u;      // 0x80 per CLEAN byte
            clean |= (uint64_t)((((z >> 7) * 0x00204081u) >> 21) & 15u) << (4 * k);
            skip |= (uint64_t)(((((t3 >> 2) & 0x01010101u) * 0x00204081u) >> 21) & 15u) << (4 * k);
        }
        const uint32_t rc = __brev((uint32_t)(clean >> sh)), rs = __brev((uint32_t)(skip >> sh));     // bit i <-> byte p - i
        const uint32_t letter = rc & ~rs, bad = ~rc;
        const uint32_t t = __fns(letter, 0, TL);                  // position of the 2k-th letter, or none
        if (t != 0xffffffffu) return (bad & (0xffffffffu >> (31 - t))) == 0;
        if (bad) return false;                                    // something else before 2k letters were seen
        if (p - gs < 32) return false;                            // the genome starts before them
        cnt = __popc(letter);
        p -= 32;
    }
    for (;;) {
        const uint32_t b = __ldg(seq + p);
        const uint32_t l = b | 0x20u;
        if (l == 'a' || l == 'c' || l == 'g' || l == 't') {
            if (++cnt == TL) return true;
        } else if (b != '\n' && b != '\r') return false;
        if (p == gs) return false;
        p--;
    }
}

// FASTQ reads (sketch_fastq3.cuh): the 2k bytes ending at absolute offset p are all of ACGTacgt, none before the file's first byte
__device__ __forceinline__ bool verify_window_strict(const uint8_t *__restrict__ seq, uint64_t gs, uint64_t p, int TL)
{
    if (p - gs + 1 < (uint64_t)TL) return false;
    for (int i = 0; i < TL; i++) {
        const uint32_t l = __ldg(seq + p - i) | 0x20u;
        if (!(l == 'a' || l == 'c' || l == 'g' || l == 't')) return false;
    }
    return true;
}

// exact resolution of queued candidates (scan representation), up to 32 at a time.  Out of line: three call sites,
// a few thousand calls per launch -- the clean loop should not carry this code in its instruction-cache footprint.
__device__ __noinline__ void resolve3(const SketchParams &P, const ScanArgs &A, const WarpQ3 &q, uint32_t first, uint32_t m)
{
    const uint32_t lane = lane_id();
    bool found = false;
    uint64_t key = 0, ordv = 0;
    uint32_t gid = 0;
    if (lane < m) {
        const uint64_t y = ((uint64_t)q.hi[first + lane] << 32) | q.lo[first + lane];
        // byte of the occurrence's last base: the lane's offset plus the (j+1)-th byte of the lane without a skip flag
        const uint32_t oh = q.ordhi[first + lane];
        uint64_t lane_ord = ((uint64_t)(oh & 0x07ffffffu) << 32) | q.ordlo[first + lane];       // 59 bits, two's complement: a lane may start
        if (lane_ord >> 58) lane_ord |= ~((1ull << 59) - 1ull);                                    // a few bytes before its genome does
        ordv = lane_ord + __fns(~q.f[first + lane], 0, (int)(oh >> 27) + 1);
        gid = q.gid[first + lane];
        const uint64_t gs = A.goff[gid];
        const uint64_t yf = y ^ ((y >> 1) & 0x5555555555555555ull);       // raw -> A0 C1 G2 T3
        const uint64_t fwd = rev_groups64(yf, P.TL);                      // the reference's tuple (newest base lowest)
        const uint64_t rc = ~yf & P.tupmask;                              // its crvstuple: complement, oldest base lowest
        const uint64_t u = fwd < rc ? fwd : rc;
        const uint32_t inner = (uint32_t)(u >> (2 * P.out)) & P.innermask;
        uint32_t h = mix32(inner) & P.ht_mask;
        uint32_t pf = 0;
        for (;;) {
            const uint2 e = __ldg(&P.ht[h]);
            if (e.x == inner) { found = true; pf = e.y; break; }
            if (e.x == kHtEmpty) break;
            h = (h + 1) & P.ht_mask;
        }
        if (found) {
            const uint64_t dr = (((u & P.undomask) + ((u & P.outmask) << (4 * P.s))) >> (4 * P.L)) + pf;
            key = ((dr & P.comp_mask) << 56) | ((uint64_t)gid << 28) | (dr >> P.comp_code_bits);
            found = A.strict_window ? verify_window_strict(A.seq, gs, gs + ordv, P.TL) : verify_window(A.seq, gs, gs + ordv, P.TL);
            if (found && A.drop_zero && dr == 0) { found = false; atomicAdd(&A.zero_count[gid], 1u); }
        }
    }
    if (A.bkeys) {                                        // bucket mode: straight to the (component, genome) bucket
        if (found) {
            const uint32_t b = (uint32_t)(key >> 56) * A.n_genomes + gid, lo = A.boff[b], cap = A.boff[b + 1] - lo;
            const uint32_t n = atomicAdd(&A.bcnt[b], 1u);
            if (n < cap) A.bkeys[lo + n] = ((key & 0x0fffffffull) << 36) | (ordv & 0xfffffffffull);
            else *A.boverflow = 1u;
        }
        return;
    }
    const uint32_t fm = __ballot_sync(kFull, found);
    if (fm) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(A.out_count, (uint32_t)__popc(fm));
        base = __shfl_sync(kFull, base, 0);
        if (found) {
            const uint32_t idx = base + __popc(fm & ((1u << lane) - 1u));
            if (idx < A.out_cap) { A.out_keys[idx] = key; A.out_ords[idx] = ordv; }
        }
    }
}

// append the lanes flagged `has` to the warp's candidate queue; resolve when 32 are waiting.
// lane_ord = offset of the lane's first byte in its genome (general path: of the byte itself, j = 0, f = 0).
__device__ __forceinline__ void queue_push3(const SketchParams &P, const ScanArgs &A, WarpQ3 &q, uint32_t &qn, bool has, uint64_t kmer,
                                            uint64_t lane_ord, uint32_t j, uint32_t f, uint32_t gid)
{
    const uint32_t pm = __ballot_sync(kFull, has);
    if (!pm) return;
    if (has) {
        const uint32_t slot = qn + __popc(pm & ((1u << lane_id()) - 1u));
        q.lo[slot] = (uint32_t)kmer;
        q.hi[slot] = (uint32_t)(kmer >> 32);
        q.ordlo[slot] = (uint32_t)lane_ord;
        q.ordhi[slot] = ((uint32_t)(lane_ord >> 32) & 0x07ffffffu) | (j << 27);
        q.f[slot] = f;
        q.gid[slot] = gid;
    }
    qn += __popc(pm);
    __syncwarp();
    if (qn >= 32) {
        resolve3(P, A, q, qn - 32, 32);
        qn -= 32;
        __syncwarp();
    }
}

// Parked lanes first .. first+m-1, one per lane; ONE block hit of every lane is settled per call.  ST = 3: block hit i
// (windows 3i-2 .. 3i) by one read of the block's table entry; ST = 1: the bitmap was exact, the hit is a member.  Members go to
// the candidate queue.  Lanes that still hold block hits (one in seven) are parked again, compacted at `first`, and meet the next
// drain with every lane busy -- looping here until the last lane is done ran 2.4 rounds for 1.15 hits per lane.  Returns the new
// number of parked lanes.
#ifdef KSSD_DRAIN_NOINLINE
#define KSSD_DRAIN_ATTR __noinline__
#else
#define KSSD_DRAIN_ATTR __forceinline__
#endif
template <int ST>
__device__ KSSD_DRAIN_ATTR uint32_t drain3(const SketchParams &P, const ScanArgs &A, WarpQ3 &q, uint32_t &qn, LaneQ3 &lq, uint32_t first,
                                           uint32_t m, uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    const uint32_t e = first + (lane < m ? lane : 0u);
    uint32_t cand = 0, wm = 0, F = 0, off = 0, y0 = 0, y1 = 0, y2 = 0, y3 = 0;
    if (lane < m) {
        cand = lq.cand[e]; wm = lq.wmask[e]; F = lq.flags[e]; off = lq.off[e];
        y0 = lq.y[0][e]; y1 = lq.y[1][e]; y2 = lq.y[2][e]; y3 = lq.y[3][e];
    }
    if (ST == 1) cand &= wm;
    auto yw = [&](uint32_t a) -> uint32_t { return a == 0 ? y0 : (a == 1 ? y1 : (a == 2 ? y2 : (a == 3 ? y3 : 0u))); };
    // the lane's block hit i: S = the 16 bases from X position 3i-2 on (the three windows of the hit start at its bases 2, 1, 0);
    // its table entry is requested first, the re-parking below runs under the read's latency
    int i = 0;
    uint32_t S = 0;
    unsigned long long mm = 0;
    const bool act = cand != 0;
    if (act) {
        i = __ffs(cand) - 1;
        cand &= cand - 1;
        if (ST == 3) {
            const int o = 2 * (P.out + 3 * i) - 4;
            if (o < 0) S = y0 << (-o);
            else { const uint32_t a = (uint32_t)o >> 5; S = __funnelshift_r(yw(a), yw(a + 1), (uint32_t)o); }
            mm = __ldg(&P.gtab[(S >> 4) & 0xfffffu]);
        }
    }
    // lanes with block hits left: parked again, compacted at `first` (every lane holds its entry in registers by now)
    const uint32_t rem = __ballot_sync(kFull, cand != 0);
    if (rem) {
        __syncwarp();                                     // (every lane's reads of its entry are done before a slot is rewritten)
        if (cand) {
            const uint32_t d = first + __popc(rem & ((1u << lane) - 1u));
            lq.y[0][d] = y0; lq.y[1][d] = y1; lq.y[2][d] = y2; lq.y[3][d] = y3;
            lq.flags[d] = F; lq.wmask[d] = wm; lq.cand[d] = cand; lq.off[d] = off;
        }
        __syncwarp();
    }
    uint32_t hits = 0;                                    // ST = 3: bit r <-> window 3i - r is a member
    if (act) {
        if (ST == 1) hits = 1u;
        else {
            const uint32_t m01 = (uint32_t)mm, m2 = (uint32_t)(mm >> 32);
            uint32_t e0, e1, e2;                          // what each window holds outside the block, folded to 4 bits
            if (P.s == 6) { e0 = (S >> 24) & 15u; e1 = ((S >> 2) & 3u) | ((S >> 22) & 12u); e2 = S & 15u; }
            else { e0 = gtab_ext((S >> 4) & P.innermask, 0); e1 = gtab_ext((S >> 2) & P.innermask, 1); e2 = gtab_ext(S & P.innermask, 2); }
            const int j0 = 3 * i;                         // windows j0, j0 - 1, j0 - 2 (own bases 0 .. 31 only)
            const uint32_t okm = j0 < 32 ? (wm >> j0) & 1u : 0u;
            const uint32_t ok1 = (j0 >= 1 && j0 < 33) ? (wm >> (j0 - 1)) & 1u : 0u;
            const uint32_t ok2 = j0 >= 2 ? (wm >> (j0 - 2)) & 1u : 0u;
            hits = (okm & (m01 >> e0)) | ((ok1 & (m01 >> (16 + e1))) << 1) | ((ok2 & (m2 >> e2)) << 2);
        }
    }
    while (__any_sync(kFull, hits != 0)) {
        const bool has = hits != 0;
        uint32_t j = 0;
        uint64_t kmer = 0;
        if (has) {
            const int r = __ffs(hits) - 1;
            hits &= hits - 1;
            j = (uint32_t)(ST == 1 ? i : 3 * i - r);
            const uint32_t a = (2u * j) >> 5;             // 0 or 1: the k-mer is Y bits [2j, 2j + 4k)
            const uint32_t w0 = a ? y1 : y0, w1 = a ? y2 : y1, w2 = a ? y3 : y2;
            kmer = ((((uint64_t)__funnelshift_r(w1, w2, 2u * j)) << 32) | __funnelshift_r(w0, w1, 2u * j)) & P.tupmask;
        }
        queue_push3(P, A, q, qn, has, kmer, ord_base + off, j, F, gid);
    }
    return first + __popc(rem);
}

// One 512-byte GENERAL iteration (16 bytes per lane): headers, N, IUPAC, anything -- exact per byte.  `cur` is already
// masked to the span / genome extent.  Same state machine as general_iter16 (sketch_scan.cuh); the k-mer rolls in the
// scan representation and hits go straight to the candidate queue.
__device__ __noinline__ void general_iter3(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQ3 &q, uint32_t &qn,
                                           StreamState &st, uint4 cur, uint64_t cbase, uint64_t end, bool past_end, uint32_t lane_off,
                                           uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    const int TL = P.TL;
    const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
    const uint64_t laddr = cbase + 16 * lane;
    uint32_t V = 0, NLm = 0, CRm = 0, GTm = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t b = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;
        const uint32_t l = b | 0x20u;
        V |= (uint32_t)(l == 'a' || l == 'c' || l == 'g' || l == 't') << i;
        NLm |= (uint32_t)(b == '\n') << i;
        CRm |= (uint32_t)(b == '\r') << i;
        GTm |= (uint32_t)(b == '>') << i;
    }
    // header state: '>' sets, '\n' clears; carry-propagate through the lane, then across lanes
    const uint32_t ev = GTm | NLm;
    const bool has_ev = ev != 0;
    const bool last_set = has_ev && ((GTm >> (31 - __clz(ev))) & 1u);
    const uint32_t evS = __ballot_sync(kFull, last_set);
    const uint32_t evA = __ballot_sync(kFull, has_ev);
    const uint32_t prev = evA & ((1u << lane) - 1u);
    const uint32_t h_in = prev ? ((evS >> (31 - __clz(prev))) & 1u) : st.hdr;
    const uint32_t Aa = ~NLm & 0xffffu, Bb = GTm;
    const uint32_t sum = Aa + Bb + h_in;
    const uint32_t hdrmask = (sum ^ Aa ^ Bb) & 0xffffu;      // bit i: byte i lies inside a header
    st.hdr = __shfl_sync(kFull, (sum >> 16) & 1u, 31);
    const uint32_t Veff = V & ~hdrmask;
    const uint32_t BRK = ~(V | NLm | CRm) & ~hdrmask & 0xffffu;
    // lane summary (newest base lowest, raw codes): bases after the lane's last break
    uint32_t tb = 0, tn = 0;
    bool hb = false;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t c = (w[i >> 2] >> (8 * (i & 3) + 1)) & 3u;
        if ((Veff >> i) & 1u) { tb = (tb << 2) | c; tn++; }
        else if ((BRK >> i) & 1u) { tb = 0; tn = 0; hb = true; }
    }
    // inclusive scan of (bits, n, broke) under "append unless the right part broke"
    uint64_t sb = tb;
    uint32_t sn = tn;
    uint32_t sbrk = hb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t ob = shfl_up64(sb, o);
        const uint32_t on = __shfl_up_sync(kFull, sn, o);
        const uint32_t obrk = __shfl_up_sync(kFull, sbrk, o);
        if (lane >= (uint32_t)o && !sbrk) {
            if (sn < 32) sb |= ob << (2 * sn);
            sn = min(sn + on, 32u);
            sbrk = obrk;
        }
    }
    uint64_t eb = shfl_up64(sb, 1);
    uint32_t en = __shfl_up_sync(kFull, sn, 1);
    uint32_t ebrk = __shfl_up_sync(kFull, sbrk, 1);
    if (lane == 0) { eb = 0; en = 0; ebrk = 0; }
    uint64_t hist;
    uint32_t run;
    if (ebrk) { hist = eb; run = en; }
    else { hist = (en < 32 ? (st.cw << (2 * en)) : 0ull) | eb; run = min(st.since_break + en, kRunCap); }
    // valid bases at offsets >= end (run-out accounting)
    uint32_t gem = 0;
    if (past_end) {
        const int64_t rel = (int64_t)end - (int64_t)laddr;
        gem = rel <= 0 ? 0xffffu : (rel >= 16 ? 0u : (~((1u << rel) - 1u) & 0xffffu));
    }
    const uint32_t cge = __popc(Veff & gem);
    uint32_t ginc = cge;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, ginc, o);
        if (lane >= (uint32_t)o) ginc += t;
    }
    uint32_t ae = st.after_end + ginc - cge;
    // walk the lane's bytes; the last 2k bases roll in the scan representation (age a at group 2k-1-a).  A real loop
    // (this path is rare): one copy of the queue hand-over, bytes picked out of two 64-bit halves.
    uint64_t fr = rev_groups64(hist & P.tupmask, TL);
    const uint64_t wlo = ((uint64_t)cur.y << 32) | cur.x, whi = ((uint64_t)cur.w << 32) | cur.z;
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
        bool hit = false;
        if ((Veff >> i) & 1u) {
            const uint64_t c = ((i < 8 ? wlo : whi) >> (8 * (i & 7) + 1)) & 3ull;
            fr = (fr >> 2) | (c << (2 * TL - 2));
            run = min(run + 1, kRunCap);
            ae += (gem >> i) & 1u;
            if (run >= (uint32_t)TL && ae <= (uint32_t)(TL - 1)) {
                const uint32_t v = (uint32_t)(fr >> (2 * P.out));
                if (pf3_probe(pf, v)) hit = P.gtab ? gtab_probe0(P.gtab, v & P.innermask) : true;
            }
        } else if ((BRK >> i) & 1u) run = 0;
        queue_push3(P, A, q, qn, hit, fr, ord_base + lane_off + i, 0u, 0u, gid);
    }
    // warp carry = inclusive value of lane 31 on top of the old carry
    const uint64_t sb31 = shfl64(sb, 31);
    const uint32_t sn31 = __shfl_sync(kFull, sn, 31), sbrk31 = __shfl_sync(kFull, sbrk, 31);
    if (sbrk31) { st.cw = sb31; st.since_break = sn31; }
    else { st.cw = (sn31 < 32 ? (st.cw << (2 * sn31)) : 0ull) | sb31; st.since_break = min(st.since_break + sn31, kRunCap); }
    st.after_end += __shfl_sync(kFull, ginc, 31);
}

// ---- lazy classification of two words (8 bytes): codes of each word in the top byte of c0 / c1 (oldest base lowest),
// the eight skip flags (bit 3 of each byte) in the top byte of g (oldest lowest), dacc |= nonzero iff a byte with
// bit 3 set has a non-zero high nibble.  The multipliers only ever add terms at distinct bit positions: no carries.
__device__ __forceinline__ void classify_lazy8(uint32_t w0, uint32_t w1, uint32_t &dacc, uint32_t &c0, uint32_t &c1, uint32_t &g)
{
    const uint32_t f0 = w0 & 0x08080808u, f1 = w1 & 0x08080808u;
    c0 = (w0 & 0x06060606u) * 0x00820820u;       // bits 1-2 of bytes 0..3 -> bits 24..31
    c1 = (w1 & 0x06060606u) * 0x00820820u;
    dacc |= w0 & (f0 * 0x1eu);                   // bit 3 spread over bits 4-7 of the same byte
    dacc |= w1 & (f1 * 0x1eu);
    g = f0 * 0x00204081u + f1 * 0x02040810u;     // flags of w0 -> bits 24..27, of w1 -> bits 28..31
}
// top bytes of four words -> one word, a's byte lowest
__device__ __forceinline__ uint32_t top_bytes4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return prmt(prmt(a, b, 0x0073u), prmt(c, d, 0x0073u), 0x5410u);
}

#ifdef KSSD_SCAN_TMA
// A/B variant (profiles/r2_sketch_ncu_summary.md): the steady iterations' text arrives by 1-D bulk copies (TMA,
// cp.async.bulk + mbarrier complete_tx) into a two-stage, 1 KiB-per-stage ring per warp, two iterations ahead, and the
// lanes read their 32 bytes back with two LDS.128 -- instead of one LDG.E.256 per lane into registers.
struct TmaRing { alignas(16) uint8_t buf[2][1024]; alignas(8) unsigned long long bar[2]; };
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_issue(TmaRing &r, int stage, const uint8_t *src)
{
    const uint32_t bar = smem_u32(&r.bar[stage]), dst = smem_u32(r.buf[stage]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 1024, [%2];" ::"r"(dst), "l"(src), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_wait(TmaRing &r, int stage, uint32_t parity)
{
    const uint32_t bar = smem_u32(&r.bar[stage]);
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
#endif

// Cold per-span state lives in shared memory (one record per warp): the span's byte extents are 64-bit and only the
// first and last iterations of a span look at them -- in registers they pushed the loop counters out to local memory.
struct SpanCtx { uint64_t gs, ge, start, end, chunk0; uint32_t gid, after_end, ln, qn; };

// ST: bases per first-level probe (3 or 1); BIG: 2k-1 history bases need more than 32 bits (k >= 9)
template <int ST, bool BIG>
__device__ void scan_span3(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQ3 &q, LaneQ3 &lq,
                           volatile SpanCtx &sc
#ifdef KSSD_SCAN_TMA
                           , TmaRing &ring
#endif
                           )
{
    constexpr int NPROBE = ST == 3 ? 12 : 32;
    const uint32_t lane = lane_id();
    const int TL = P.TL;
    const uint32_t hsh = (2u * (uint32_t)(TL - 1)) & 31u;      // own bases sit above the 2k-1 history bases: bit 32*BIG + hsh
    // stream state in registers; packed into a StreamState only around the out-of-line general iterations
    uint32_t cw0 = 0, cw1 = 0;                                   // the last 2k-1 bases of the stream, oldest lowest
    uint32_t since_break = 0, hdr = 0;
    uint32_t ln = 0;                                             // parked lanes of this warp
    uint32_t n_steady;
    Bytes32 cur;
    {
        const uint64_t start = sc.start, end = sc.end, ge = sc.ge, chunk0 = start & ~127ull;
        if (lane == 0) { sc.chunk0 = chunk0; sc.after_end = 0; }
        __syncwarp();
        // iterations 1 .. n_steady are "steady": wholly inside [start, min(end, ge)) -- no masking, no run-out logic
        const uint64_t lim = end < ge ? end : ge;
        const uint64_t full = (lim - chunk0) >> 10;
        n_steady = full > 1 ? (uint32_t)(full - 1 < 0x3fffffffull ? full - 1 : 0x3fffffffull) : 0u;
        cur = load_chunk32_guarded(A, chunk0 + 32 * lane);
#ifdef KSSD_SCAN_TMA
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring.bar[0])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&ring.bar[1])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (n_steady >= 1) tma_issue(ring, 1, A.seq + chunk0 + 1024);          // iteration j lives in stage j & 1
            if (n_steady >= 2) tma_issue(ring, 0, A.seq + chunk0 + 2048);
        }
        __syncwarp();
#endif
    }
    bool at_eof = false;

    // One CLEAN iteration from the three words of the classification on: squeeze, history, Y / X, the block probes, parking.
    // FAST: a steady iteration with a saturated run and no header pending -- the bulk of every span -- where nothing of the
    // span's edges applies (no masked halves for k >= 9, no cut lanes, no position filter, no run accounting).
    auto clean_iter = [&](auto fast_tag, uint32_t it, uint32_t lane_off, uint32_t PA, uint32_t PB, uint32_t F, uint32_t n, uint32_t nA, bool steady,
                          bool past_end) {
        constexpr bool FAST = decltype(fast_tag)::value;
        {
            uint32_t fa = F & 0xffffu, fb = F >> 16;
            if (!FAST || !BIG) {                             // (a steady lane with >= 2k-1 >= 17 bases has no half without one)
                if (fa == 0xffffu) { fa = 0; PA = 0; }       // a half masked out whole (span start, genome end)
                if (fb == 0xffffu) { fb = 0; PB = 0; }
            }
            for (;;) {      // squeeze the skipped bytes out of both halves; first round is branch-free
                const uint32_t ia = fa & (0u - fa), ib = fb & (0u - fb);
                const uint32_t la = ia * ia - 1u, lb = ib * ib - 1u;
                PA = ((PA >> 2) & ~la) | (PA & la);
                PB = ((PB >> 2) & ~lb) | (PB & lb);
                fa = (fa ^ ia) >> 1;
                fb = (fb ^ ib) >> 1;
                if (!__any_sync(kFull, (fa | fb) != 0)) break;
            }
        }
        // the lane's n bases, oldest lowest: PA | PB << 2 nA
        const uint32_t Q0 = PA | __funnelshift_lc(0u, PB, 2 * nA), Q1 = __funnelshift_lc(PB, 0u, 2 * nA);
        // what the next lane needs of them: the last 2k-1, again oldest lowest
        uint32_t S0, S1;
        {
            const int d = 2 * ((int)n - (TL - 1));
            if (FAST || steady || d >= 0) {
                if (BIG) { S0 = __funnelshift_rc(Q0, Q1, d); S1 = __funnelshift_rc(Q1, 0u, d); }      // d <= 32
                else { S0 = (uint32_t)((((uint64_t)Q1 << 32) | Q0) >> d); S1 = 0u; }                  // d <= 62, 2k-1 <= 15 bases left
            } else {                                                                                  // a cut lane with fewer bases: they end at group 2k-2
                const uint64_t s = (((uint64_t)Q1 << 32) | Q0) << (-d);
                S0 = (uint32_t)s; S1 = (uint32_t)(s >> 32);
            }
        }
        uint32_t H0 = __shfl_up_sync(kFull, S0, 1), H1 = BIG ? __shfl_up_sync(kFull, S1, 1) : 0u;
        if (lane == 0) { H0 = cw0; H1 = cw1; }
        // Y = history | own bases << 2(2k-1): the k-mer that ends at own base j is Y bits [2j, 2j + 4k)
        uint32_t Y0, Y1, Y2, Y3;
        if (BIG) {
            Y0 = H0;
            Y1 = H1 | (Q0 << hsh);
            Y2 = __funnelshift_l(Q0, Q1, hsh);
            Y3 = __funnelshift_lc(Q1, 0u, hsh);
        } else {
            Y0 = H0 | (Q0 << hsh);
            Y1 = __funnelshift_l(Q0, Q1, hsh);
            Y2 = __funnelshift_lc(Q1, 0u, hsh);
            Y3 = 0u;
        }
        // X = Y >> 2 out: the central 2s-mer of that k-mer is X bits [2j, 2j + 4s)
        const uint32_t X0 = __funnelshift_r(Y0, Y1, 2 * P.out), X1 = __funnelshift_r(Y1, Y2, 2 * P.out), X2 = __funnelshift_r(Y2, Y3, 2 * P.out);
        auto xsh = [&](int k) -> uint32_t {      // X >> k for a constant k in [-2, 95]
            return k < 0 ? (X0 << (-k)) : (k < 32 ? __funnelshift_r(X0, X1, k) : (k < 64 ? __funnelshift_r(X1, X2, k - 32) : (X2 >> (k - 64))));
        };
        uint32_t cand = 0;
#pragma unroll
        for (int i = NPROBE - 1; i >= 0; i--) {
            const uint32_t word = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(pf) + (xsh(2 * ST * i - 2) & 0x1fffcu));
            cand = __funnelshift_l(__funnelshift_l(0u, word, xsh(2 * ST * i + 15)), cand, 1);      // cand = cand << 1 | flag
        }

        uint32_t wm = low_mask((int)n);                          // windows (own bases) the stream position allows
        if (!FAST && (since_break < kRunCap || past_end)) {
            const uint32_t N = __reduce_add_sync(kFull, n);
            if (since_break < (uint32_t)(TL - 1) || past_end) {
                // start of a span / run-out past its end: filter by position inside the iteration
                uint32_t incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                const int o_l = (int)(incl - n);                     // bases before this lane
                const int need = TL - 1 - (int)since_break - o_l;    // own base j ends a k-mer of this span iff j >= need
                if (need > 0) wm &= ~low_mask(min(need, 32));
                if (past_end) {
                    uint32_t E;                                      // bases of this iteration before `end`
                    const uint32_t after_end = sc.after_end;
                    const int64_t rel = (int64_t)sc.end - (int64_t)(sc.chunk0 + ((uint64_t)it << 10));
                    if (rel <= 0) E = 0;
                    else {
                        const int le = (int)(rel >> 5), be = (int)(rel & 31);      // bytes [0, be) of lane `le` lie before `end`
                        E = __shfl_sync(kFull, (uint32_t)o_l + (uint32_t)be - (uint32_t)__popc(F & low_mask(be)), le);
                    }
                    // the k-mer's first base lies before `end` iff after_end + (o_l + j - E + 1) <= 2k-1
                    const int keep = TL - 1 - (int)after_end + (int)E - o_l;      // j < keep
                    if (keep < 32) wm &= low_mask(max(keep, 0));
                    __syncwarp();
                    if (lane == 0) sc.after_end = after_end + N - E;
                    __syncwarp();
                }
            }
            since_break = min(since_break + N, kRunCap);
        }
        cw0 = __shfl_sync(kFull, S0, 31);
        if (BIG) cw1 = __shfl_sync(kFull, S1, 31);
        const uint32_t hit = __ballot_sync(kFull, cand != 0);
        if (hit) {
            if (cand) {
#ifdef KSSD_SCAN_GTAB_PF
                if (ST == 3 && FAST) {      // the table entry of the lane's first block hit: asked into L1 now, read by the drain an iteration or three later
                    const uint32_t sh = 6u * (uint32_t)(__ffs(cand) - 1), a = sh >> 5;
                    const uint32_t lo = a == 0 ? X0 : (a == 1 ? X1 : X2), hi = a == 0 ? X1 : (a == 1 ? X2 : 0u);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(P.gtab + (__funnelshift_r(lo, hi, sh) & 0xfffffu)));
                }
#endif
                const uint32_t i = ln + __popc(hit & ((1u << lane) - 1u));
                lq.y[0][i] = Y0; lq.y[1][i] = Y1; lq.y[2][i] = Y2; lq.y[3][i] = Y3;
                lq.flags[i] = F; lq.wmask[i] = wm; lq.cand[i] = cand; lq.off[i] = lane_off;
            }
            ln += __popc(hit);
            __syncwarp();
            if (ln >= 32) {
                uint32_t qn = sc.qn;                              // candidates waiting in the warp's queue (cold state too)
                do ln = drain3<ST>(P, A, q, qn, lq, ln - 32, 32, sc.gid, sc.chunk0 - sc.gs); while (ln >= 32);
                if (lane == 0) sc.qn = qn;
                __syncwarp();
            }
        }
    };

    for (uint32_t it = 0;; it++) {
#ifndef KSSD_SCAN_TMA
        // ---- fast loop: steady iterations 1 .. n_steady-1 with a saturated run and no header pending.  A dirty iteration leaves
        // the loop with its text still in `cur` and goes through the general iteration below as it is.
        if (it - 1u < n_steady - 1u && n_steady && since_break >= kRunCap && !hdr) {
            const uint8_t *np = A.seq + sc.chunk0 + ((uint64_t)(it + 1) << 10) + 32 * lane;      // the lane's text of the next iteration
            for (;;) {
                uint32_t dacc = 0, PA, PB, F;
                {
                    uint32_t c0, c1, c2, c3, c4, c5, c6, c7, g0, g1, g2, g3;
                    classify_lazy8(cur.lo.x, cur.lo.y, dacc, c0, c1, g0);
                    classify_lazy8(cur.lo.z, cur.lo.w, dacc, c2, c3, g1);
                    classify_lazy8(cur.hi.x, cur.hi.y, dacc, c4, c5, g2);
                    classify_lazy8(cur.hi.z, cur.hi.w, dacc, c6, c7, g3);
                    F = top_bytes4(g0, g1, g2, g3);
                    PA = top_bytes4(c0, c1, c2, c3);
                    PB = top_bytes4(c4, c5, c6, c7);
                }
                const uint32_t nA = 16 - __popc(F & 0xffffu), n = 32 - __popc(F);
                if (!__all_sync(kFull, dacc == 0 && n >= (uint32_t)(TL - 1))) break;
                cur = ldg_stream256(np);
                np += 1024;
#ifndef KSSD_SCAN_NO_L2PF
                // an iteration takes ~1 us, about what a DRAM access takes under load: the text of the iteration after next is
                // asked into L2 now (no register, no shared memory), so that the load above finds it there
#ifndef KSSD_SCAN_PFDIST
#define KSSD_SCAN_PFDIST 1
#endif
                if (it + 2 + KSSD_SCAN_PFDIST <= n_steady) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + 1024 * KSSD_SCAN_PFDIST));
#endif
                clean_iter(std::true_type{}, it, (it << 10) + 32 * lane, PA, PB, F, n, nA, true, false);
                if (++it >= n_steady) break;                      // the last steady iteration loads its successor guarded: below
            }
        }
#endif
        const bool steady = (it - 1u) < n_steady;
        const uint32_t lane_off = (it << 10) + 32 * lane;

        bool past_end = false, cut_lane = false;
        if (!steady) {
            const uint64_t start = sc.start, ge = sc.ge, end = sc.end, cbase = sc.chunk0 + ((uint64_t)it << 10);
            const uint64_t laddr = cbase + 32 * lane;
            if (cbase < start || cbase + 1024 > ge) {
                mask_lane_bytes(cur.lo, clamp16((int64_t)start - (int64_t)laddr), clamp16((int64_t)ge - (int64_t)laddr));
                mask_lane_bytes(cur.hi, clamp16((int64_t)start - (int64_t)(laddr + 16)), clamp16((int64_t)ge - (int64_t)(laddr + 16)));
            }
            past_end = cbase + 1024 > end;
            cut_lane = laddr < start || laddr + 32 > ge;
        }

        uint32_t dacc = 0, PA, PB, F;
        {
            uint32_t c0, c1, c2, c3, c4, c5, c6, c7, g0, g1, g2, g3;
            classify_lazy8(cur.lo.x, cur.lo.y, dacc, c0, c1, g0);
            classify_lazy8(cur.lo.z, cur.lo.w, dacc, c2, c3, g1);
            classify_lazy8(cur.hi.x, cur.hi.y, dacc, c4, c5, g2);
            classify_lazy8(cur.hi.z, cur.hi.w, dacc, c6, c7, g3);
            F = top_bytes4(g0, g1, g2, g3);                       // bit b: byte b of the lane is skipped
            PA = top_bytes4(c0, c1, c2, c3);
            PB = top_bytes4(c4, c5, c6, c7);
        }
        // the 32 bytes are now three words: request the next KiB into the same registers (one buffer, no copies); the
        // rest of the iteration and the other warps cover its latency.  A dirty iteration re-reads its text itself.
        if (it < n_steady) {       // (the address is rebuilt from the span record: a live 64-bit pointer costs two registers all loop long)
#ifdef KSSD_SCAN_TMA
            const uint32_t j = it + 1;                                // the iteration whose text is due now
            tma_wait(ring, j & 1, ((j - 1) >> 1) & 1);
            const uint4 *sp = reinterpret_cast<const uint4 *>(ring.buf[j & 1] + 32 * lane);
            cur.lo = sp[0];
            cur.hi = sp[1];
            __syncwarp();
            if (lane == 0 && j + 2 <= n_steady) {                     // its stage is free again: two iterations ahead
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_issue(ring, j & 1, A.seq + sc.chunk0 + ((uint64_t)(j + 2) << 10));
            }
#else
            cur = ldg_stream256(A.seq + sc.chunk0 + ((uint64_t)(it + 1) << 10) + 32 * lane);
#endif
        } else {
            const uint64_t nb = sc.chunk0 + ((uint64_t)(it + 1) << 10);
            if (nb < sc.ge) cur = load_chunk32_guarded(A, nb + 32 * lane);
        }
        const uint32_t nA = 16 - __popc(F & 0xffffu), n = 32 - __popc(F);
        const bool lane_ok = dacc == 0 && (n >= (uint32_t)(TL - 1) || cut_lane);
        const bool clean = __all_sync(kFull, lane_ok) && !hdr;

        if (clean) {
            clean_iter(std::false_type{}, it, lane_off, PA, PB, F, n, nA, steady, past_end);
        } else {
            // two general 512-byte iterations with a 16-byte lane mapping (reloaded: L2 hits); the carry changes
            // representation on the way in and out
            const uint64_t cw = ((uint64_t)cw1 << 32) | cw0;
            const uint64_t start = sc.start, ge = sc.ge, end = sc.end, cbase = sc.chunk0 + ((uint64_t)it << 10);
            StreamState st = {rev_groups64(cw, TL - 1), since_break, sc.after_end, hdr};
            uint32_t qn = sc.qn;
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint64_t sbase = cbase + 512ull * h;
                if (sbase >= ge) break;
                const uint64_t laddr = sbase + 16 * lane;
                uint4 c16 = load_chunk16_guarded(A, laddr);
                if (sbase < start || sbase + 512 > ge)
                    mask_lane_bytes(c16, clamp16((int64_t)start - (int64_t)laddr), clamp16((int64_t)ge - (int64_t)laddr));
                general_iter3(P, A, pf, q, qn, st, c16, sbase, end, sbase + 512 > end, (it << 10) + 512u * h + 16 * lane, sc.gid, sc.chunk0 - sc.gs);
            }
            const uint64_t cwr = rev_groups64(st.cw & (P.tupmask >> 2), TL - 1);
            cw0 = (uint32_t)cwr; cw1 = (uint32_t)(cwr >> 32);
            since_break = st.since_break; hdr = st.hdr;
            __syncwarp();
            if (lane == 0) { sc.after_end = st.after_end; sc.qn = qn; }
            __syncwarp();
        }

        if (!steady) {
            const uint64_t nb = sc.chunk0 + ((uint64_t)(it + 1) << 10);
            if (nb >= sc.ge) { at_eof = true; break; }          // genome exhausted
            if (nb >= sc.end) {                                 // run-out: stop when no owned k-mer can still end
                const uint32_t after_end = sc.after_end;
                if (after_end >= (uint32_t)(TL - 1) || since_break <= after_end) break;
            }
        }
    }
    {
        if (ln) {
            uint32_t qn = sc.qn;
            while (ln) {                                              // one block hit per parked lane and pass
                const uint32_t m = ln < 32u ? ln : 32u;
                ln = drain3<ST>(P, A, q, qn, lq, ln - m, m, sc.gid, sc.chunk0 - sc.gs);
            }
            __syncwarp();
            if (lane == 0) sc.qn = qn;
            __syncwarp();
        }
    }
    if (hdr && at_eof && lane == 0) atomicOr(&A.gstatus[sc.gid], 1);   // the text ended inside a '>' line
}

template <int ST, bool BIG>
#ifdef KSSD_SCAN_MAXNREG
__global__ void __maxnreg__(KSSD_SCAN_MAXNREG) sketch_fasta3_kernel(
#else
__global__ void __launch_bounds__(kScanThreads, 1) sketch_fasta3_kernel(
#endif
    const __grid_constant__ SketchParams P, const __grid_constant__ ScanArgs A, const uint32_t *__restrict__ pf_global)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *pf = reinterpret_cast<uint32_t *>(smem_raw);
    WarpQ3 *queues = reinterpret_cast<WarpQ3 *>(smem_raw + kPf3Words * 4);
    LaneQ3 *lqueues = reinterpret_cast<LaneQ3 *>(queues + kScanWarps);
    SpanCtx *spans = reinterpret_cast<SpanCtx *>(lqueues + kScanWarps);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(pf_global);
        uint4 *dst = reinterpret_cast<uint4 *>(pf);
        for (uint32_t i = threadIdx.x; i < kPf3Words / 4; i += blockDim.x) dst[i] = __ldg(&src[i]);
    }
    __syncthreads();
    WarpQ3 &q = queues[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    if (lane == 0) spans[threadIdx.x >> 5].qn = 0;       // candidates waiting in the warp's queue (it outlives spans)
    __syncwarp();
    for (;;) {
        uint32_t si = 0;
        if (lane == 0) si = atomicAdd(A.ticket, 1u);
        si = __shfl_sync(kFull, si, 0);
        if (si >= A.n_spans) break;
        const uint32_t gid = A.span_gid[si];
        const uint64_t gs = A.goff[gid], ge = gs + A.glen[gid];
        uint64_t start, end;
        if (!span_extent(A, si, gid, gs, ge, start, end)) continue;
        SpanCtx &sc = spans[threadIdx.x >> 5];
        __syncwarp();
        if (lane == 0) { sc.gs = gs; sc.ge = ge; sc.start = start; sc.end = end; sc.gid = gid; }
        __syncwarp();
#ifdef KSSD_SCAN_TMA
        scan_span3<ST, BIG>(P, A, pf, q, lqueues[threadIdx.x >> 5], sc, reinterpret_cast<TmaRing *>(spans + kScanWarps)[threadIdx.x >> 5]);
#else
        scan_span3<ST, BIG>(P, A, pf, q, lqueues[threadIdx.x >> 5], sc);
#endif
    }
    {
        __syncwarp();
        const uint32_t qn = reinterpret_cast<volatile SpanCtx *>(spans)[threadIdx.x >> 5].qn;
        if (qn) resolve3(P, A, q, 0, qn);
    }
}

}  // namespace kssd
