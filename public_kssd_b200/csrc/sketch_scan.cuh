// sketch_scan.cuh -- Stage I scan kernel: FASTA text -> sampled (component, genome, id) occurrences.
//
// Replaces the byte-serial loop of fasta2co / uniq_fasta2co (reference iseq2comem.c:205-270,
// :633-700).  Semantics reproduced exactly:
//   ACGTacgt -> 0..3 and extend the run; '\n' '\r' are skipped WITHOUT breaking the run; '>' opens a
//   header that is skipped through the next '\n' and breaks the run; every other byte breaks the
//   run.  A 2k-mer is considered once 2k valid bases have been seen since the last break; its
//   canonical form min(fwd, revcomp) is sampled iff shuf[central 2s-mer] < dim_end, and re-encoded
//   as drtuple = [left outer | right outer | pf].
//
// B200 mapping (DESIGN.md "Stage I"):
//   * one persistent CTA per SM, the 128 KiB prefilter bitmap resident in shared memory;
//   * a WARP is the unit of streaming: it pulls a span (a run of whole lines of one genome) from an
//     atomic ticket and walks it 1 KiB per iteration, 32 B per lane (sketch_scan32.cuh), with the next
//     KiB already in flight (register double buffering); no block-level barrier in the steady state;
//   * the central 2s-mer of the FORWARD strand is tested against the bitmap of S u RC(S): the
//     central window of the reverse complement is the reverse complement of the central window,
//     so the canonical choice cannot create a hit the forward window does not announce.  Only
//     ~1/128 of the positions survive; they are queued per warp and resolved densely (revcomp,
//     canonical min, exact lookup in the sampled-set hash table, drtuple) 32 at a time;
//   * k-mers are OWNED by the span their first base lies in; a warp runs past the end of its span
//     until 2k-1 valid bases or a break, so spans never exchange state.
//
// The kernel is instruction bound (profiles/r1_sketch_ncu_summary.md): every choice below is about the
// number of ALU-pipe instructions per input byte, not about bytes moved.
//
// This header holds the pieces shared by the scan kernels: byte classification, the squeeze, the prefilter
// probes, the candidate queue and its exact resolver, span boundaries, and the general (dirty) iteration.
// The clean-path loop and the kernel live in sketch_scan32.cuh; the FASTQ walk in sketch_fastq.cuh.
#pragma once
#include "kssd_device.cuh"

namespace kssd {

#ifndef KSSD_SCAN_THREADS
#define KSSD_SCAN_THREADS 512
#endif
constexpr int kScanThreads = KSSD_SCAN_THREADS; // warps per SM = threads / 32 (one CTA per SM)
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kQueueCap = 64;                   // per-warp candidate stack entries
constexpr uint64_t kNoSpan = ~0ull;
constexpr uint32_t kRunCap = 64;                // saturation of "valid bases since last break"

struct ScanArgs {
    const uint8_t *seq;          // batch text
    uint64_t seq_bytes;          // readable bytes
    const uint64_t *goff;        // per genome offset  (device)
    const uint64_t *glen;        // per genome length  (device)
    const uint32_t *span_gid;    // per span genome id (device)
    const uint64_t *span_nom;    // per span nominal start (absolute offset)
    uint32_t n_spans;
    uint32_t *ticket;            // span dispenser
    uint64_t *out_keys;          // (comp << 56) | (gid << 28) | id
    uint64_t *out_ords;          // byte offset of the occurrence inside its genome (monotone in stream order)
    uint32_t out_cap;
    uint32_t *out_count;
    int32_t *gstatus;            // per genome flags (bit 0: header ran into EOF)
    uint32_t *zero_count;        // per genome: occurrences of code 0 dropped by the FASTA quirk (they still count as keys)
    int drop_zero;               // FASTA quirk: drtuple == 0 is never stored (iseq2comem.c:258)
    // bucket mode (sketch_scan3.cuh resolver, csrc/sketch_buckets.cuh): occurrences go straight to the (component, genome)
    // bucket they belong to -- bkeys[boff[b] + n] = id << 36 | offset -- instead of one list that needs a global sort
    uint64_t *bkeys;             // null = list mode
    const uint32_t *boff;        // bucket starts (n_buckets + 1), capacities from the genome lengths
    uint32_t *bcnt;              // occurrences per bucket (may exceed the capacity: then *boverflow is set)
    uint32_t *boverflow;
    uint32_t n_genomes;
    int strict_window;           // FASTQ reads: the final check wants the 2k BYTES ending at the position to be letters (no line ends inside)
};

struct WarpQueue {
    uint32_t lo[kQueueCap];
    uint32_t hi[kQueueCap];
    uint32_t ord[kQueueCap];
};

// Lanes of a clean iteration whose first-level probe hit something park here -- their four base words, the hit mask and
// their offset -- until 32 of them are waiting; the second-level probe and the hand-over to the exact queue then run
// with every lane busy instead of once per iteration for the few lanes that hit.
struct LaneQueue {
    uint32_t w0[kQueueCap], w1[kQueueCap], w2[kQueueCap], w3[kQueueCap], cand[kQueueCap], off[kQueueCap];
};
constexpr size_t kScanSmemBytes = (size_t)(kPfWords + kPf2Words) * 4 + (size_t)kScanWarps * (sizeof(WarpQueue) + sizeof(LaneQueue));

// ---- byte classification, 4 bytes at a time (bit tricks checked exhaustively in tests/test_bittricks.py) ----
//   dacc |= nonzero byte  <=>  that byte is NOT one of ACGTacgt, '\n', '\r'
//   t3   : per byte, bits 0-1 = (b >> 1) & 3 (A0 C1 T2 G3), bit 2 = bit 3 of b -- clear in every letter, set in '\n' and
//          '\r' (a skip byte when clean); the three bits are adjacent in b, so one AND of w >> 1 yields them and they
//          index the eight-entry PRMT table of the bytes a clean byte must equal
//   m    : top byte = the four 2-bit codes A0 C1 G2 T3, first byte in the top two bits
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &dacc, uint32_t &t3, uint32_t &m)
{
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    t3 = s1 & 0x07070707u;
    const uint32_t u = w & ~(s1 & 0x20202020u);                   // fold case of letters only
    const uint32_t a = t3 | (t3 >> 4);
    const uint32_t sel = prmt(a, 0u, 0x4420u);
    const uint32_t e = prmt(0x47544341u, 0xFF0D0AFFu, sel);        // A C T G | - \n \r -
    dacc |= u ^ e;
    const uint32_t t2 = (s1 ^ s2) & 0x03030303u;                  // bit 0 = b1 ^ b2, bit 1 = b2 ^ b3 = b2 for letters
    m = t2 * 0x40100401u;
}

// 8 bits from two t3 words (tOld = the older four bytes): bit (7 - b) = flag (t3 bit 2) of byte b of the eight,
// i.e. reversed order, newest byte in bit 0.  The multiplier both gathers and reverses the flags.
__device__ __forceinline__ uint32_t rev_flags8(uint32_t tOld, uint32_t tNew)
{
    const uint32_t g = ((tNew >> 2) | (tOld << 2)) & 0x11111111u;
    return (g * 0x08040201u) >> 24;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
    return ((uint64_t)__shfl_sync(kFull, (uint32_t)(v >> 32), src) << 32) | __shfl_sync(kFull, (uint32_t)v, src);
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d)
{
    return ((uint64_t)__shfl_up_sync(kFull, (uint32_t)(v >> 32), d) << 32) | __shfl_up_sync(kFull, (uint32_t)v, d);
}

// (hi:lo) << s for s in [0, 32]: the three result words; clamp-mode funnel shifts make s == 32 exact
__device__ __forceinline__ void shl96(uint32_t lo, uint32_t hi, uint32_t s, uint32_t &r0, uint32_t &r1, uint32_t &r2)
{
    r0 = __funnelshift_lc(0u, lo, s);
    r1 = __funnelshift_lc(lo, hi, s);
    r2 = __funnelshift_lc(hi, 0u, s);
}

// A span starts right after the first '\n' at or after nominal-1 (or at the genome start); the search stops at `lim`
// (the next span's nominal-1, or the genome end).  First probe: 128 bytes at once, four per lane -- enough for any
// ordinary line width; longer lines take the 32-bytes-per-round loop.
__device__ __forceinline__ uint64_t find_span_start(const uint8_t *seq, uint64_t gs, uint64_t ge, uint64_t nominal, uint64_t lim)
{
    if (nominal <= gs) return gs;
    const uint32_t lane = lane_id();
    if (lim > ge) lim = ge;
    uint64_t p = nominal - 1;
    {
        uint32_t hit = 0;                       // bit j: byte p + 4*lane + j is a newline inside [p, lim)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t a = p + 4 * lane + j;
            hit |= (uint32_t)(a < lim && seq[a] == '\n') << j;
        }
        const uint32_t m = __ballot_sync(kFull, hit != 0);
        if (m) {
            const int l = __ffs(m) - 1;
            const uint32_t h = __shfl_sync(kFull, hit, l);
            const uint64_t st = p + 4 * l + (__ffs(h) - 1) + 1;
            return st < ge ? st : kNoSpan;
        }
        p += 128;
    }
    for (; p < lim; p += 32) {
        const uint64_t a = p + lane;
        const bool nl = (a < lim) && (seq[a] == '\n');
        const uint32_t m = __ballot_sync(kFull, nl);
        if (m) {
            const uint64_t st = p + (__ffs(m) - 1) + 1;
            return st < ge ? st : kNoSpan;
        }
    }
    return kNoSpan;
}

// start and end of span `si` (its end is the start of the next non-empty span of the same genome)
__device__ __forceinline__ bool span_extent(const ScanArgs &A, uint32_t si, uint32_t gid, uint64_t gs, uint64_t ge, uint64_t &start, uint64_t &end)
{
    auto next_nom = [&](uint32_t j) -> uint64_t {      // where span j's search must stop
        return (j + 1 < A.n_spans && A.span_gid[j + 1] == gid) ? A.span_nom[j + 1] - 1 : ge;
    };
    start = find_span_start(A.seq, gs, ge, A.span_nom[si], next_nom(si));
    if (start == kNoSpan) return false;
    end = ge;
    for (uint32_t j = si + 1; j < A.n_spans && A.span_gid[j] == gid; j++) {
        const uint64_t e = find_span_start(A.seq, gs, ge, A.span_nom[j], next_nom(j));
        if (e != kNoSpan) { end = e; break; }
    }
    return true;
}

// first-level prefilter probe of a window value v (layout in kssd_device.cuh); returns the flag in bit 31
__device__ __forceinline__ uint32_t pf_probe(const uint32_t *__restrict__ pf, uint32_t v)
{
    return __funnelshift_l(0u, pf[v & kPfWordMask], v >> kPfBitShift);
}
// second-level probe of the exact inner 2s-mer
__device__ __forceinline__ bool pf2_probe(const uint32_t *__restrict__ pf, uint32_t inner)
{
    const uint32_t i = pf2_index(inner);
    return (pf[kPfWords + (i >> 5)] >> (i & 31)) & 1u;
}

// ---- exact resolution of queued candidates (dense: up to 32 at a time) ----
__device__ __forceinline__ void resolve_candidates(const SketchParams &P, const ScanArgs &A, WarpQueue &q, uint32_t first,
                                                   uint32_t m, uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    bool found = false;
    uint64_t key = 0, ordv = 0;
    if (lane < m) {
        const uint64_t fwd = ((uint64_t)q.hi[first + lane] << 32) | q.lo[first + lane];
        ordv = ord_base + q.ord[first + lane];
        const uint64_t rc = revcomp2(fwd, P.TL);
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
            if (A.drop_zero && dr == 0) { found = false; atomicAdd(&A.zero_count[gid], 1u); }
            key = ((dr & P.comp_mask) << 56) | ((uint64_t)gid << 28) | (dr >> P.comp_code_bits);
        }
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

// push the candidates flagged in `cand` (bit d = k-mer ending d valid bases before the lane's newest
// base); W3:W2:W1:W0 holds the lane's history + own bases, newest base in the low bits.
// vmask: 0 for a clean lane (ord sub-index = span-1-d, span = 16 or 32 bytes per lane), else the lane's
// effective-valid byte mask (16-byte general path).
__device__ __forceinline__ void push_candidates(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQueue &q,
                                                uint32_t &qn, uint32_t cand, uint32_t n, uint32_t W0, uint32_t W1, uint32_t W2, uint32_t W3,
                                                uint32_t lane_off, uint32_t vmask, uint32_t span, uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    while (__any_sync(kFull, cand != 0)) {
        bool has = cand != 0;
        int d = 0;
        uint32_t lo = 0, hi = 0;
        if (has) {
            d = __ffs(cand) - 1;
            cand &= cand - 1;
            const bool up = d >= 16;                       // k-mer starts in the upper three words
            const uint32_t a0 = up ? W1 : W0, a1 = up ? W2 : W1, a2 = up ? W3 : W2;
            const uint32_t sh = 2 * (d & 15);
            lo = __funnelshift_r(a0, a1, sh);
            hi = __funnelshift_r(a1, a2, sh);
            // second-level filter on the forward inner 2s-mer: 15 of 16 first-level hits stop here
            has = pf2_probe(pf, __funnelshift_r(lo, hi, 2 * P.out) & P.innermask);
        }
        const uint32_t pm = __ballot_sync(kFull, has);
        if (has) {
            const uint64_t fwd = (((uint64_t)hi << 32) | lo) & P.tupmask;
            uint32_t sub;
            if (vmask == 0) sub = span - 1 - d;
            else sub = __fns(vmask, 0, (int)(n - d));          // byte index of the (n-d)-th valid base
            const uint32_t slot = qn + __popc(pm & ((1u << lane) - 1u));
            q.lo[slot] = (uint32_t)fwd;
            q.hi[slot] = (uint32_t)(fwd >> 32);
            q.ord[slot] = lane_off + sub;
        }
        qn += __popc(pm);
        __syncwarp();
        if (qn >= 32) {
            resolve_candidates(P, A, q, qn - 32, 32, gid, ord_base);
            qn -= 32;
            __syncwarp();
        }
    }
}

// guarded 16-byte load for the first / last chunks of a buffer
__device__ __forceinline__ uint4 load_chunk16_guarded(const ScanArgs &A, uint64_t addr)
{
    if (addr + 16 <= A.seq_bytes) return ldg_stream(reinterpret_cast<const uint4 *>(A.seq + addr));
    uint32_t w[4] = {0x0d0d0d0du, 0x0d0d0d0du, 0x0d0d0d0du, 0x0d0d0d0du};
    for (int i = 0; i < 16; i++)
        if (addr + i < A.seq_bytes) {
            w[i >> 2] = (w[i >> 2] & ~(0xffu << (8 * (i & 3)))) | ((uint32_t)A.seq[addr + i] << (8 * (i & 3)));
        }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// bytes [0,lo) and [hi,16) of the lane become '\r' (skipped, not a line end)
__device__ __forceinline__ void mask_lane_bytes(uint4 &q, int lo, int hi)
{
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 16; i++)
        if (i < lo || i >= hi) w[i >> 2] = (w[i >> 2] & ~(0xffu << (8 * (i & 3)))) | (0x0du << (8 * (i & 3)));
    q = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ int clamp16(int64_t v) { return v < 0 ? 0 : (v > 16 ? 16 : (int)v); }

// Warp state carried from one 512-byte iteration to the next.
struct StreamState {
    uint64_t cw;           // most recent valid bases (>= TL-1 of them are meaningful when run allows)
    uint32_t since_break;  // valid bases since the last break, saturating at kRunCap
    uint32_t after_end;    // valid bases at offsets >= span end seen so far (run-out accounting)
    uint32_t hdr;          // inside a '>' header line
};

// One 512-byte GENERAL iteration (16 bytes per lane): headers, N, IUPAC, anything.  `cur` is already masked to the
// span / genome extent, `codes` are its 2-bit codes (byte 0 in the top two bits).  Rare, so kept out of line.
__device__ __noinline__ void general_iter16(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQueue &q, uint32_t &qn,
                                            StreamState &st, uint4 cur, uint32_t codes, uint64_t cbase, uint64_t end, bool past_end,
                                            uint32_t lane_off, uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    const int TL = P.TL;
    uint32_t cand = 0, W0, W1, W2, vmask = 0, n;
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
    vmask = Veff;
    n = __popc(Veff);
    // lane summary: bases after the lane's last break (tail) and all effective bases (Pl)
    uint32_t tb = 0, tn = 0, Pl = 0;
    bool hb = false;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t c = (codes >> (30 - 2 * i)) & 3u;
        if ((Veff >> i) & 1u) { tb = (tb << 2) | c; tn++; Pl = (Pl << 2) | c; }
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
    // walk the lane's bytes
    uint64_t fwd = hist;
    uint32_t j = 0;   // valid bases consumed in this lane
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if ((Veff >> i) & 1u) {
            fwd = (fwd << 2) | ((codes >> (30 - 2 * i)) & 3u);
            run = min(run + 1, kRunCap);
            ae += (gem >> i) & 1u;
            j++;
            if (run >= (uint32_t)TL && ae <= (uint32_t)(TL - 1)) {
                if (pf_probe(pf, (uint32_t)(fwd >> (2 * P.out))) >> 31) cand |= 1u << (n - j);
            }
        } else if ((BRK >> i) & 1u) run = 0;
    }
    shl96((uint32_t)hist, (uint32_t)(hist >> 32), 2 * n, W0, W1, W2);
    W0 |= Pl;
    // warp carry = inclusive value of lane 31 on top of the old carry
    const uint64_t sb31 = shfl64(sb, 31);
    const uint32_t sn31 = __shfl_sync(kFull, sn, 31), sbrk31 = __shfl_sync(kFull, sbrk, 31);
    if (sbrk31) { st.cw = sb31; st.since_break = sn31; }
    else { st.cw = (sn31 < 32 ? (st.cw << (2 * sn31)) : 0ull) | sb31; st.since_break = min(st.since_break + sn31, kRunCap); }
    st.after_end += __shfl_sync(kFull, ginc, 31);
    push_candidates(P, A, pf, q, qn, cand, n, W0, W1, W2, 0u, lane_off, vmask, 16u, gid, ord_base);
}

}  // namespace kssd
