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
//     atomic ticket and walks it 512 B per iteration, 16 B per lane, with the next 512 B already
//     in flight (register double buffering); no block-level barrier in the steady state;
//   * the central 2s-mer of the FORWARD strand is tested against the bitmap of S u RC(S): the
//     central window of the reverse complement is the reverse complement of the central window,
//     so the canonical choice cannot create a hit the forward window does not announce.  Only
//     ~1/128 of the positions survive; they are queued per warp and resolved densely (revcomp,
//     canonical min, exact lookup in the sampled-set hash table, drtuple) 32 at a time;
//   * k-mers are OWNED by the span their first base lies in; a warp runs past the end of its span
//     until 2k-1 valid bases or a break, so spans never exchange state.
#pragma once
#include "kssd_device.cuh"

namespace kssd {

constexpr int kScanThreads = 512;               // 16 warps per SM
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
    uint32_t span_bytes;
    uint32_t *ticket;            // span dispenser
    uint64_t *out_keys;          // (comp << 56) | (gid << 28) | id
    uint64_t *out_ords;          // byte offset of the occurrence inside its genome (monotone in stream order)
    uint32_t out_cap;
    uint32_t *out_count;
    int32_t *gstatus;            // per genome flags (bit 0: header ran into EOF)
    int drop_zero;               // FASTA quirk: drtuple == 0 is never stored (iseq2comem.c:258)
};

struct WarpQueue {
    uint32_t lo[kQueueCap];
    uint32_t hi[kQueueCap];
    uint32_t ord[kQueueCap];
};

// ---- byte classification, 4 bytes at a time (verified exhaustively, see tests/test_bittricks.py) ----
// diff byte == 0  <=>  byte is one of ACGTacgt, '\n', '\r'
// sc bit 6 of a byte set  <=> byte has bit 6 clear (for a clean byte: it is '\n' or '\r')
// p8 = the four 2-bit codes, first byte in the top two bits
__device__ __forceinline__ void classify4(uint32_t w, uint32_t &diff, uint32_t &sc, uint32_t &p8)
{
    const uint32_t t3 = ((w >> 1) & 0x03030303u) | ((~w >> 4) & 0x04040404u);
    const uint32_t u = w & ~((w >> 1) & 0x20202020u);           // fold case of letters only
    const uint32_t a = t3 | (t3 >> 4);
    const uint32_t sel = prmt(a, 0u, 0x4420u);
    const uint32_t e = prmt(0x47544341u, 0xFF0D0AFFu, sel);      // A C T G | - \n \r -
    diff = u ^ e;
    sc = ~w & 0x40404040u;
    const uint32_t t = t3 & 0x03030303u;
    const uint32_t t2 = t ^ ((t >> 1) & 0x01010101u);            // A0 C1 T2 G3 -> A0 C1 G2 T3
    p8 = (t2 * 0x40100401u) >> 24;
}

// 16 bits, bit b = byte b of the lane's 16 bytes has bit 6 clear (natural order)
__device__ __forceinline__ uint32_t skip_mask16(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3)
{
    const uint32_t g01 = (s0 >> 6) | (s1 >> 2);
    const uint32_t g23 = (s2 >> 6) | (s3 >> 2);
    return (((g01 * 0x00204081u) >> 21) & 0xffu) | (((g23 * 0x00204081u) >> 13) & 0xff00u);
}

// remove the 2-bit groups flagged in m (bit b = group of byte b, byte 0 in the top bits);
// the survivors end up right-aligned
__device__ __forceinline__ uint32_t squeeze_groups(uint32_t c, uint32_t m)
{
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t low = (1u << (30 - 2 * b)) - 1u;
        c = ((c >> 2) & ~low) | (c & low);
    }
    return c;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src)
{
    return ((uint64_t)__shfl_sync(kFull, (uint32_t)(v >> 32), src) << 32) | __shfl_sync(kFull, (uint32_t)v, src);
}
__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int d)
{
    return ((uint64_t)__shfl_up_sync(kFull, (uint32_t)(v >> 32), d) << 32) | __shfl_up_sync(kFull, (uint32_t)v, d);
}

// first '\n' at or after nominal-1, searched over at most span_bytes bytes; returns the offset just
// after it (a span always starts right after a newline, or at the genome start).
__device__ __forceinline__ uint64_t find_span_start(const uint8_t *seq, uint64_t gs, uint64_t ge, uint64_t nominal,
                                                    uint32_t span_bytes)
{
    if (nominal <= gs) return gs;
    const uint32_t lane = lane_id();
    uint64_t lim = nominal - 1 + span_bytes;
    if (lim > ge) lim = ge;
    for (uint64_t p = nominal - 1; p < lim; p += 32) {
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
            if (A.drop_zero && dr == 0) found = false;
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
// base); W2:W1:W0 holds the lane's history + own bases, newest base in the low bits.
// vmask: 0 for the clean path (ord sub-index = 15-d), else the lane's effective-valid byte mask.
__device__ __forceinline__ void push_candidates(const SketchParams &P, const ScanArgs &A, WarpQueue &q, uint32_t &qn,
                                                uint32_t cand, uint32_t n, uint32_t W0, uint32_t W1, uint32_t W2,
                                                uint32_t lane_off, uint32_t vmask, uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    while (__any_sync(kFull, cand != 0)) {
        const bool has = cand != 0;
        const int d = has ? (__ffs(cand) - 1) : 0;
        cand &= cand - 1;
        const uint32_t pm = __ballot_sync(kFull, has);
        if (has) {
            const uint32_t lo = __funnelshift_r(W0, W1, 2 * d);
            const uint32_t hi = __funnelshift_r(W1, W2, 2 * d);
            const uint64_t fwd = (((uint64_t)hi << 32) | lo) & P.tupmask;
            uint32_t sub;
            if (vmask == 0) sub = 15 - d;
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

__device__ __forceinline__ uint4 load_chunk16(const ScanArgs &A, uint64_t addr)
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

// One span: [start, end) of genome [gs, ge); returns nothing, appends occurrences to the output.
__device__ void scan_span(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQueue &q,
                          uint32_t gid, uint64_t gs, uint64_t ge, uint64_t start, uint64_t end)
{
    const uint32_t lane = lane_id();
    const int TL = P.TL;
    StreamState st = {0ull, 0u, 0u, 0u};
    uint32_t qn = 0;
    const uint64_t chunk0 = start & ~127ull;
    const uint64_t ord_base = chunk0 - gs + 0;   // chunk0 >= gs - 127 can be below gs: handled by signed add below
    uint64_t chunk = chunk0;
    uint4 nxt = load_chunk16(A, chunk + 16 * lane);

    for (;;) {
        uint4 cur = nxt;
        const uint64_t cbase = chunk;
        const uint64_t laddr = cbase + 16 * lane;
        chunk += 512;
        if (chunk < ge) nxt = load_chunk16(A, chunk + 16 * lane);
        if (cbase < start || cbase + 512 > ge)
            mask_lane_bytes(cur, clamp16((int64_t)start - (int64_t)laddr), clamp16((int64_t)ge - (int64_t)laddr));

        uint32_t d0, d1, d2, d3, s0, s1, s2, s3, p0, p1, p2, p3;
        classify4(cur.x, d0, s0, p0);
        classify4(cur.y, d1, s1, p1);
        classify4(cur.z, d2, s2, p2);
        classify4(cur.w, d3, s3, p3);
        const uint32_t codes = (p0 << 24) | (p1 << 16) | (p2 << 8) | p3;   // byte 0 in the top two bits
        const bool dirty = (d0 | d1 | d2 | d3) != 0;
        const uint32_t skm = skip_mask16(s0, s1, s2, s3);
        uint32_t n = 16 - __popc(skm);
        const bool past_end = cbase + 512 > end;
        const uint32_t lane_off = (uint32_t)(cbase - chunk0) + 16 * lane;

        // a lane may be short of bases only where the text itself is cut (before `start`, after the genome end)
        const bool lane_ok = !dirty && (n >= (uint32_t)P.hist_min_n || laddr < start || laddr + 16 > ge);
        const bool clean = __all_sync(kFull, lane_ok) && !st.hdr;
        uint32_t cand = 0, W0, W1, W2, vmask = 0;

        if (clean) {
            // ---------------- clean iteration: only bases and line ends, no header pending ----------------
            const uint32_t Pl = squeeze_groups(codes, skm);
            const uint32_t A1 = __shfl_up_sync(kFull, Pl, 1);
            const uint32_t B2 = __shfl_up_sync(kFull, Pl, 2);
            const uint32_t nA = __shfl_up_sync(kFull, n, 1);
            uint64_t H;
            if (lane == 0) H = st.cw;
            else {
                const uint64_t older = lane >= 2 ? (uint64_t)B2 : st.cw;
                H = (nA >= 16 ? (older << 32) : (older << (2 * nA))) | A1;
            }
            // W = (H << 2n) | Pl, 96 bits
            const uint64_t x0 = (n >= 16) ? ((uint64_t)(uint32_t)H << 32) : ((uint64_t)(uint32_t)H << (2 * n));
            const uint64_t x1 = (n >= 16) ? ((uint64_t)(uint32_t)(H >> 32) << 32) : ((uint64_t)(uint32_t)(H >> 32) << (2 * n));
            W0 = (uint32_t)x0 | Pl;
            W1 = (uint32_t)(x0 >> 32) | (uint32_t)x1;
            W2 = (uint32_t)(x1 >> 32);
            // prefilter on the central 2s-mer of the k-mer ending at each own base
            const uint32_t Xlo = __funnelshift_r(W0, W1, 2 * P.out);
            const uint32_t Xhi = __funnelshift_r(W1, W2, 2 * P.out);
#pragma unroll
            for (int d = 0; d < 16; d++) {
                const uint32_t tmp = __funnelshift_r(Xlo, Xhi, 2 * d) & P.pfmask;
                const uint32_t word = pf[tmp >> 5];
                cand |= ((word >> (tmp & 31)) & 1u) << d;
            }
            cand &= (1u << n) - 1u;

            const uint32_t N = __reduce_add_sync(kFull, n);
            if (st.since_break < (uint32_t)(TL - 1) || past_end) {
                // start of a span / run-out past its end: filter by position inside the iteration
                uint32_t incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                const int o_l = (int)(incl - n);                     // valid bases before this lane
                // ok1: since_break + t_local + 1 >= TL  with t_local = o_l + n-1-d
                const int need = TL - 1 - (int)st.since_break - o_l; // n-1-d >= need
                if (need > 0) {
                    const int keep = (int)n - need;                   // d <= keep-1
                    cand &= keep <= 0 ? 0u : ((1u << keep) - 1u);
                }
                uint32_t E = N;                                      // valid bases of this iteration before `end`
                if (past_end) {
                    // lane holding `end` (or lane 0 when the whole iteration is past it)
                    const int64_t rel = (int64_t)end - (int64_t)cbase;
                    if (rel <= 0) E = 0;
                    else {
                        const int le = (int)(rel >> 4);
                        const int be = (int)(rel & 15);
                        const uint32_t before = (uint32_t)o_l + (uint32_t)__popc(~skm & ((1u << be) - 1u));
                        E = __shfl_sync(kFull, before, le);
                    }
                    // ok2: after_end + (t_local - E + 1) <= TL-1   for t_local >= E
                    const int lim = TL - 2 - (int)st.after_end + (int)E - o_l;   // n-1-d <= lim
                    const int drop = (int)n - 1 - lim;                            // d >= drop
                    if (drop > 0) cand &= drop >= 16 ? 0u : ~((1u << drop) - 1u);
                    st.after_end += N - E;
                }
            }
            st.since_break = min(st.since_break + N, kRunCap);
            // carry: the last two lanes hold at least TL-1 bases
            const uint32_t P30 = __shfl_sync(kFull, Pl, 30), P31 = __shfl_sync(kFull, Pl, 31);
            const uint32_t n31 = __shfl_sync(kFull, n, 31);
            st.cw = (n31 >= 16 ? ((uint64_t)P30 << 32) : ((uint64_t)P30 << (2 * n31))) | P31;
        } else {
            // ---------------- general iteration: headers, N, IUPAC, anything ----------------
            const uint32_t w[4] = {cur.x, cur.y, cur.z, cur.w};
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
                        const uint32_t tmp = (uint32_t)(fwd >> (2 * P.out)) & P.pfmask;
                        if ((pf[tmp >> 5] >> (tmp & 31)) & 1u) cand |= 1u << (n - j);
                    }
                } else if ((BRK >> i) & 1u) run = 0;
            }
            // W = (hist << 2n) | Pl
            const uint64_t x0 = (n >= 16) ? ((uint64_t)(uint32_t)hist << 32) : ((uint64_t)(uint32_t)hist << (2 * n));
            const uint64_t x1 = (n >= 16) ? ((uint64_t)(uint32_t)(hist >> 32) << 32) : ((uint64_t)(uint32_t)(hist >> 32) << (2 * n));
            W0 = (uint32_t)x0 | Pl;
            W1 = (uint32_t)(x0 >> 32) | (uint32_t)x1;
            W2 = (uint32_t)(x1 >> 32);
            // warp carry = inclusive value of lane 31 on top of the old carry
            const uint64_t sb31 = shfl64(sb, 31);
            const uint32_t sn31 = __shfl_sync(kFull, sn, 31), sbrk31 = __shfl_sync(kFull, sbrk, 31);
            if (sbrk31) { st.cw = sb31; st.since_break = sn31; }
            else { st.cw = (sn31 < 32 ? (st.cw << (2 * sn31)) : 0ull) | sb31; st.since_break = min(st.since_break + sn31, kRunCap); }
            st.after_end += __shfl_sync(kFull, ginc, 31);
        }

        push_candidates(P, A, q, qn, cand, n, W0, W1, W2, lane_off, vmask, gid, ord_base);

        if (cbase + 512 >= ge) break;                       // genome exhausted
        if (cbase + 512 >= end) {                           // run-out: stop when no owned k-mer can still end
            if (st.after_end >= (uint32_t)(TL - 1) || st.since_break <= st.after_end) break;
        }
    }
    if (qn) { resolve_candidates(P, A, q, 0, qn, gid, ord_base); __syncwarp(); }
    if (st.hdr && chunk >= ge && lane == 0) atomicOr(&A.gstatus[gid], 1);
}

__global__ void __launch_bounds__(kScanThreads, 1) sketch_fasta_kernel(const SketchParams P, const ScanArgs A)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *pf = reinterpret_cast<uint32_t *>(smem_raw);
    WarpQueue *queues = reinterpret_cast<WarpQueue *>(smem_raw + kPfWords * 4);
    {   // stage the prefilter bitmap (L2 -> shared), 16 B per thread per step
        const uint4 *src = reinterpret_cast<const uint4 *>(P.prefilter);
        uint4 *dst = reinterpret_cast<uint4 *>(pf);
        for (uint32_t i = threadIdx.x; i < kPfWords / 4; i += blockDim.x) dst[i] = __ldg(&src[i]);
    }
    __syncthreads();
    WarpQueue &q = queues[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    for (;;) {
        uint32_t si = 0;
        if (lane == 0) si = atomicAdd(A.ticket, 1u);
        si = __shfl_sync(kFull, si, 0);
        if (si >= A.n_spans) break;
        const uint32_t gid = A.span_gid[si];
        const uint64_t gs = A.goff[gid], ge = gs + A.glen[gid];
        const uint64_t start = find_span_start(A.seq, gs, ge, A.span_nom[si], A.span_bytes);
        if (start == kNoSpan) continue;
        uint64_t end = ge;
        for (uint32_t j = si + 1; j < A.n_spans && A.span_gid[j] == gid; j++) {
            const uint64_t e = find_span_start(A.seq, gs, ge, A.span_nom[j], A.span_bytes);
            if (e != kNoSpan) { end = e; break; }
        }
        scan_span(P, A, pf, q, gid, gs, ge, start, end);
    }
}

}  // namespace kssd
