// sketch_fastq3.cuh -- FASTQ read walk, warp formulation: the block prefilter and the lazy codes of sketch_scan3.cuh applied to reads.
//
// Replaces the per-read loops of fastq2co / mt_shortreads2koc (reference iseq2comem.c:277-356, :554-615) whenever the quality bytes
// cannot change the outcome: -A (quality ignored), or -Q <= 0 in a file whose single-pass line index saw no byte >= 0x80 (a quality
// byte fails -Q <= 0 only as a negative signed char).  -Q > 0 and files beyond the single-pass index keep the thread-per-read walk
// of sketch_fastq.cuh, which was as ALU bound as round 1's FASTA scan (exact classification, one probe per base, ~2000 instructions
// per read).
//
// A warp takes 32 records at a time and lane l walks record l: it frames the record from the line index (positions of its five line
// ends, requested two batches ahead; the record rules and the fgets-length check are the ones of sketch_fastq_kernel) and cuts the
// sequence line into PIECES of 32 text-aligned bytes (the first one starts at s0 & ~31: `lead` bytes of it belong to the header
// line).  Round p of the batch handles piece p of every lane's read -- one aligned 32-byte load per lane, requested a round ahead (the
// pieces of the next batch are asked into L2 a batch ahead) -- so the history of a piece, the last 2k-1 bases before it, is the
// lane's own previous piece: no shuffle, no search.  From there on a piece is a lane of the FASTA scan's steady loop without skipped
// bytes: codes from bits 1-2 (no exact classification: an N is a fake base until the end), 12 probes of the block bitmap per 32
// bases, lanes with a block hit parked and drained 32 at a time through the block table, candidates resolved exactly.  Which k-mers
// of a piece may count is a window mask from the read's extent alone: first base at or after s0, last base before the line end.
// Reads of one sequencing run have one length, so the lanes of a round are all busy; mixed lengths idle the lanes of the shorter reads.
// The final check is against the text: the 2k bytes ending at the position are all letters (strict: a read knows no line ends).
#pragma once
#include "sketch_fastq.cuh"
#include "sketch_scan3.cuh"

namespace kssd {

constexpr size_t kFq3SmemBytes = (size_t)kPf3Words * 4 + (size_t)kScanWarps * (sizeof(WarpQ3) + sizeof(LaneQ3));

template <int ST, bool BIG>
__global__ void __launch_bounds__(kScanThreads, 1) sketch_fastq3_kernel(const __grid_constant__ SketchParams P, const __grid_constant__ ScanArgs A,
                                                                           const __grid_constant__ FastqArgs Fq, const uint32_t *__restrict__ pf_global)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *pf = reinterpret_cast<uint32_t *>(smem_raw);
    WarpQ3 *queues = reinterpret_cast<WarpQ3 *>(smem_raw + kPf3Words * 4);
    LaneQ3 *lqueues = reinterpret_cast<LaneQ3 *>(queues + kScanWarps);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(pf_global);
        uint4 *dst = reinterpret_cast<uint4 *>(pf);
        for (uint32_t i = threadIdx.x; i < kPf3Words / 4; i += blockDim.x) dst[i] = __ldg(&src[i]);
    }
    __syncthreads();
    if (Fq.idx->overflow) return;                                          // the host redoes this file with the two-pass index
    if (!Fq.abund && Fq.Q > -128 && Fq.idx->highbit) return;               // a quality byte may fail -Q <= 0: the thread-per-read walk reads them
    constexpr int NPROBE = ST == 3 ? 12 : 32;
    const uint32_t lane = lane_id(), wid = threadIdx.x >> 5;
    WarpQ3 &q = queues[wid];
    LaneQ3 &lq = lqueues[wid];
    const int TL = P.TL;
    const uint32_t hsh = (2u * (uint32_t)(TL - 1)) & 31u;
    const uint64_t n_nl = (uint64_t)Fq.idx->n_nl;
    const uint64_t n_lines = n_nl + (A.seq[Fq.ge - 1] != '\n' ? 1u : 0u);
    const uint64_t n_records = (n_lines + 3) / 4, n_batches = (n_records + 31) / 32;
    const uint64_t base = Fq.pos_base & ~31ull;                             // piece starts are kept as 32-bit offsets from here
    const uint64_t ord_base = base - Fq.gs;                                 // (may wrap below zero: occurrences add back past it)
    uint32_t qn = 0, ln = 0;

    // the five line ends that frame record r (lines 4r-1 .. 4r+3), as offsets; asked for two batches ahead of their use
    auto fetch_nl = [&](uint64_t bb, uint32_t (&v)[5]) {
        const uint64_t r = 32 * bb + lane;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint64_t idx = 4 * r + k;                                  // line 4r - 1 + k
            v[k] = (bb < n_batches && idx >= 1 && idx - 1 < n_nl) ? __ldg(&Fq.nlpos32[idx - 1]) : 0u;
        }
    };
    // record r = 32 bb + lane from its line ends: first piece (offset from `base`), header bytes in it, bases, pieces; the record rules
    // and the fgets-length check of sketch_fastq_kernel
    auto frame = [&](uint64_t bb, const uint32_t (&v)[5], uint32_t &start_o, uint32_t &lead_o, uint32_t &len_o, uint32_t &pieces_o) {
        const uint64_t r = 32 * bb + lane;
        uint64_t s0 = 0, len = 0;
        const uint64_t l_seq = 4 * r + 1, l_q = 4 * r + 3;
        auto NL5 = [&](int k) -> uint64_t { return Fq.pos_base + v[k]; };                // line 4r - 1 + k (valid where the rules below look)
        if (bb < n_batches && r < n_records && l_seq < n_lines) {
            const bool process = Fq.abund ? (4 * r + 4 <= n_lines) : (r == 0 || 4 * r + 4 <= n_nl);      // iseq2comem.c:567 / :300-307
            if (process) {
                s0 = NL5(1) + 1;                                              // line 4r (the header) is terminated: l_seq < n_lines
                const uint64_t s1 = l_seq < n_nl ? NL5(2) : Fq.ge;
                uint64_t qlen = 0;
                if (l_q < n_lines) {
                    const uint64_t q0 = NL5(3) + 1;
                    qlen = l_q < n_nl ? NL5(4) + 1 - q0 : Fq.ge - q0;
                }
                const uint64_t h0 = r == 0 ? Fq.gs : NL5(0) + 1;
                const uint64_t hlen = NL5(1) - h0;
                const uint64_t plen = (l_seq + 1 < n_nl) ? NL5(3) - (NL5(2) + 1) : 0;
                len = s1 - s0;
                if (len > Fq.line_cap || hlen > Fq.line_cap || plen > Fq.line_cap || qlen > (uint64_t)Fq.line_cap + 1) {
                    atomicOr(&A.gstatus[Fq.gid], 2);                          // the reference would mis-frame every later record
                    len = 0;
                }
                if (len < (uint64_t)TL) len = 0;                              // no k-mer fits
            }
        }
        lead_o = (uint32_t)(s0 & 31);
        start_o = (uint32_t)((s0 & ~31ull) - base);
        len_o = (uint32_t)len;
        pieces_o = len ? (uint32_t)((lead_o + len + 31) >> 5) : 0u;
    };
    auto load_piece = [&](uint32_t off) -> Bytes32 {
        const uint64_t addr = base + off;
        if (addr + 32 <= A.seq_bytes) return ldg_stream256(A.seq + addr);
        return load_chunk32_guarded(A, addr);
    };

    // A lane walks ITS record of the batch, one piece per round: no search for the piece's read, and the history of piece p is the
    // lane's own piece p - 1 -- no shuffle.  Rounds of a batch = the longest read's pieces (reads of one run have one length).
    const uint64_t b_step = (uint64_t)gridDim.x * kScanWarps;
    uint64_t b = (uint64_t)blockIdx.x * kScanWarps + wid;
    uint32_t nlA[5], nlB[5];
    fetch_nl(b, nlA);
    fetch_nl(b + b_step, nlB);
    uint32_t start, lead, len, pieces;
    frame(b, nlA, start, lead, len, pieces);
    for (; b < n_batches; b += b_step) {
        // the next batch: framed now (its line ends were asked for a batch ago), its first pieces asked into L2, the line ends of the
        // batch after it requested
        uint32_t start_n, lead_n, len_n, pieces_n;
        frame(b + b_step, nlB, start_n, lead_n, len_n, pieces_n);
        for (uint32_t pp = 0; pp < pieces_n && pp < 8; pp++)                    // (a batch takes longer than a DRAM access: its text is in L2 when asked for)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.seq + base + start_n + 32u * pp));
        fetch_nl(b + 2 * b_step, nlB);
        const uint32_t rounds = __reduce_max_sync(kFull, pieces);
        Bytes32 nxt;
        nxt.lo = make_uint4(0, 0, 0, 0); nxt.hi = nxt.lo;
        if (pieces) nxt = load_piece(start);
        uint32_t Sp0 = 0, Sp1 = 0;                                          // the last 2k-1 bases of the lane's previous piece
        for (uint32_t p = 0; p < rounds; p++) {
            const Bytes32 cur = nxt;
            const bool active = p < pieces;
            const uint32_t off = start + 32u * p;
            if (p + 1 < pieces) nxt = load_piece(off + 32u);                 // the next round's text is requested before this one is used
            // own byte j ends a k-mer of the read iff the k-mer starts at or after the read's first base and j lies before the line end
            uint32_t wm = 0;
            if (active) {
                const int shift = (int)(32u * p) - (int)lead;                // read position of the piece's byte 0
                const int lo = TL - 1 - shift, hi = (int)len - shift;
                wm = low_mask(hi < 0 ? 0 : (hi > 32 ? 32 : hi)) & ~low_mask(lo < 0 ? 0 : (lo > 32 ? 32 : lo));
            }
            // ---- codes (lazy: bits 1-2 of every byte), history, Y / X, probes: the steady loop of the FASTA scan without skipped bytes ----
            uint32_t Q0, Q1;
            {
                const uint32_t c0 = (cur.lo.x & 0x06060606u) * 0x00820820u, c1 = (cur.lo.y & 0x06060606u) * 0x00820820u;
                const uint32_t c2 = (cur.lo.z & 0x06060606u) * 0x00820820u, c3 = (cur.lo.w & 0x06060606u) * 0x00820820u;
                const uint32_t c4 = (cur.hi.x & 0x06060606u) * 0x00820820u, c5 = (cur.hi.y & 0x06060606u) * 0x00820820u;
                const uint32_t c6 = (cur.hi.z & 0x06060606u) * 0x00820820u, c7 = (cur.hi.w & 0x06060606u) * 0x00820820u;
                Q0 = top_bytes4(c0, c1, c2, c3);
                Q1 = top_bytes4(c4, c5, c6, c7);
            }
            const uint32_t H0 = Sp0, H1 = Sp1;
            {
                const int d = 2 * (32 - (TL - 1));
                if (BIG) { Sp0 = __funnelshift_rc(Q0, Q1, d); Sp1 = __funnelshift_rc(Q1, 0u, d); }
                else { Sp0 = (uint32_t)((((uint64_t)Q1 << 32) | Q0) >> d); Sp1 = 0u; }
            }
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
            const uint32_t X0 = __funnelshift_r(Y0, Y1, 2 * P.out), X1 = __funnelshift_r(Y1, Y2, 2 * P.out), X2 = __funnelshift_r(Y2, Y3, 2 * P.out);
            auto xsh = [&](int k) -> uint32_t {      // X >> k for a constant k in [-2, 95]
                return k < 0 ? (X0 << (-k)) : (k < 32 ? __funnelshift_r(X0, X1, k) : (k < 64 ? __funnelshift_r(X1, X2, k - 32) : (X2 >> (k - 64))));
            };
            uint32_t cand = 0;
#pragma unroll
            for (int k = NPROBE - 1; k >= 0; k--) {
                const uint32_t word = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(pf) + (xsh(2 * ST * k - 2) & 0x1fffcu));
                cand = __funnelshift_l(__funnelshift_l(0u, word, xsh(2 * ST * k + 15)), cand, 1);
            }
            if (wm == 0) cand = 0;                                          // header bytes, the read's first 2k-1 bases, lanes past their last piece
            const uint32_t hit = __ballot_sync(kFull, cand != 0);
            if (hit) {
                if (cand) {
                    const uint32_t e = ln + __popc(hit & ((1u << lane) - 1u));
                    lq.y[0][e] = Y0; lq.y[1][e] = Y1; lq.y[2][e] = Y2; lq.y[3][e] = Y3;
                    lq.flags[e] = 0u; lq.wmask[e] = wm; lq.cand[e] = cand; lq.off[e] = off;
                }
                ln += __popc(hit);
                __syncwarp();
                if (ln >= 32) {
                    do ln = drain3<ST>(P, A, q, qn, lq, ln - 32, 32, Fq.gid, ord_base); while (ln >= 32);
                    __syncwarp();
                }
            }
        }
        start = start_n; lead = lead_n; len = len_n; pieces = pieces_n;
    }
    while (ln) {                                                           // one block hit per parked lane and pass
        const uint32_t m = ln < 32u ? ln : 32u;
        ln = drain3<ST>(P, A, q, qn, lq, ln - m, m, Fq.gid, ord_base);
    }
    __syncwarp();
    if (qn) resolve3(P, A, q, 0, qn);
}

}  // namespace kssd
