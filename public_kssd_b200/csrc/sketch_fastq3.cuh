// sketch_fastq3.cuh -- FASTQ read walk, warp formulation: the block prefilter and the lazy codes of sketch_scan3.cuh applied to reads.
//
// Replaces the per-read loops of fastq2co / mt_shortreads2koc (reference iseq2comem.c:277-356, :554-615) whenever the quality bytes
// cannot change the outcome: -A (quality ignored), or -Q <= 0 in a file whose single-pass line index saw no byte >= 0x80 (a quality
// byte fails -Q <= 0 only as a negative signed char).  -Q > 0 and files beyond the single-pass index keep the thread-per-read walk
// of sketch_fastq.cuh, which was as ALU bound as round 1's FASTA scan (exact classification, one probe per base, ~2000 instructions
// per read).
//
// A warp takes 32 records at a time.  Lane l frames record l from the line index (positions of its five line ends; the record rules
// and the fgets-length check are the ones of sketch_fastq_kernel), and the sequence line is cut into PIECES of 32 text-aligned bytes
// (the first one starts at s0 & ~31: `lead` bytes of it belong to the header line).  A prefix sum over the 32 records numbers the
// pieces; the warp then goes through them 32 at a time, one piece per lane -- one aligned 32-byte load per lane, the piece's read
// found by a five-step search of the prefix in shared memory.  From there on a piece is a lane of the FASTA scan's steady loop
// without skipped bytes: codes from bits 1-2 (no exact classification: an N is a fake base until the end), the previous lane's last
// 2k-1 bases as history (the previous piece of the same read -- the first piece needs none), 12 probes of the block bitmap per 32
// bases, lanes with a block hit parked and drained 32 at a time through the block table, candidates resolved exactly.  Which k-mers
// of a piece may count is a window mask from the read's extent alone: first base at or after s0, last base before the line end.
// The final check is against the text: the 2k bytes ending at the position are all letters (strict: a read knows no line ends).
#pragma once
#include "sketch_fastq.cuh"
#include "sketch_scan3.cuh"

namespace kssd {

struct Fq3Batch { uint32_t pre[33], start[32], len[32], lead[32]; };      // pieces before record i of the batch | aligned start | bases | header bytes in piece 0
constexpr size_t kFq3SmemBytes = (size_t)kPf3Words * 4 + (size_t)kScanWarps * (sizeof(WarpQ3) + sizeof(LaneQ3) + sizeof(Fq3Batch));

template <int ST, bool BIG>
__global__ void __launch_bounds__(kScanThreads, 1) sketch_fastq3_kernel(const __grid_constant__ SketchParams P, const __grid_constant__ ScanArgs A,
                                                                           const __grid_constant__ FastqArgs Fq, const uint32_t *__restrict__ pf_global)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *pf = reinterpret_cast<uint32_t *>(smem_raw);
    WarpQ3 *queues = reinterpret_cast<WarpQ3 *>(smem_raw + kPf3Words * 4);
    LaneQ3 *lqueues = reinterpret_cast<LaneQ3 *>(queues + kScanWarps);
    Fq3Batch *batches = reinterpret_cast<Fq3Batch *>(lqueues + kScanWarps);
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
    Fq3Batch &bt = batches[wid];
    const int TL = P.TL;
    const uint32_t hsh = (2u * (uint32_t)(TL - 1)) & 31u;
    const uint64_t n_nl = (uint64_t)Fq.idx->n_nl;
    const uint64_t n_lines = n_nl + (A.seq[Fq.ge - 1] != '\n' ? 1u : 0u);
    const uint64_t n_records = (n_lines + 3) / 4, n_batches = (n_records + 31) / 32;
    const uint64_t base = Fq.pos_base & ~31ull;                             // piece starts are kept as 32-bit offsets from here
    const uint64_t ord_base = base - Fq.gs;                                 // (may wrap below zero: occurrences add back past it)
    uint32_t qn = 0, ln = 0, cw0 = 0, cw1 = 0;

    // the five line ends that frame record r (lines 4r-1 .. 4r+3), as offsets; asked for one batch ahead of their use
    auto fetch_nl = [&](uint64_t bb, uint32_t (&v)[5]) {
        const uint64_t r = 32 * bb + lane;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const uint64_t idx = 4 * r + k;                                  // line 4r - 1 + k
            v[k] = (bb < n_batches && idx >= 1 && idx - 1 < n_nl) ? __ldg(&Fq.nlpos32[idx - 1]) : 0u;
        }
    };
    const uint64_t b_step = (uint64_t)gridDim.x * kScanWarps;
    uint64_t b = (uint64_t)blockIdx.x * kScanWarps + wid;
    uint32_t nl5[5];
    fetch_nl(b, nl5);
    for (; b < n_batches; b += b_step) {
        // ---- frame the batch's records (one per lane): sequence line extent, record rules, fgets-length check ----
        {
            const uint64_t r = 32 * b + lane;
            uint64_t s0 = 0, len = 0;
            const uint64_t l_seq = 4 * r + 1, l_q = 4 * r + 3;
            auto NL5 = [&](int k) -> uint64_t { return Fq.pos_base + nl5[k]; };          // line 4r - 1 + k (valid where the rules below look)
            if (r < n_records && l_seq < n_lines) {
                const bool process = Fq.abund ? (4 * r + 4 <= n_lines) : (r == 0 || 4 * r + 4 <= n_nl);      // iseq2comem.c:567 / :300-307
                if (process) {
                    s0 = NL5(1) + 1;                                          // line 4r (the header) is terminated: l_seq < n_lines
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
                        atomicOr(&A.gstatus[Fq.gid], 2);                      // the reference would mis-frame every later record
                        len = 0;
                    }
                    if (len < (uint64_t)TL) len = 0;                          // no k-mer fits
                }
            }
            fetch_nl(b + b_step, nl5);                                        // the next batch's line ends travel under this batch's pieces
            const uint32_t lead = (uint32_t)(s0 & 31), pieces = len ? (uint32_t)((lead + len + 31) >> 5) : 0u;
            uint32_t incl = pieces;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, incl, o);
                if (lane >= (uint32_t)o) incl += t;
            }
            __syncwarp();                                                  // (the previous batch's pieces are done with the record table)
            bt.pre[lane + 1] = incl;
            if (lane == 0) bt.pre[0] = 0;
            bt.start[lane] = (uint32_t)((s0 & ~31ull) - base);
            bt.len[lane] = (uint32_t)len;
            bt.lead[lane] = lead;
            __syncwarp();
        }
        const uint32_t total = bt.pre[32];
        // the lane's piece of the round that starts at slot s_first: record by a search of the prefix (pre[i] <= s < pre[i + 1]), piece
        // number, window mask, and the text on its way
        auto stage = [&](uint32_t s_first, Bytes32 &text, uint32_t &wm_o, uint32_t &off_o) {
            const uint32_t s = s_first + lane;
            const bool active = s < total;
            uint32_t i = 0;
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1)
                if (bt.pre[i + step] <= s) i += step;
            if (!active) i = 0;
            const uint32_t p = s - bt.pre[i], off = bt.start[i] + 32u * p;          // byte offset of the piece from `base`
            const int shift = (int)(32u * p) - (int)bt.lead[i];                     // read position of the piece's byte 0
            const uint64_t addr = base + off;
            if (active && addr + 32 <= A.seq_bytes) text = ldg_stream256(A.seq + addr);
            else if (active) text = load_chunk32_guarded(A, addr);
            else { text.lo = make_uint4(0, 0, 0, 0); text.hi = text.lo; }
            // own byte j ends a k-mer of the read iff the k-mer starts at or after the read's first base and j lies before the line end
            uint32_t wm = 0;
            if (active) {
                const int lo = TL - 1 - shift, hi = (int)bt.len[i] - shift;
                wm = low_mask(hi < 0 ? 0 : (hi > 32 ? 32 : hi)) & ~low_mask(lo < 0 ? 0 : (lo > 32 ? 32 : lo));
            }
            wm_o = wm;
            off_o = off;
        };
        Bytes32 nxt;
        uint32_t wm_n = 0, off_n = 0;
        if (total) stage(0, nxt, wm_n, off_n);
        for (uint32_t s_first = 0; s_first < total; s_first += 32) {
            const Bytes32 cur = nxt;
            const uint32_t wm = wm_n, off = off_n;
            if (s_first + 32 < total) stage(s_first + 32, nxt, wm_n, off_n);          // the next round's text is requested before this one is used
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
            uint32_t S0, S1;
            {
                const int d = 2 * (32 - (TL - 1));
                if (BIG) { S0 = __funnelshift_rc(Q0, Q1, d); S1 = __funnelshift_rc(Q1, 0u, d); }
                else { S0 = (uint32_t)((((uint64_t)Q1 << 32) | Q0) >> d); S1 = 0u; }
            }
            uint32_t H0 = __shfl_up_sync(kFull, S0, 1), H1 = BIG ? __shfl_up_sync(kFull, S1, 1) : 0u;
            if (lane == 0) { H0 = cw0; H1 = cw1; }
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
            if (wm == 0) cand = 0;                                          // header bytes, the read's first 2k-1 bases, lanes past the last piece
            cw0 = __shfl_sync(kFull, S0, 31);
            if (BIG) cw1 = __shfl_sync(kFull, S1, 31);
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
    }
    while (ln) {                                                           // one block hit per parked lane and pass
        const uint32_t m = ln < 32u ? ln : 32u;
        ln = drain3<ST>(P, A, q, qn, lq, ln - m, m, Fq.gid, ord_base);
    }
    __syncwarp();
    if (qn) resolve3(P, A, q, 0, qn);
}

}  // namespace kssd
