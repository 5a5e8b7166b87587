// sketch_scan32.cuh -- Stage I scan, 32 bytes per lane (1 KiB per warp iteration).
//
// The clean-path loop of the FASTA scan (helpers in sketch_scan.cuh): one 256-bit load per lane (LDG.E.256 on
// sm_100), the lane's 32 bases in a 64-bit word, history from ONE neighbour lane (it holds >= 2k-1 bases), 32
// prefilter windows per lane.  The per-iteration coordination (shuffles, votes, carries, loop control, candidate
// hand-off) is paid once per KiB; an earlier 16-bytes-per-lane loop paid it twice as often and ran at 0.17 of the HBM
// roofline against 0.21 (profiles/r1_sketch_ncu_summary.md).
// Dirty iterations fall back to two 512-byte general iterations with a 16-byte lane mapping (general_iter16).
#pragma once
#include "sketch_scan.cuh"

namespace kssd {

struct Bytes32 { uint4 lo, hi; };    // lo = bytes 0..15 (older), hi = bytes 16..31

__device__ __forceinline__ Bytes32 ldg_stream256(const uint8_t *p)
{
    Bytes32 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ Bytes32 load_chunk32_guarded(const ScanArgs &A, uint64_t addr)
{
    Bytes32 r;
    r.lo = load_chunk16_guarded(A, addr);
    r.hi = load_chunk16_guarded(A, addr + 16);
    return r;
}

// (hi:lo) << s for s in [0, 64] -> four result words
__device__ __forceinline__ void shl128(uint32_t lo, uint32_t hi, uint32_t s, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    const bool big = s >= 32;
    const uint32_t a0 = big ? 0u : lo, a1 = big ? lo : hi, a2 = big ? hi : 0u;
    const uint32_t t = big ? s - 32 : s;      // 0..32
    r0 = __funnelshift_lc(0u, a0, t);
    r1 = __funnelshift_lc(a0, a1, t);
    r2 = __funnelshift_lc(a1, a2, t);
    r3 = __funnelshift_lc(a2, 0u, t);
}

__device__ __forceinline__ uint32_t low_mask(int bits)     // bits in [0, 32]
{
    return bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
}

// second-level probe + hand-over of `m` parked lanes (entries first .. first+m-1), one per lane
__device__ __forceinline__ void drain_lanes(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQueue &q, uint32_t &qn,
                                            const LaneQueue &lq, uint32_t first, uint32_t m, uint32_t gid, uint64_t ord_base)
{
    const uint32_t lane = lane_id();
    uint32_t c = 0, w0 = 0, w1 = 0, w2 = 0, w3 = 0, off = 0;
    if (lane < m) {
        const uint32_t i = first + lane;
        c = lq.cand[i]; w0 = lq.w0[i]; w1 = lq.w1[i]; w2 = lq.w2[i]; w3 = lq.w3[i]; off = lq.off[i];
    }
    push_candidates(P, A, pf, q, qn, c, 32u, w0, w1, w2, w3, off, 0u, 32u, gid, ord_base);
}

__device__ void scan_span32(const SketchParams &P, const ScanArgs &A, const uint32_t *__restrict__ pf, WarpQueue &q, LaneQueue &lq,
                            uint32_t gid, uint64_t gs, uint64_t ge, uint64_t start, uint64_t end)
{
    const uint32_t lane = lane_id();
    const int TL = P.TL;
    // stream state in registers; it is packed into a StreamState only around the out-of-line general iterations
    uint64_t cw = 0;
    uint32_t since_break = 0, after_end = 0, hdr = 0;
    uint32_t qn = 0, ln = 0;
    const uint64_t chunk0 = start & ~127ull;
    const uint64_t ord_base = chunk0 - gs;           // may wrap below zero; real occurrences add back past it
    // iterations 1 .. n_steady are "steady": wholly inside [start, min(end, ge)) -- no masking, no run-out logic
    const uint64_t lim = end < ge ? end : ge;
    const uint64_t full = (lim - chunk0) >> 10;
    const uint32_t n_steady = full > 1 ? (uint32_t)(full - 1 < 0x3fffffffull ? full - 1 : 0x3fffffffull) : 0u;
    const uint8_t *lp = A.seq + chunk0 + 32 * lane;
    Bytes32 nxt = load_chunk32_guarded(A, chunk0 + 32 * lane);
    bool at_eof = false;

    for (uint32_t it = 0;; it++) {
        Bytes32 cur = nxt;
        const bool steady = (it - 1u) < n_steady;
        const uint64_t cbase = chunk0 + ((uint64_t)it << 10);
        const uint32_t lane_off = (it << 10) + 32 * lane;
        if (it < n_steady) nxt = ldg_stream256(lp + 1024);
        else if (cbase + 1024 < ge) nxt = load_chunk32_guarded(A, cbase + 1024 + 32 * lane);
        lp += 1024;

        bool past_end = false, cut_lane = false;
        if (!steady) {
            const uint64_t laddr = cbase + 32 * lane;
            if (cbase < start || cbase + 1024 > ge) {
                mask_lane_bytes(cur.lo, clamp16((int64_t)start - (int64_t)laddr), clamp16((int64_t)ge - (int64_t)laddr));
                mask_lane_bytes(cur.hi, clamp16((int64_t)start - (int64_t)(laddr + 16)), clamp16((int64_t)ge - (int64_t)(laddr + 16)));
            }
            past_end = cbase + 1024 > end;
            cut_lane = laddr < start || laddr + 32 > ge;
        }

        uint32_t dacc = 0, t0, t1, t2, t3, t4, t5, t6, t7, m0, m1, m2, m3, m4, m5, m6, m7;
        classify4(cur.lo.x, dacc, t0, m0);
        classify4(cur.lo.y, dacc, t1, m1);
        classify4(cur.lo.z, dacc, t2, m2);
        classify4(cur.lo.w, dacc, t3, m3);
        classify4(cur.hi.x, dacc, t4, m4);
        classify4(cur.hi.y, dacc, t5, m5);
        classify4(cur.hi.z, dacc, t6, m6);
        classify4(cur.hi.w, dacc, t7, m7);
        uint32_t rskA = (rev_flags8(t0, t1) << 8) | rev_flags8(t2, t3);     // older 16 bytes
        uint32_t rskB = (rev_flags8(t4, t5) << 8) | rev_flags8(t6, t7);     // newer 16 bytes
        const uint32_t nA = 16 - __popc(rskA), nB = 16 - __popc(rskB);
        const uint32_t n = nA + nB;
        const bool lane_ok = dacc == 0 && (n >= (uint32_t)(TL - 1) || cut_lane);
        const bool clean = __all_sync(kFull, lane_ok) && !hdr;

        if (clean) {
            uint32_t PA = prmt(prmt(m3, m2, 0x0073u), prmt(m1, m0, 0x0073u), 0x5410u);
            uint32_t PB = prmt(prmt(m7, m6, 0x0073u), prmt(m5, m4, 0x0073u), 0x5410u);
            for (;;) {      // squeeze the line ends out of both halves; first round is branch-free
                const uint32_t ia = rskA & (0u - rskA), ib = rskB & (0u - rskB);
                const uint32_t la = ia * ia - 1u, lb = ib * ib - 1u;
                PA = ((PA >> 2) & ~la) | (PA & la);
                PB = ((PB >> 2) & ~lb) | (PB & lb);
                rskA = (rskA ^ ia) >> 1;
                rskB = (rskB ^ ib) >> 1;
                if (!__any_sync(kFull, (rskA | rskB) != 0)) break;
            }
            // the lane's bases, newest in the low bits: (PA << 2 nB) | PB
            uint32_t P0, P1, junk0, junk1;
            shl128(PA, 0u, 2 * nB, P0, P1, junk0, junk1);
            P0 |= PB;
            // history = the previous lane's bases (it holds >= 2k-1 of them), the warp carry for lane 0
            uint32_t H0 = __shfl_up_sync(kFull, P0, 1), H1 = __shfl_up_sync(kFull, P1, 1);
            if (lane == 0) { H0 = (uint32_t)cw; H1 = (uint32_t)(cw >> 32); }
            uint32_t W0, W1, W2, W3;
            shl128(H0, H1, 2 * n, W0, W1, W2, W3);
            W0 |= P0;
            W1 |= P1;
            // prefilter on the central 2s-mer of the k-mer ending at each own base (d = distance from the newest)
            const uint32_t X0 = __funnelshift_r(W0, W1, 2 * P.out);
            const uint32_t X1 = __funnelshift_r(W1, W2, 2 * P.out);
            const uint32_t X2 = __funnelshift_r(W2, W3, 2 * P.out);
            uint32_t cand = 0;
            // t(e) = X >> 2e; window d reads the word at t(d-1) & 0x1fffc and shifts it by t(d+8) (layout in kssd_device.cuh)
            auto tsh = [&](int e) -> uint32_t {
                return e < 0 ? (X0 << 2) : (e < 16 ? __funnelshift_r(X0, X1, 2 * e) : (e < 32 ? __funnelshift_r(X1, X2, 2 * e - 32) : (X2 >> (2 * e - 64))));
            };
#pragma unroll
            for (int d = 31; d >= 0; d--) {
                const uint32_t word = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(pf) + (tsh(d - 1) & (kPfWordMask << 2)));
                cand = __funnelshift_l(__funnelshift_l(0u, word, tsh(d + 8)), cand, 1);      // cand = cand << 1 | flag
            }
            cand &= low_mask((int)n);

            const uint32_t N = __reduce_add_sync(kFull, n);
            if (since_break < (uint32_t)(TL - 1) || past_end) {
                // start of a span / run-out past its end: filter by position inside the iteration
                uint32_t incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                const int o_l = (int)(incl - n);                     // valid bases before this lane
                // ok1: since_break + t_local + 1 >= TL  with t_local = o_l + n-1-d
                const int need = TL - 1 - (int)since_break - o_l; // n-1-d >= need
                if (need > 0) cand &= low_mask(max((int)n - need, 0));
                if (past_end) {
                    uint32_t E;                                      // valid bases of this iteration before `end`
                    const int64_t rel = (int64_t)end - (int64_t)cbase;
                    if (rel <= 0) E = 0;
                    else {
                        const int le = (int)(rel >> 5);
                        const int be = (int)(rel & 31);              // bytes [0, be) of lane `le` lie before `end`
                        // the squeeze consumed rskA/rskB: recount the skipped bytes from nA / nB and the fresh flags
                        const uint32_t fA = (rev_flags8(t0, t1) << 8) | rev_flags8(t2, t3);
                        const uint32_t fB = (rev_flags8(t4, t5) << 8) | rev_flags8(t6, t7);
                        uint32_t vb;                                 // valid bytes among the first `be`
                        if (be <= 16) vb = (uint32_t)be - (uint32_t)__popc(fA >> (16 - be));
                        else vb = nA + (uint32_t)(be - 16) - (uint32_t)__popc(fB >> (32 - be));
                        E = __shfl_sync(kFull, (uint32_t)o_l + vb, le);
                    }
                    // ok2: after_end + (t_local - E + 1) <= TL-1   for t_local >= E
                    const int lim2 = TL - 2 - (int)after_end + (int)E - o_l;   // n-1-d <= lim2
                    const int drop = (int)n - 1 - lim2;                           // d >= drop
                    if (drop > 0) cand &= ~low_mask(min(drop, 32));
                    after_end += N - E;
                }
            }
            since_break = min(since_break + N, kRunCap);
            cw = ((uint64_t)__shfl_sync(kFull, P1, 31) << 32) | __shfl_sync(kFull, P0, 31);
            const uint32_t hit = __ballot_sync(kFull, cand != 0);
            if (hit) {
                if (cand) {
                    const uint32_t i = ln + __popc(hit & ((1u << lane) - 1u));
                    lq.w0[i] = W0; lq.w1[i] = W1; lq.w2[i] = W2; lq.w3[i] = W3; lq.cand[i] = cand; lq.off[i] = lane_off;
                }
                ln += __popc(hit);
                __syncwarp();
                if (ln >= 32) {
                    drain_lanes(P, A, pf, q, qn, lq, ln - 32, 32, gid, ord_base);
                    ln -= 32;
                    __syncwarp();
                }
            }
        } else {
            // two general 512-byte iterations with the 16-byte lane mapping (reloaded: L1/L2 hits)
            StreamState st = {cw, since_break, after_end, hdr};
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint64_t sbase = cbase + 512ull * h;
                if (sbase >= ge) break;
                const uint64_t laddr = sbase + 16 * lane;
                uint4 c16 = load_chunk16_guarded(A, laddr);
                if (sbase < start || sbase + 512 > ge)
                    mask_lane_bytes(c16, clamp16((int64_t)start - (int64_t)laddr), clamp16((int64_t)ge - (int64_t)laddr));
                uint32_t dd = 0, u0, u1, u2, u3, k0, k1, k2, k3;
                classify4(c16.x, dd, u0, k0);
                classify4(c16.y, dd, u1, k1);
                classify4(c16.z, dd, u2, k2);
                classify4(c16.w, dd, u3, k3);
                const uint32_t codes = prmt(prmt(k3, k2, 0x0073u), prmt(k1, k0, 0x0073u), 0x5410u);
                general_iter16(P, A, pf, q, qn, st, c16, codes, sbase, end, sbase + 512 > end, (it << 10) + 512u * h + 16 * lane, gid, ord_base);
            }
            cw = st.cw; since_break = st.since_break; after_end = st.after_end; hdr = st.hdr;
        }

        if (!steady) {
            if (cbase + 1024 >= ge) { at_eof = true; break; }   // genome exhausted
            if (cbase + 1024 >= end) {                          // run-out: stop when no owned k-mer can still end
                if (after_end >= (uint32_t)(TL - 1) || since_break <= after_end) break;
            }
        }
    }
    if (ln) { drain_lanes(P, A, pf, q, qn, lq, 0, ln, gid, ord_base); __syncwarp(); }
    if (qn) { resolve_candidates(P, A, q, 0, qn, gid, ord_base); __syncwarp(); }
    if (hdr && at_eof && lane == 0) atomicOr(&A.gstatus[gid], 1);   // the text ended inside a '>' line
}

__global__ void __launch_bounds__(kScanThreads, 1) sketch_fasta32_kernel(const SketchParams P, const ScanArgs A)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *pf = reinterpret_cast<uint32_t *>(smem_raw);
    WarpQueue *queues = reinterpret_cast<WarpQueue *>(smem_raw + (kPfWords + kPf2Words) * 4);
    LaneQueue *lqueues = reinterpret_cast<LaneQueue *>(queues + kScanWarps);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.prefilter);
        uint4 *dst = reinterpret_cast<uint4 *>(pf);
        for (uint32_t i = threadIdx.x; i < (kPfWords + kPf2Words) / 4; i += blockDim.x) dst[i] = __ldg(&src[i]);
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
        uint64_t start, end;
        if (!span_extent(A, si, gid, gs, ge, start, end)) continue;
        scan_span32(P, A, pf, q, lqueues[threadIdx.x >> 5], gid, gs, ge, start, end);
    }
}

}  // namespace kssd
