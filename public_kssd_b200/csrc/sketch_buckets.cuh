// sketch_buckets.cuh -- Stage I post-pass without a global sort and without a host round trip in the middle.
//
// The reference inserts every sampled k-mer of a genome into that genome's own hash table (iseq2comem.c:255-268) and the
// writers emit one id list per (component, genome) (:525-551).  The list-mode post-pass (kssd_b200.cu) rebuilds that
// grouping with a 64-bit radix sort of all occurrences of the batch, which needs their number on the host first.  Here
// the scan's resolver drops each occurrence into its (component, genome) BUCKET directly -- bucket capacities follow
// from the genome lengths alone, so the host lays them out before the launch -- and one CTA per bucket sorts its few
// thousand (id, offset) keys in shared memory, collapses runs (multiplicity, first occurrence), applies the mode's keep
// rule and leaves the bucket compacted.  A scan of the kept counts and one move kernel give the final layout.  A bucket
// that overflows (an extremely repetitive genome) sets a flag and the whole batch is redone in list mode.
#pragma once
#include "kssd_device.cuh"

namespace kssd {

constexpr int kBucketThreads = 256;
constexpr uint32_t kBucketMaxCap = 8192;                 // keys one CTA sorts: 64 KiB of shared memory + 16 KiB of run starts

// bkeys[boff[b] .. +min(bcnt[b], cap)) : id << 36 | byte offset.  Out (compacted at the bucket's start): ids, multiplicity
// (saturated u16), first-occurrence offset; kept[b]; distinct keys per genome (before the keep rule).
__global__ void __launch_bounds__(kBucketThreads) bucket_finish_kernel(const uint64_t *__restrict__ bkeys, const uint32_t *__restrict__ boff,
                                                                        const uint32_t *__restrict__ bcnt, const uint32_t *__restrict__ overflow,
                                                                        uint32_t n_genomes, int mode, int M, uint32_t *__restrict__ t_ids,
                                                                        uint16_t *__restrict__ t_ab, uint64_t *__restrict__ t_ord,
                                                                        uint32_t *__restrict__ kept, uint32_t *__restrict__ distinct_pg, uint32_t *__restrict__ n_occ_total)
{
    extern __shared__ __align__(16) uint8_t bucket_sm[];
    if (*overflow) return;                                // the batch is redone in list mode anyway
    const uint32_t b = blockIdx.x, lo = boff[b], cap = boff[b + 1] - lo;
    const uint32_t n = min(bcnt[b], cap);
    if (n == 0) { if (threadIdx.x == 0) kept[b] = 0; return; }
    uint32_t P2 = 32;
    while (P2 < n) P2 <<= 1;
    uint64_t *key = reinterpret_cast<uint64_t *>(bucket_sm);
    uint16_t *hpos = reinterpret_cast<uint16_t *>(key + P2);                 // run starts, then (reused) kept-run numbers
    __shared__ uint32_t wsum[kBucketThreads / 32], tot_s;
    for (uint32_t i = threadIdx.x; i < P2; i += kBucketThreads) key[i] = i < n ? bkeys[lo + i] : ~0ull;
    __syncthreads();
    // bitonic sort, ascending: ids ascend, and inside an id the offsets do -- a run's first key is its first occurrence
    for (uint32_t k = 2; k <= P2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < P2 / 2; t += kBucketThreads) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), x = i | j;       // the pair (i, i ^ j) with bit j clear in i
                const uint64_t a = key[i], c = key[x];
                if ((a > c) == ((i & k) == 0)) { key[i] = c; key[x] = a; }
            }
            __syncthreads();
        }
    // block-wide exclusive scan of one value per thread
    auto block_scan = [&](uint32_t v, uint32_t &total) -> uint32_t {
        const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        __syncthreads();
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        uint32_t off = incl - v, tot = 0;
#pragma unroll
        for (uint32_t w = 0; w < kBucketThreads / 32; w++) { const uint32_t s = wsum[w]; off += w < wid ? s : 0u; tot += s; }
        total = tot;
        return off;
    };
    // run starts: every thread owns a contiguous stretch of the sorted keys
    const uint32_t per = P2 / kBucketThreads ? P2 / kBucketThreads : 1;
    const uint32_t i0 = min(threadIdx.x * per, n), i1 = min(i0 + per, n);
    uint32_t hc = 0;
    for (uint32_t i = i0; i < i1; i++) hc += (i == 0 || (key[i] >> 36) != (key[i - 1] >> 36)) ? 1u : 0u;
    uint32_t nruns;
    uint32_t hb = block_scan(hc, nruns);
    for (uint32_t i = i0; i < i1; i++)
        if (i == 0 || (key[i] >> 36) != (key[i - 1] >> 36)) hpos[hb++] = (uint16_t)i;
    __syncthreads();
    // keep rule per run (contiguous stretches of runs per thread), compaction by a second scan
    const uint32_t rper = (nruns + kBucketThreads - 1) / kBucketThreads;
    const uint32_t r0 = min(threadIdx.x * rper, nruns), r1 = min(r0 + rper, nruns);
    auto run_count = [&](uint32_t r) -> uint32_t { return (r + 1 < nruns ? (uint32_t)hpos[r + 1] : n) - (uint32_t)hpos[r]; };
    auto keep_rule = [&](uint32_t cnt) -> bool {
        if (mode == KSSD_MODE_FASTA_UNIQ) return cnt == 1;               // iseq2comem.c:694-695 + :540
        if (mode == KSSD_MODE_FASTQ) return cnt >= (uint32_t)M;          // iseq2comem.c:336-346 + :514
        return true;
    };
    uint32_t kc = 0;
    for (uint32_t r = r0; r < r1; r++) kc += keep_rule(run_count(r)) ? 1u : 0u;
    uint32_t nkept;
    uint32_t kb = block_scan(kc, nkept);
    for (uint32_t r = r0; r < r1; r++) {
        const uint32_t cnt = run_count(r);
        if (!keep_rule(cnt)) continue;
        const uint64_t k0 = key[hpos[r]];
        t_ids[lo + kb] = (uint32_t)(k0 >> 36);
        t_ab[lo + kb] = (uint16_t)min(cnt, 65535u);                       // iseq2comem.c:602-604
        t_ord[lo + kb] = k0 & 0xfffffffffull;
        kb++;
    }
    if (threadIdx.x == 0) {
        kept[b] = nkept;
        atomicAdd(n_occ_total, n);
        atomicAdd(&distinct_pg[b % n_genomes], nruns);                    // every distinct key took a slot (keycount, :262 / :689)
    }
}

// kept entries of bucket b: from the bucket's start in the temporaries to their final place foff[b] ..
__global__ void bucket_move_kernel(const uint32_t *__restrict__ boff, const uint32_t *__restrict__ foff, const uint32_t *__restrict__ t_ids,
                                   const uint16_t *__restrict__ t_ab, const uint64_t *__restrict__ t_ord, uint32_t *__restrict__ ids,
                                   uint16_t *__restrict__ abund, uint64_t *__restrict__ ord)
{
    const uint32_t b = blockIdx.x, src = boff[b], dst = foff[b], n = foff[b + 1] - dst;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        ids[dst + i] = t_ids[src + i];
        abund[dst + i] = t_ab[src + i];
        ord[dst + i] = t_ord[src + i];
    }
}

}  // namespace kssd
