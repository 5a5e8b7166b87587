// composite.cuh -- `kssd composite` (reference get_species_abundance, command_composite.c:389-547): for every query
// sketch with abundances (-A) and every reference, the abundances of the k-mers they share, and from them
// kmer_num / mean / the 98-99 % band mean / median / max.
//
// The reference probes a per-query hash table with every reference code and qsorts each reference's list.  Here the
// reference side is the inverted index the search already has: every (query code, posting) pair becomes one 64-bit key
// (query | ref | abundance); ONE radix sort groups them by (query, ref) with the abundances ascending inside a group,
// a run-length pass delimits the groups, and one thread per group reads its order statistics straight out of the
// sorted keys.  A second small sort puts the rows in the reference's print order (most shared k-mers first, ties in
// reference order -- glibc's qsort is a stable merge sort at these sizes).
#pragma once
#include "index_dist.cuh"

namespace kssd {

struct CompRow {            // mirrors kssd_comp_row_t
    uint32_t qry, ref, kmer_num, median, max;
    float mean, pct;
    uint32_t pad;
};

constexpr int kCompQryShift = 40, kCompRefShift = 16;     // key = qry << 40 | ref << 16 | abundance

// number of (query code, posting) pairs of every query code (one component)
__global__ void comp_count_kernel(const uint32_t *__restrict__ qcodes, uint64_t n, const CodeLookup L, uint32_t *__restrict__ len)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = qcodes[i];
    uint32_t s0, s1;
    code_lookup(L, c, s0, s1);
    len[i] = s1 - s0;
}

// query number of every query code (binary search in the per-query index), then its pairs at off[i]..
__global__ void comp_emit_kernel(const uint32_t *__restrict__ qcodes, const uint16_t *__restrict__ qabund, const uint64_t *__restrict__ qindex,
                                 int n_qry, uint64_t n, const CodeLookup L, const uint32_t *__restrict__ mco,
                                 const uint64_t *__restrict__ off, uint64_t *__restrict__ keys)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = n_qry;                                   // largest q with qindex[q] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (qindex[mid] <= i) lo = mid; else hi = mid;
    }
    const uint32_t c = qcodes[i];
    const uint64_t base = ((uint64_t)lo << kCompQryShift) | qabund[i];
    uint64_t o = off[i];
    uint32_t s0, s1;
    code_lookup(L, c, s0, s1);
    for (uint32_t p = s0; p < s1; p++) keys[o++] = base | ((uint64_t)mco[p] << kCompRefShift);
}

// one thread per (query, ref) group: statistics out of the sorted abundances (low 16 bits of the keys)
__global__ void comp_rows_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ run_key, const uint32_t *__restrict__ run_len,
                                 const uint64_t *__restrict__ run_off, uint32_t n_runs, uint32_t min_kmers, CompRow *__restrict__ rows,
                                 uint64_t *__restrict__ order_key, uint32_t *__restrict__ kept)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint32_t k = run_len[r];
    const uint64_t g = run_key[r];                            // key >> 16: qry << 24 | ref
    CompRow row;
    row.qry = (uint32_t)(g >> (kCompQryShift - kCompRefShift));
    row.ref = (uint32_t)(g & ((1u << (kCompQryShift - kCompRefShift)) - 1u));
    row.kmer_num = k;
    row.pad = 0;
    if (k < min_kmers) {                                      // not printed (command_composite.c:511)
        row.median = row.max = 0; row.mean = row.pct = 0.f;
        rows[r] = row;
        order_key[r] = ~0ull;
        return;
    }
    atomicAdd(kept, 1u);
    const uint64_t *a = keys + run_off[r] - 1;                // 1-based like ref_abund[rn][1..k]
    long long sum = 0;
    for (uint32_t n = 1; n <= k; n++) sum += (long long)(a[n] & 0xffffu);
    int lastsum = 0, lastn = 0;
    for (int n = (int)((double)k * 0.98); (double)n <= (double)k * 0.99; n++) {
        lastsum += n == 0 ? (int)k : (int)(a[n] & 0xffffu);     // (k == 1: the reference reads ref_abund[rn][0], the count itself)
        lastn++;
    }
    row.median = (uint32_t)(a[k / 2] & 0xffffu);
    row.max = (uint32_t)(a[k] & 0xffffu);
    row.mean = (float)(int)sum / (float)(int)k;               // (float)sum/kmer_num with int operands (:531)
    row.pct = (float)lastsum / (float)lastn;
    rows[r] = row;
    // print order: query ascending, kmer_num descending, reference ascending
    order_key[r] = ((uint64_t)row.qry << 44) | ((uint64_t)(0xfffffu - min(k, 0xfffffu)) << 24) | row.ref;
}

__global__ void comp_gather_kernel(const CompRow *__restrict__ rows, const uint32_t *__restrict__ perm, uint32_t n, CompRow *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rows[perm[i]];
}

__global__ void comp_iota_kernel(uint32_t *p, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// run-length key of every pair: everything but the abundance
__global__ void comp_group_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ groups)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) groups[i] = keys[i] >> kCompRefShift;
}

}  // namespace kssd
