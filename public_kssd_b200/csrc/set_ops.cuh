// set_ops.cuh -- `kssd set` on sketches (reference command_set.c): union / uniq union of all genomes of a component and
// the intersect / subtract filters against a pan sketch.  The reference walks a 2^28-bit dictionary serially; here the
// same dictionary lives in 32 MiB of device memory, is filled with atomic ORs, and the ordered outputs come from
// two-pass (count, scan, fill) compactions.
#pragma once
#include "kssd_device.cuh"

namespace kssd {

constexpr int kSetThreads = 256;
constexpr int kSetWordsPerBlock = kSetThreads * 4;

// dictionary bits of the codes; `twice` (optional) gets the bit of every code seen more than once
__global__ void set_mark_kernel(const uint32_t *__restrict__ codes, uint64_t n, uint32_t *__restrict__ seen, uint32_t *__restrict__ twice)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = codes[i], bit = 1u << (c & 31);
    const uint32_t old = atomicOr(&seen[c >> 5], bit);
    if (twice && (old & bit)) atomicOr(&twice[c >> 5], bit);
}

__device__ __forceinline__ uint32_t set_word(const uint32_t *seen, const uint32_t *twice, uint64_t w)
{
    return twice ? (seen[w] & ~twice[w]) : seen[w];
}

// members of the dictionary in ascending order: per block of words count, (scan outside), then write
__global__ void __launch_bounds__(kSetThreads) set_count_kernel(const uint32_t *__restrict__ seen, const uint32_t *__restrict__ twice, uint64_t n_words,
                                                                 uint32_t *__restrict__ block_counts)
{
    const uint64_t w0 = (uint64_t)blockIdx.x * kSetWordsPerBlock + 4ull * threadIdx.x;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (w0 + j < n_words) c += __popc(set_word(seen, twice, w0 + j));
    c = __reduce_add_sync(kFull, c);
    __shared__ uint32_t red[kSetThreads / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int i = 0; i < kSetThreads / 32; i++) s += red[i];
        block_counts[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kSetThreads) set_fill_kernel(const uint32_t *__restrict__ seen, const uint32_t *__restrict__ twice, uint64_t n_words,
                                                                const uint32_t *__restrict__ block_offsets, uint32_t *__restrict__ out)
{
    const uint64_t w0 = (uint64_t)blockIdx.x * kSetWordsPerBlock + 4ull * threadIdx.x;
    uint32_t m[4], c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        m[j] = w0 + j < n_words ? set_word(seen, twice, w0 + j) : 0u;
        c += __popc(m[j]);
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    __shared__ uint32_t wsum[kSetThreads / 32];
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    uint32_t o = block_offsets[blockIdx.x] + incl - c;
    for (uint32_t w = 0; w < wid; w++) o += wsum[w];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t v = m[j];
        while (v) {
            const int b = __ffs(v) - 1;
            v &= v - 1;
            out[o++] = (uint32_t)((w0 + j) * 32 + b);
        }
    }
}

// intersect / subtract: flag = (code in dictionary) == intersect
__global__ void set_flag_kernel(const uint32_t *__restrict__ codes, uint64_t n, const uint32_t *__restrict__ dict, int intersect,
                                uint32_t *__restrict__ flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = codes[i];
    flags[i] = (uint32_t)(((dict[c >> 5] >> (c & 31)) & 1u) == (uint32_t)intersect);
}

// order-preserving compaction + the rebuilt per-genome index (the exclusive scan read at the genome boundaries)
__global__ void set_scatter_kernel(const uint32_t *__restrict__ codes, uint64_t n, const uint32_t *__restrict__ flags,
                                   const uint32_t *__restrict__ pos, uint32_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) out[pos[i]] = codes[i];
}

__global__ void set_index_kernel(const uint64_t *__restrict__ index, int n_genomes, uint64_t n, const uint32_t *__restrict__ flags,
                                 const uint32_t *__restrict__ pos, uint64_t *__restrict__ out_index)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_genomes) return;
    const uint64_t i = index[g];
    out_index[g] = i < n ? pos[i] : (n ? (uint64_t)pos[n - 1] + flags[n - 1] : 0ull);
}


// ---- kssd set -g (grouping_genomes, reference command_set.c:698-790): the sketches of a group's member genomes, walked in the
// order the reference inserts them into the group's hash table; what the table ends up holding is the DISTINCT codes, and the
// order they were first met is all its slot layout depends on.  Here: member m's codes go to positions m_dst[m] .. of one
// group-major sequence as (group << 32 | code) keys with their position as the value; a stable radix sort groups equal
// (group, code) pairs with the earliest position first, the run heads are kept and sorted back by position.
__global__ void group_gather_kernel(const uint32_t *__restrict__ combco, const uint64_t *__restrict__ m_src, const uint64_t *__restrict__ m_dst,
                                    const uint32_t *__restrict__ m_group, uint32_t n_members, unsigned long long *__restrict__ keys, uint32_t *__restrict__ pos)
{
    const uint32_t m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (m >= n_members) return;
    const uint64_t s = m_src[m], d = m_dst[m], n = m_dst[m + 1] - d;
    const unsigned long long g = (unsigned long long)m_group[m] << 32;
    for (uint64_t i = lane; i < n; i += 32) { keys[d + i] = g | combco[s + i]; pos[d + i] = (uint32_t)(d + i); }
}

__global__ void group_heads_kernel(const unsigned long long *__restrict__ keys, uint64_t n, uint32_t *__restrict__ flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// heads only: (first position, code), compacted
__global__ void group_compact_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ pos, const uint32_t *__restrict__ flags,
                                     const uint32_t *__restrict__ excl, uint64_t n, uint32_t *__restrict__ h_pos, uint32_t *__restrict__ h_code)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) { h_pos[excl[i]] = pos[i]; h_code[excl[i]] = (uint32_t)keys[i]; }
}

}  // namespace kssd
