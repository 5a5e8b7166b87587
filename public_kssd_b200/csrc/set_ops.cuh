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

}  // namespace kssd
