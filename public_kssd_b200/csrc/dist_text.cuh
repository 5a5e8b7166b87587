// distance.out text on the GPU (reference dist_print_nobin / output_ctrl, command_dist.c:1239-1242, 1267-1285): one line per
// statistics row, "%s\t%s\t" + the numbers of fmt_exact.cuh -- byte for byte what glibc's printf writes.  Two passes over the
// rows: line lengths -> exclusive scan -> every line written at its offset, a warp copying its 32 lines with consecutive lanes.
// A value the integer formatter hands back (fmt_exact.cuh) is counted; the host then formats that output with snprintf.
#pragma once
#include <cstdint>

#include "fmt_exact.cuh"
#include "index_dist.cuh"

namespace kssd {

constexpr int kTextThreads = 128;

// strnlen of every fixed-stride name record
__global__ void name_len_kernel(const char *__restrict__ names, size_t stride, uint32_t n, uint16_t *__restrict__ len)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char *s = names + (size_t)i * stride;
    uint32_t l = 0;
    while (l < stride && s[l]) l++;
    len[i] = (uint16_t)l;
}

__device__ __forceinline__ fmt::RowNumbers row_numbers(const StatRow &r)
{
    fmt::RowNumbers v;
    v.shared = r.shared; v.rs_u = r.rs_u; v.ref_size = r.ref_size; v.qry_size = r.qry_size;
    v.metric = r.metric; v.dist = r.dist; v.pvalue = r.pvalue; v.fdr = r.fdr;
    v.ci_m_lo = r.ci_m_lo; v.ci_m_hi = r.ci_m_hi; v.ci_d_lo = r.ci_d_lo; v.ci_d_hi = r.ci_d_hi;
    return v;
}

// pass 1: the length of every line; len[n] is left to the caller (the scan's total)
// (handed_back[0]: values the formatter declined; handed_back[1]: set when a row names a query or reference outside the lists)
__global__ void __launch_bounds__(kTextThreads) dist_text_len_kernel(const StatRow *__restrict__ rows, uint64_t n, uint32_t n_qry, uint32_t n_ref,
                                                                     const uint16_t *__restrict__ qlen, const uint16_t *__restrict__ rlen, int outfields,
                                                                     uint32_t *__restrict__ len, uint32_t *__restrict__ handed_back)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const StatRow r = rows[i];
    if (r.qry >= n_qry || r.ref >= n_ref) { handed_back[1] = 1u; len[i] = 0; return; }
    char buf[fmt::kRowNumMax];
    int nl = fmt::put_row_numbers(buf, row_numbers(r), outfields);
    if (nl < 0) { atomicAdd(handed_back, 1u); nl = 0; }
    len[i] = (uint32_t)qlen[r.qry] + (uint32_t)rlen[r.ref] + 2u + (uint32_t)nl;
}

// pass 2: the lines themselves.  A thread formats its row's numbers into shared memory; the warp then copies line after line.
__global__ void __launch_bounds__(kTextThreads) dist_text_write_kernel(const StatRow *__restrict__ rows, uint64_t n, const char *__restrict__ qn,
                                                                       const char *__restrict__ rn, size_t stride, const uint16_t *__restrict__ qlen,
                                                                       const uint16_t *__restrict__ rlen, int outfields, const uint64_t *__restrict__ off,
                                                                       char *__restrict__ text)
{
    __shared__ char slot[kTextThreads][fmt::kRowNumMax];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, w0 = threadIdx.x & ~31u;
    uint32_t q = 0, r = 0, ql = 0, rl = 0, nl = 0;
    uint64_t o = 0;
    if (i < n) {
        const StatRow row = rows[i];
        q = row.qry; r = row.ref; ql = qlen[q]; rl = rlen[r]; o = off[i];
        const int t = fmt::put_row_numbers(slot[threadIdx.x], row_numbers(row), outfields);
        nl = t < 0 ? 0u : (uint32_t)t;
    }
    __syncwarp();
    const uint64_t warp_row0 = i - lane;
    for (uint32_t l = 0; l < 32 && warp_row0 + l < n; l++) {
        const uint32_t lq = __shfl_sync(0xffffffffu, q, l), lr = __shfl_sync(0xffffffffu, r, l);
        const uint32_t a = __shfl_sync(0xffffffffu, ql, l), b = __shfl_sync(0xffffffffu, rl, l), c = __shfl_sync(0xffffffffu, nl, l);
        const uint64_t lo = __shfl_sync(0xffffffffu, o, l);
        const char *sq = qn + (size_t)lq * stride, *sr = rn + (size_t)lr * stride, *sn = slot[w0 + l];
        const uint32_t total = a + b + 2u + c;
        for (uint32_t p = lane; p < total; p += 32) {
            char ch;
            if (p < a) ch = sq[p];
            else if (p == a) ch = '\t';
            else if (p < a + 1 + b) ch = sr[p - a - 1];
            else if (p == a + 1 + b) ch = '\t';
            else ch = sn[p - a - b - 2];
            text[lo + p] = ch;
        }
    }
}

}  // namespace kssd
