// kssd_device.cuh -- shared device-side definitions for libkssd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libkssd_b200 is written for sm_100a (B200) only"
#endif

namespace kssd {

constexpr int kWarp = 32;
constexpr uint32_t kFull = 0xffffffffu;

// Prefilter: 2^20-bit membership bitmap of (S u RC(S)) projected on the low 20 bits of the inner
// 2s-mer.  128 KiB -- it lives in shared memory of the one persistent CTA per SM.
constexpr int kPfBits = 20;
constexpr uint32_t kPfWords = 1u << (kPfBits - 5);
// Bit layout of the prefilter for a window value v (the central 2s-mer in the low bits of v):
//   word = v & 0x7fff (bits 0..14),  bit = 31 - ((v >> 16) & 31) (bits 16..20, reversed so that a left funnel
//   shift by v >> 16 brings the flag to bit 31).  Bit 15 is skipped on purpose: with t[e] = X >> 2e the word offset
//   of window d is t[d-1] & 0x1fffc and its shift amount is t[d+8] -- one funnel shift per window serves both.
constexpr uint32_t kPfWordMask = 0x7fffu;
constexpr int kPfBitShift = 16;
// Second-level filter (same set, independent hash of the whole window): 2^17 bits, consulted only for the ~1/128
// windows that pass the first level, so that only ~1/2000 reach the exact path.
constexpr uint32_t kPf2Words = 1u << 12;

__host__ __device__ __forceinline__ uint32_t pf2_index(uint32_t inner) { return (inner * 0x9E3779B1u) >> 15; }

constexpr uint32_t kHtEmpty = 0xffffffffu;

// Everything the scan kernels need from (k, subk, drlevel, .shuf); filled by kssd_ctx_create.
// Field meanings follow seq2co_global_var_initial (reference iseq2comem.c:54-77).
struct SketchParams {
    int k, s, L;
    int TL;              // 2k bases per k-mer
    int out;             // k - s outer bases on each side
    int hist_min_n;      // ceil((TL-1)/2): min valid bases per lane for the clean path
    uint64_t tupmask;    // low 4k bits
    uint64_t undomask;   // left outer bases of the canonical k-mer
    uint64_t outmask;    // right outer bases (low 2*out bits)
    uint32_t innermask;  // low 4s bits
    uint32_t dim_end;
    int comp_code_bits;
    uint32_t comp_mask;  // component_num - 1
    uint32_t ht_mask;    // sampled-set hash table size - 1
    const uint32_t *prefilter;  // kPfWords words (global copy), then kPf2Words words of the second level
    const uint2 *ht;            // {inner, pf}
    const unsigned long long *gtab;   // lazy scan, 2s >= 12: per 10-base block, members among the three windows around it (else null)
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// streaming 16-byte load: read once, do not pollute L1
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// reverse complement of an nb-base 2-bit string held in the low 2*nb bits (newest base lowest)
__host__ __device__ __forceinline__ uint64_t revcomp2(uint64_t x, int nb)
{
    // reverse the order of 2-bit groups of the 64-bit word, complement, then right-align
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
    x = ((x >> 8) & 0x00ff00ff00ff00ffull) | ((x & 0x00ff00ff00ff00ffull) << 8);
    x = ((x >> 16) & 0x0000ffff0000ffffull) | ((x & 0x0000ffff0000ffffull) << 16);
    x = (x >> 32) | (x << 32);
    x = ~x;
    return nb == 32 ? x : (x >> (64 - 2 * nb));
}

}  // namespace kssd
