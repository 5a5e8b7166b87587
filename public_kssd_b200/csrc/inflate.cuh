// gzip / DEFLATE (RFC 1952 / 1951) decoding, one compressed file per thread -- what the reference gets from popen("zcat -fc")
// (iseq2comem.c:187-200, :283-290).  A batch of Stage I is hundreds of independent .gz genomes: decoding them on the GPU lets the
// compressed bytes cross PCIe (3-4x fewer) and takes inflate off the host cores.  A single stream is serial (every symbol's
// position depends on the one before), so the parallelism is over files; `stage1_files.cuh` picks this path when a batch holds
// enough of them and leaves large single files to zlib on the host.
//
// Decoder: 64-bit bit buffer refilled by bytes; literal/length codes through a 10-bit lookup table (codes of DNA text are 2-9 bits),
// distance codes through an 8-bit one, longer codes by the canonical count/symbol walk; tables live in memory the caller provides
// (shared memory on the GPU).  Members are decoded one after another (multi-member files, `cat a.gz b.gz`); every member's ISIZE and
// CRC-32 are checked.  The same code compiles for the host, where the CPU test suite runs it against zlib.
#pragma once
#include <cstdint>

namespace kssd {
namespace gz {

#ifdef __CUDACC__
#define KGZ __host__ __device__ inline
#else
#define KGZ inline
#endif

enum : int { kOk = 0, kBadHeader = -1, kBadData = -2, kOutputFull = -3, kTruncated = -4, kBadCrc = -5, kBadSize = -6 };

constexpr int kLRoot = 10, kDRoot = 8;

struct Tables {
    uint16_t lfast[1 << kLRoot];      // (symbol << 4) | length, 0 = longer than the root
    uint16_t dfast[1 << kDRoot];
    uint16_t lcount[16], dcount[16];  // canonical walk: codes per length, symbols in code order
    uint16_t lsym[288], dsym[32];
};

struct Bits {
    const uint8_t *p;
    uint64_t pos, end;
    uint64_t bb;
    int bc;
};

KGZ void refill(Bits &b)
{
    while (b.bc <= 56) {
        const uint64_t byte = b.pos < b.end ? b.p[b.pos] : 0u;      // zeros past the end; the caller checks pos against end
        b.pos++;
        b.bb |= byte << b.bc;
        b.bc += 8;
    }
}
KGZ uint32_t take(Bits &b, int n)      // n <= 32, bits are there
{
    const uint32_t v = (uint32_t)(b.bb & ((1ull << n) - 1ull));
    b.bb >>= n;
    b.bc -= n;
    return v;
}

KGZ uint32_t reverse_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// canonical Huffman tables from code lengths; false: over-subscribed, or incomplete with more than one code (the fixed distance
// code of RFC 1951 3.2.6 is incomplete by definition: `fixed`)
KGZ bool build(const uint8_t *len, int n, uint16_t *count, uint16_t *sym, uint16_t *fast, int root, bool fixed = false)
{
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[len[i]]++;
    for (int i = 0; i < (1 << root); i++) fast[i] = 0;
    if (count[0] == n) return true;                       // no codes at all: legal for distances of a literal-only block
    int left = 1;
    for (int l = 1; l < 16; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    if (left > 0 && n - count[0] != 1 && !fixed) return false;
    uint16_t offs[16], next[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l < 16; l++) {
        code = (code + count[l - 1]) << 1;
        if (l == 1) code = 0;
        next[l] = (uint16_t)code;
    }
    // (next[l] = first code of length l: code_1 = 0, code_l = (code_{l-1} + count_{l-1}) << 1)
    for (int s = 0; s < n; s++) {
        const int l = len[s];
        if (!l) continue;
        sym[offs[l]++] = (uint16_t)s;
        const uint32_t c = next[l]++;
        if (l <= root) {
            const uint32_t r = reverse_bits(c, l);
            for (uint32_t k = r; k < (1u << root); k += 1u << l) fast[k] = (uint16_t)((s << 4) | l);
        }
    }
    return true;
}

// one symbol; -1: invalid code
KGZ int decode(Bits &b, const uint16_t *fast, int root, const uint16_t *count, const uint16_t *sym)
{
    const uint32_t e = fast[b.bb & ((1u << root) - 1u)];
    if (e & 15u) {
        b.bb >>= (e & 15u);
        b.bc -= (int)(e & 15u);
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;
    uint64_t bits = b.bb;
    for (int l = 1; l <= 15; l++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) {
            b.bb >>= l;
            b.bc -= l;
            return sym[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

KGZ uint32_t crc32_bytes(const uint32_t *tab, uint32_t crc, const uint8_t *p, uint64_t n)      // tab: the 256-entry table of 0xEDB88320
{
    crc = ~crc;
    for (uint64_t i = 0; i < n; i++) crc = tab[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}
KGZ void crc32_table(uint32_t *tab, int first, int step)
{
    for (int i = first; i < 256; i += step) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[i] = c;
    }
}

// one raw DEFLATE stream from b into out[o ..]; returns kOk with o advanced
KGZ int inflate_stream(Bits &b, Tables &T, uint8_t *out, uint64_t &o, uint64_t cap)
{
    const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const uint64_t start = o;
    for (;;) {
        refill(b);
        const uint32_t last = take(b, 1), type = take(b, 2);
        if (type == 0) {
            take(b, b.bc & 7);                            // to the byte boundary
            refill(b);
            const uint32_t n = take(b, 16), nn = take(b, 16);
            if ((n ^ 0xffffu) != nn) return kBadData;
            if (o + n > cap) return kOutputFull;
            for (uint32_t i = 0; i < n; i++) {
                if (b.bc < 8) refill(b);
                out[o++] = (uint8_t)take(b, 8);
            }
            if (b.pos - (uint64_t)(b.bc >> 3) > b.end) return kTruncated;
        } else if (type == 1 || type == 2) {
            uint8_t lens[320];
            int nlen, ndist;
            if (type == 1) {
                for (int i = 0; i < 144; i++) lens[i] = 8;
                for (int i = 144; i < 256; i++) lens[i] = 9;
                for (int i = 256; i < 280; i++) lens[i] = 7;
                for (int i = 280; i < 288; i++) lens[i] = 8;
                for (int i = 0; i < 30; i++) lens[288 + i] = 5;
                nlen = 288; ndist = 30;
            } else {
                nlen = (int)take(b, 5) + 257;
                ndist = (int)take(b, 5) + 1;
                const int ncode = (int)take(b, 4) + 4;
                if (nlen > 286 || ndist > 30) return kBadData;
                uint8_t cl[19];
                for (int i = 0; i < 19; i++) cl[i] = 0;
                refill(b);
                for (int i = 0; i < ncode; i++) {
                    if (b.bc < 3) refill(b);
                    cl[order[i]] = (uint8_t)take(b, 3);
                }
                // the code-length code goes through the distance tables' memory (7-bit codes, 19 symbols)
                if (!build(cl, 19, T.dcount, T.dsym, T.dfast, 7)) return kBadData;
                int i = 0;
                while (i < nlen + ndist) {
                    refill(b);
                    const int s = decode(b, T.dfast, 7, T.dcount, T.dsym);
                    if (s < 0) return kBadData;
                    if (s < 16) lens[i++] = (uint8_t)s;
                    else {
                        int rep, v = 0;
                        if (s == 16) {
                            if (i == 0) return kBadData;
                            v = lens[i - 1];
                            rep = 3 + (int)take(b, 2);
                        } else if (s == 17) rep = 3 + (int)take(b, 3);
                        else rep = 11 + (int)take(b, 7);
                        if (i + rep > nlen + ndist) return kBadData;
                        while (rep--) lens[i++] = (uint8_t)v;
                    }
                    if (b.pos > b.end + 8) return kTruncated;
                }
                if (lens[256] == 0) return kBadData;      // no end-of-block code
            }
            if (!build(lens, nlen, T.lcount, T.lsym, T.lfast, kLRoot)) return kBadData;
            if (!build(lens + nlen, ndist, T.dcount, T.dsym, T.dfast, kDRoot, type == 1)) return kBadData;
            for (;;) {
                if (b.bc < 48) refill(b);
                const int s = decode(b, T.lfast, kLRoot, T.lcount, T.lsym);
                if (s < 256) {
                    if (s < 0) return kBadData;
                    if (o >= cap) return kOutputFull;
                    out[o++] = (uint8_t)s;
                    continue;
                }
                if (s == 256) break;
                const int li = s - 257;
                if (li >= 29) return kBadData;
                const uint32_t len = lbase[li] + take(b, lext[li]);
                const int ds = decode(b, T.dfast, kDRoot, T.dcount, T.dsym);
                if (ds < 0 || ds >= 30) return kBadData;
                const uint64_t dist = dbase[ds] + take(b, dext[ds]);
                if (dist > o - start) return kBadData;    // (a member never reaches back into the one before)
                if (o + len > cap) return kOutputFull;
                const uint8_t *src = out + o - dist;
                uint8_t *dst = out + o;
                for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
                o += len;
                if (b.pos > b.end + 8) return kTruncated;
            }
            if (b.pos - (uint64_t)(b.bc >> 3) > b.end) return kTruncated;
        } else return kBadData;
        if (last) return kOk;
    }
}

// a whole .gz file (one or more members) -> out; *out_len = decoded bytes
KGZ int gunzip(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, Tables &T, const uint32_t *crc_tab, uint64_t *out_len)
{
    uint64_t pos = 0, o = 0;
    int members = 0;
    while (pos < n) {
        if (n - pos < 18 || in[pos] != 0x1f || in[pos + 1] != 0x8b || in[pos + 2] != 8) {
            if (members == 0) return kBadHeader;
            bool zeros = true;                            // gzip ignores zero padding after the last member
            for (uint64_t i = pos; i < n; i++) zeros &= in[i] == 0;
            if (zeros) break;
            return kBadHeader;
        }
        const uint32_t flg = in[pos + 3];
        if (flg & 0xe0u) return kBadHeader;
        pos += 10;
        if (flg & 4u) {
            if (pos + 2 > n) return kTruncated;
            pos += 2 + ((uint64_t)in[pos] | ((uint64_t)in[pos + 1] << 8));
        }
        if (flg & 8u) { while (pos < n && in[pos]) pos++; pos++; }
        if (flg & 16u) { while (pos < n && in[pos]) pos++; pos++; }
        if (flg & 2u) pos += 2;
        if (pos >= n) return kTruncated;
        Bits b;
        b.p = in; b.pos = pos; b.end = n; b.bb = 0; b.bc = 0;
        const uint64_t o0 = o;
        const int rc = inflate_stream(b, T, out, o, cap);
        if (rc != kOk) return rc;
        pos = b.pos - (uint64_t)(b.bc >> 3);              // bytes still whole in the bit buffer were not consumed
        if (pos + 8 > n) return kTruncated;
        const uint32_t crc = (uint32_t)in[pos] | ((uint32_t)in[pos + 1] << 8) | ((uint32_t)in[pos + 2] << 16) | ((uint32_t)in[pos + 3] << 24);
        const uint32_t isz = (uint32_t)in[pos + 4] | ((uint32_t)in[pos + 5] << 8) | ((uint32_t)in[pos + 6] << 16) | ((uint32_t)in[pos + 7] << 24);
        pos += 8;
        if ((uint32_t)(o - o0) != isz) return kBadSize;
        if (crc_tab && crc32_bytes(crc_tab, 0, out + o0, o - o0) != crc) return kBadCrc;
        members++;
    }
    *out_len = o;
    return members ? kOk : kBadHeader;
}

struct Job { uint64_t in_off, in_len, out_off, out_cap; };      // one file: compressed bytes -> its place in the text buffer
struct Result { uint64_t out_len; int32_t status, pad; };

#ifdef __CUDACC__
// One file per thread, pulled by ticket (the host orders the jobs largest first).  A CTA is one warp of which the first
// `active` lanes decode: a stream is latency-bound (table lookup -> shift -> next lookup), so a batch of a few hundred files is
// spread one per warp over every SM, and only a batch of many thousands packs 32 files into a warp.  Tables in shared memory.
__global__ void __launch_bounds__(32) gunzip_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const Job *__restrict__ jobs,
                                                    Result *__restrict__ res, uint32_t n, uint32_t active, uint32_t *__restrict__ ticket, int check_crc)
{
    extern __shared__ __align__(16) uint8_t gz_smem[];
    uint32_t *crc_tab = reinterpret_cast<uint32_t *>(gz_smem);
    crc32_table(crc_tab, (int)threadIdx.x, 32);
    __syncwarp();
    if (threadIdx.x >= active) return;
    Tables &T = *(reinterpret_cast<Tables *>(gz_smem + 1024) + threadIdx.x);
    for (;;) {
        const uint32_t i = atomicAdd(ticket, 1u);
        if (i >= n) return;
        const Job j = jobs[i];
        uint64_t len = 0;
        const int rc = gunzip(in + j.in_off, j.in_len, out + j.out_off, j.out_cap, T, check_crc ? crc_tab : nullptr, &len);
        Result r;
        r.out_len = len; r.status = rc; r.pad = 0;
        res[i] = r;
    }
}

// '\n' from the end of every file's text to the next 16-byte boundary (the scan reads whole 16-byte groups)
__global__ void pad_text_kernel(uint8_t *__restrict__ text, const uint64_t *__restrict__ goff, const uint64_t *__restrict__ glen, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t e = goff[i] + glen[i], e16 = (e + 15) & ~15ull;
    for (uint64_t p = e; p < e16; p++) text[p] = '\n';
}
#endif

}  // namespace gz
}  // namespace kssd
