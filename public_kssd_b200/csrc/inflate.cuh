// gzip / DEFLATE (RFC 1952 / 1951) decoding, one compressed file per thread -- what the reference gets from popen("zcat -fc")
// (iseq2comem.c:187-200, :283-290).  A batch of Stage I is hundreds of independent .gz genomes: decoding them on the GPU lets the
// compressed bytes cross PCIe (3-4x fewer) and takes inflate off the host cores.  A single stream is serial (every symbol's
// position depends on the one before), so the parallelism is over files; `stage1_files.cuh` picks this path when a call holds
// enough of them and leaves the others to zlib on the host.
//
// What one thread does per symbol is therefore what counts:
//   * the bit buffer (64 bits) is refilled 32 bits at a time from aligned words, the next word requested one refill ahead;
//   * literal/length codes go through a 10-bit table whose entries hold up to TWO literals, or a length code's base and
//     extra-bit count; distance codes through an 8-bit table of base / extra bits; longer codes by the canonical count/symbol walk;
//   * match bytes are loaded in groups before they are stored (independent loads; an overlapping match stays bytewise);
//   * offsets inside a file are 32-bit (a file decodes to < 4 GiB here; the caller sends larger ones to zlib).
// Measured (B200, `gzip -1` of random DNA = nothing but 3-6 byte matches): ~1,000 cycles per match however the copy is done -- a ring
// of the last 8 KiB in shared memory, eight-byte stores from a register, base / extra bits folded into the table entries each left
// the 0.66-0.68 s per 5 MB file where it was: one thread issues a dependent instruction every ~10 cycles, and two files in one
// warp take turns (divergence).  What does scale is the number of WARPS, so the tables are kept small (5.8 KB: 32 streams per SM).
// Tables live in memory the caller provides (shared memory on the GPU).  Members are decoded one after another (multi-member
// files, `cat a.gz b.gz`); every member's ISIZE and CRC-32 (slicing-by-8) are checked.  The same code compiles for the host, where
// the CPU test suite runs it against zlib.  The input must be readable kInPad bytes past its end (zeros).
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
constexpr uint32_t kInPad = 16;       // readable zero bytes behind the compressed data

static const uint32_t h_crc8[2048] = {
#include "crc32_slice8.inc"
};
#ifdef __CUDACC__
__device__ const uint32_t d_crc8[2048] = {
#include "crc32_slice8.inc"
};
#endif

// build() leaves symbol | code length << 16 in an entry (0: the code is longer than the root -> canonical walk); then
//   literal/length table: literals   lit0 | lit1 << 8 | bits consumed << 16 | n << 24 (n = 1, 2)
//                         length     base | extra-bit count << 9 | code length << 16 | kLen
//                         end of block                              code length << 16 | kEob      (symbols 286, 287: kBad)
//   distance table:       base | extra-bit count << 15 | code length << 19                        (symbols 30, 31: kBad)
constexpr uint32_t kLen = 1u << 31, kEob = 1u << 30, kBad = 1u << 29;
struct Tables {
    uint32_t lfast[1 << kLRoot];
    uint32_t dfast[1 << kDRoot];
    uint16_t lcount[16], dcount[16];  // canonical walk: codes per length, symbols in code order
    uint16_t lsym[288], dsym[32];
};

#define KSSD_GZ_CONST_TABLES(Q)                                                                                                                      \
    Q const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258}; \
    Q const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};                              \
    Q const uint16_t dbase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,                               \
                                  193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};                            \
    Q const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
namespace host_tab { KSSD_GZ_CONST_TABLES(static) }
#ifdef __CUDACC__
namespace dev_tab { KSSD_GZ_CONST_TABLES(__device__) }
#endif
#ifdef __CUDA_ARCH__
namespace tab = dev_tab;
#else
namespace tab = host_tab;
#endif

struct Bits {
    const uint32_t *wp, *wend;        // next aligned word to request; first word that is not inside the padded input (zeros from there on)
    uint32_t nextw;                   // the word requested last (not in bb yet)
    uint64_t bb;
    int bc;
    uint64_t loaded;                  // bytes of the file that went into bb so far (position = loaded - bc / 8)
};

KGZ void bits_open(Bits &b, const uint8_t *in, uint64_t pos, uint64_t n)
{
    b.wend = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(in + n + kInPad) & ~(uintptr_t)3);
    b.bb = 0;
    b.bc = 0;
    b.loaded = pos;
    while ((reinterpret_cast<uintptr_t>(in + b.loaded) & 3u) != 0) {
        b.bb |= (uint64_t)in[b.loaded] << b.bc;
        b.bc += 8;
        b.loaded++;
    }
    b.wp = reinterpret_cast<const uint32_t *>(in + b.loaded);
    b.nextw = b.wp < b.wend ? *b.wp : 0u;
    b.wp++;
}
KGZ void refill(Bits &b)                                  // afterwards bc > 32
{
    if (b.bc <= 32) {
        b.bb |= (uint64_t)b.nextw << b.bc;
        b.bc += 32;
        b.loaded += 4;
        b.nextw = b.wp < b.wend ? *b.wp : 0u;      // (a damaged stream that runs past the end reads zeros until its output is full)
        b.wp++;
    }
}
KGZ uint64_t bits_pos(const Bits &b) { return b.loaded - (uint64_t)(b.bc >> 3); }      // first byte not consumed whole
KGZ uint32_t take(Bits &b, int n)                         // n <= 16, bits are there
{
    const uint32_t v = (uint32_t)b.bb & ((1u << n) - 1u);
    b.bb >>= n;
    b.bc -= n;
    return v;
}

KGZ uint32_t bfe(uint32_t v, uint32_t pos, uint32_t n)      // n bits of v from bit pos on (pos + n <= 32)
{
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(n));
    return r;
#else
    return (v >> pos) & (uint32_t)((1ull << n) - 1ull);
#endif
}
KGZ uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)      // bits sh .. sh + 31 of hi:lo, sh < 32
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}
KGZ void drop(Bits &b, uint32_t n)
{
    b.bb >>= n;
    b.bc -= (int)n;
}

KGZ uint32_t reverse_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// canonical Huffman code from code lengths: count / sym for the walk, and for every code of at most `root` bits the entries
// fast[reversed code + k * 2^len] = symbol | len << 16.  false: over-subscribed, or incomplete with more than one code (the fixed
// distance code of RFC 1951 3.2.6 is incomplete by definition: `fixed`).
KGZ bool build(const uint8_t *len, int n, uint16_t *count, uint16_t *sym, uint32_t *fast, int root, bool fixed)
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
    for (int l = 1; l < 16; l++) {                        // first code of length l
        code = l == 1 ? 0u : (code + count[l - 1]) << 1;
        next[l] = (uint16_t)code;
    }
    for (int s = 0; s < n; s++) {
        const int l = len[s];
        if (!l) continue;
        sym[offs[l]++] = (uint16_t)s;
        const uint32_t c = next[l]++;
        if (l <= root) {
            const uint32_t e = (uint32_t)s | ((uint32_t)l << 16);
            for (uint32_t k = reverse_bits(c, l); k < (1u << root); k += 1u << l) fast[k] = e;
        }
    }
    return true;
}

// literal/length table into its final form; two literals per entry where the bits of the index hold two whole literal codes
// (descending: the entry read for the second literal is not converted yet)
KGZ void finish_litlen(uint32_t *fast)
{
    for (int i = (1 << kLRoot) - 1; i >= 0; i--) {
        const uint32_t e = fast[i], l1 = (e >> 16) & 31u, s1 = e & 0x1ffu;
        if (l1 == 0) continue;
        uint32_t f;
        if (s1 < 256u) {
            f = s1 | (l1 << 16) | (1u << 24);
            if (l1 < (uint32_t)kLRoot) {
                const uint32_t e2 = fast[i >> l1], l2 = (e2 >> 16) & 31u, s2 = e2 & 0x1ffu;
                if ((e2 >> 24) == 0 && l2 != 0 && s2 < 256u && l1 + l2 <= (uint32_t)kLRoot) f = s1 | (s2 << 8) | ((l1 + l2) << 16) | (2u << 24);
            }
        } else if (s1 == 256u) f = (l1 << 16) | kEob;
        else if (s1 < 286u) f = tab::lbase[s1 - 257u] | ((uint32_t)tab::lext[s1 - 257u] << 9) | (l1 << 16) | kLen;
        else f = (l1 << 16) | kBad;
        fast[i] = f;
    }
}
KGZ void finish_dist(uint32_t *fast)
{
    for (int i = 0; i < (1 << kDRoot); i++) {
        const uint32_t e = fast[i], l = (e >> 16) & 31u, s = e & 0x1ffu;
        if (l == 0) continue;
        fast[i] = s < 30u ? tab::dbase[s] | ((uint32_t)tab::dext[s] << 15) | (l << 19) : kBad;
    }
}

// a symbol by the canonical walk (any code length); -1: invalid code
KGZ int decode_walk(Bits &b, const uint16_t *count, const uint16_t *sym)
{
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
KGZ int decode_small(Bits &b, const uint32_t *fast, int root, const uint16_t *count, const uint16_t *sym)      // tables as build() leaves them
{
    const uint32_t e = fast[(uint32_t)b.bb & ((1u << root) - 1u)], l = e >> 16;
    if (l) {
        b.bb >>= l;
        b.bc -= (int)l;
        return (int)(e & 0xffffu);
    }
    return decode_walk(b, count, sym);
}

// the output of one file
struct Out {
    uint8_t *base;
    uint32_t o;
};
KGZ void put(Out &w, uint32_t c) { w.base[w.o++] = (uint8_t)c; }

// CRC-32 of p[0 .. n): bytes up to the first 8-byte boundary one by one, then eight at a time; t: the slicing-by-8 tables
KGZ uint32_t crc32_slice8(const uint32_t *t, const uint8_t *p, uint32_t n)
{
    uint32_t crc = 0xffffffffu, i = 0;
    for (; i < n && (reinterpret_cast<uintptr_t>(p + i) & 7u); i++) crc = t[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p + i);
    for (; i + 8 <= n; i += 8, w += 2) {
        const uint32_t lo = w[0] ^ crc, hi = w[1];
        crc = t[7 * 256 + (lo & 255u)] ^ t[6 * 256 + ((lo >> 8) & 255u)] ^ t[5 * 256 + ((lo >> 16) & 255u)] ^ t[4 * 256 + (lo >> 24)] ^
              t[3 * 256 + (hi & 255u)] ^ t[2 * 256 + ((hi >> 8) & 255u)] ^ t[1 * 256 + ((hi >> 16) & 255u)] ^ t[hi >> 24];
    }
    for (; i < n; i++) crc = t[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    return ~crc;
}

// one raw DEFLATE stream from b; returns kOk with w advanced.  `end`: size of the compressed file.
KGZ int inflate_stream(Bits &b, uint64_t end, Tables &T, Out &w, uint32_t cap)
{
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const uint32_t start = w.o;
    for (;;) {
        refill(b);
        const uint32_t last = take(b, 1), type = take(b, 2);
        if (type == 0) {
            take(b, b.bc & 7);                            // to the byte boundary
            refill(b);
            const uint32_t n = take(b, 16);
            refill(b);
            const uint32_t nn = take(b, 16);
            if ((n ^ 0xffffu) != nn) return kBadData;
            if (n > cap - w.o) return kOutputFull;
            for (uint32_t i = 0; i < n; i++) {
                refill(b);
                put(w, take(b, 8));
                if ((i & 1023u) == 0 && bits_pos(b) > end) return kTruncated;
            }
            if (bits_pos(b) > end) return kTruncated;
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
                for (int i = 0; i < ncode; i++) {
                    refill(b);
                    cl[order[i]] = (uint8_t)take(b, 3);
                }
                // the code-length code goes through the distance tables' memory (7-bit codes, 19 symbols)
                if (!build(cl, 19, T.dcount, T.dsym, T.dfast, 7, false)) return kBadData;
                int i = 0;
                while (i < nlen + ndist) {
                    refill(b);
                    const int s = decode_small(b, T.dfast, 7, T.dcount, T.dsym);
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
                }
                if (bits_pos(b) > end) return kTruncated;
                if (lens[256] == 0) return kBadData;      // no end-of-block code
            }
            if (!build(lens, nlen, T.lcount, T.lsym, T.lfast, kLRoot, false)) return kBadData;
            if (!build(lens + nlen, ndist, T.dcount, T.dsym, T.dfast, kDRoot, type == 1)) return kBadData;
            finish_litlen(T.lfast);
            finish_dist(T.dfast);
            uint32_t guard = 0;
            for (;;) {
                refill(b);                                // > 32 bits: a literal/length code and its extra bits (15 + 5)
                uint32_t lo = (uint32_t)b.bb;
                const uint32_t e = T.lfast[lo & ((1u << kLRoot) - 1u)];
                uint32_t len = 0;
                if (e & kLen) {
                    const uint32_t cl = (e >> 16) & 31u, xb = (e >> 9) & 7u;
                    len = (e & 0x1ffu) + bfe(lo, cl, xb);
                    drop(b, cl + xb);
                } else if (e != 0) {
                    drop(b, (e >> 16) & 31u);
                    const uint32_t nl = (e >> 24) & 3u;
                    if (nl) {                             // one or two literals
                        if (cap - w.o < nl) return kOutputFull;
                        put(w, e & 0xffu);
                        if (nl == 2u) put(w, (e >> 8) & 0xffu);
                        continue;
                    }
                    if (e & kEob) break;
                    return kBadData;
                } else {                                  // a code longer than the root
                    const int s = decode_walk(b, T.lcount, T.lsym);
                    if (s < 0 || s >= 286) return kBadData;
                    if (s < 256) {
                        if (cap == w.o) return kOutputFull;
                        put(w, (uint32_t)s);
                        continue;
                    }
                    if (s == 256) break;
                    len = tab::lbase[s - 257] + take(b, tab::lext[s - 257]);
                }
                refill(b);                                // a distance code and its extra bits (15 + 13)
                lo = (uint32_t)b.bb;
                const uint32_t d = T.dfast[lo & ((1u << kDRoot) - 1u)];
                uint32_t dist;
                if (d != 0) {
                    if (d & kBad) return kBadData;
                    const uint32_t dl = (d >> 19) & 15u, xd = (d >> 15) & 15u;
                    dist = (d & 0x7fffu) + bfe(lo, dl, xd);
                    drop(b, dl + xd);
                } else {
                    const int ds = decode_walk(b, T.dcount, T.dsym);
                    if (ds < 0 || ds >= 30) return kBadData;
                    dist = tab::dbase[ds] + take(b, tab::dext[ds]);
                }
                if (dist > w.o - start) return kBadData;  // (a member never reaches back into the one before)
                if (len > cap - w.o) return kOutputFull;
                const uint8_t *src = w.base + (w.o - dist);
                uint8_t *dst = w.base + w.o;
                uint32_t i = 0;
                if (dist >= 8u) {
                    for (; i + 8 <= len; i += 8) {
                        const uint8_t c0 = src[i], c1 = src[i + 1], c2 = src[i + 2], c3 = src[i + 3], c4 = src[i + 4], c5 = src[i + 5], c6 = src[i + 6], c7 = src[i + 7];
                        dst[i] = c0; dst[i + 1] = c1; dst[i + 2] = c2; dst[i + 3] = c3; dst[i + 4] = c4; dst[i + 5] = c5; dst[i + 6] = c6; dst[i + 7] = c7;
                    }
                    if (i + 4 <= len) {
                        const uint8_t c0 = src[i], c1 = src[i + 1], c2 = src[i + 2], c3 = src[i + 3];
                        dst[i] = c0; dst[i + 1] = c1; dst[i + 2] = c2; dst[i + 3] = c3;
                        i += 4;
                    }
                    if (i < len) {                        // 1 .. 3 left
                        const uint8_t c0 = src[i], c1 = i + 1 < len ? src[i + 1] : (uint8_t)0, c2 = i + 2 < len ? src[i + 2] : (uint8_t)0;
                        dst[i] = c0;
                        if (i + 1 < len) dst[i + 1] = c1;
                        if (i + 2 < len) dst[i + 2] = c2;
                    }
                } else
                    for (; i < len; i++) dst[i] = src[i];
                w.o += len;
                if ((++guard & 255u) == 0 && bits_pos(b) > end) return kTruncated;
            }
            if (bits_pos(b) > end) return kTruncated;
        } else return kBadData;
        if (last) return kOk;
    }
}

// a whole .gz file (one or more members) -> out; *out_len = decoded bytes.  in[n .. n + kInPad) must be readable.
KGZ int gunzip(const uint8_t *in, uint64_t n, uint8_t *out, uint32_t cap, Tables &T, const uint32_t *crc_tab, uint32_t *out_len)
{
    uint64_t pos = 0;
    Out w;
    w.base = out; w.o = 0;
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
        bits_open(b, in, pos, n);
        const uint32_t o0 = w.o;
        const int rc = inflate_stream(b, n, T, w, cap);
        if (rc != kOk) return rc;
        pos = bits_pos(b);
        if (pos + 8 > n) return kTruncated;
        const uint32_t crc = (uint32_t)in[pos] | ((uint32_t)in[pos + 1] << 8) | ((uint32_t)in[pos + 2] << 16) | ((uint32_t)in[pos + 3] << 24);
        const uint32_t isz = (uint32_t)in[pos + 4] | ((uint32_t)in[pos + 5] << 8) | ((uint32_t)in[pos + 6] << 16) | ((uint32_t)in[pos + 7] << 24);
        pos += 8;
        if (w.o - o0 != isz) return kBadSize;
        if (crc_tab && crc32_slice8(crc_tab, out + o0, w.o - o0) != crc) return kBadCrc;
        members++;
    }
    *out_len = w.o;
    return members ? kOk : kBadHeader;
}

struct Job { uint64_t in_off, in_len, out_off, out_cap; };      // one file: compressed bytes -> its place in the text buffer
struct Result { uint64_t out_len; int32_t status, pad; };

#ifdef __CUDACC__
// One file per thread, pulled by ticket (the host orders the jobs largest first).  A CTA is one warp of which the first
// `active` lanes decode: a stream is latency-bound (table lookup -> shift -> next lookup), so a batch of a few hundred files is
// spread one per warp over every SM, and only a batch of many thousands packs several files into a warp.  Tables in shared memory.
__global__ void __launch_bounds__(32) gunzip_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const Job *__restrict__ jobs,
                                                    Result *__restrict__ res, uint32_t n, uint32_t active, uint32_t *__restrict__ ticket, int check_crc)
{
    extern __shared__ __align__(16) uint8_t gz_smem[];
    if (threadIdx.x >= active) return;
    Tables &T = *(reinterpret_cast<Tables *>(gz_smem) + threadIdx.x);
    for (;;) {
        const uint32_t i = atomicAdd(ticket, 1u);
        if (i >= n) return;
        const Job j = jobs[i];
        uint32_t len = 0;
        const int rc = gunzip(in + j.in_off, j.in_len, out + j.out_off, (uint32_t)j.out_cap, T, check_crc ? d_crc8 : nullptr, &len);
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
