/*
 * kssd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see kssd_oracle.h).
 *
 * Scalar CPU restatement of the reference algorithm, written from SURVEY.md s8a and the cited
 * reference lines.  It works on in-memory byte buffers (the reference streams through
 * popen("zcat -fc")); the streaming artefacts that are well defined (fgets line limits, EOF
 * handling of the last fastq record) are reproduced, the one that is undefined behaviour in the
 * reference (a '>' header straddling a 64 KiB refill reads seqin_buff[-1], iseq2comem.c:225-235)
 * is given its intended meaning: the header is skipped up to the next '\n'.
 */
#include "kssd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* global_basic.c:74-81 */
static const uint32_t k_primer[25] = {
    251u, 509u, 1021u, 2039u, 4093u, 8191u, 16381u, 32749u, 65521u, 131071u, 262139u, 524287u,
    1048573u, 2097143u, 4194301u, 8388593u, 16777213u, 33554393u, 67108859u, 134217689u,
    268435399u, 536870909u, 1073741789u, 2147483647u, 4294967291u};

#define CTX_SPC_USE_L 8          /* global_basic.h:45-47 */
#define LD_FCTR 0.6              /* global_basic.h:49    */
#define MIN_SUBCTX_DIM_SMP_SZ 4096u /* command_shuffle.h:29 */
#define HIBIT 0x8000000000000000ULL

size_t orc_ctx_sizeof(void) { return sizeof(orc_ctx_t); }

int orc_ctx_init(orc_ctx_t *c, int k, int s, int L, int component_sz, const int32_t *shuf)
{
    memset(c, 0, sizeof(*c));
    c->k = k; c->s = s; c->L = L; c->component_sz = component_sz; c->shuf = shuf;
    c->TL = 2 * k;
    c->out = k - s;
    c->crvsaddmove = 4 * k - 2;
    c->tupmask = ~0ULL >> (64 - 4 * k);
    c->domask = ((1ULL << (4 * s)) - 1) << (2 * c->out);
    c->undomask = ((1ULL << (2 * c->out)) - 1) << (2 * (k + s));
    uint64_t subspace = 1ULL << (4 * (s - L));
    c->dim_end = (uint32_t)(subspace > MIN_SUBCTX_DIM_SMP_SZ ? subspace : MIN_SUBCTX_DIM_SMP_SZ);
    c->component_num = (k - L > component_sz) ? (1 << (4 * (k - L - component_sz))) : 1;
    c->comp_code_bits = (k - L > component_sz) ? 4 * (k - L - component_sz) : 0;
    int pi = 4 * (k - L) - CTX_SPC_USE_L - 7;
    if (pi < 0 || pi > 24) return -1;
    c->hashsize = k_primer[pi];
    c->hashlimit = (uint32_t)(c->hashsize * LD_FCTR);
    return 0;
}

static inline int base_code(uint8_t ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/* Sampling test + re-encoding of one canonical 2k-mer (iseq2comem.c:245-253).
 * Returns 1 and sets *dr when the k-mer is sampled. */
static inline int sample_kmer(const orc_ctx_t *c, uint64_t fwd, uint64_t rc, uint64_t *dr)
{
    uint64_t u = fwd < rc ? fwd : rc;
    uint32_t inner = (uint32_t)((u & c->domask) >> (2 * c->out));
    int64_t pf = c->shuf[inner];
    if (pf < 0 || pf >= (int64_t)c->dim_end) return 0;
    uint64_t right = u & ((1ULL << (2 * c->out)) - 1);
    *dr = (((u & c->undomask) + (right << (2 * c->TL - 4 * c->out))) >> (4 * c->L)) + (uint64_t)pf;
    return 1;
}

/* Double-hash probe sequence, global_basic.h:228-230. */
static inline uint32_t probe(uint64_t key, uint32_t i, uint32_t H)
{
    return (uint32_t)(((key % H) + (uint64_t)i * (1 + key % (H - 1))) % H);
}

int orc_fasta2co(const orc_ctx_t *c, const uint8_t *buf, size_t len, int uniq, uint64_t *co)
{
    const uint32_t H = c->hashsize;
    memset(co, 0, (size_t)H * sizeof(uint64_t));
    uint64_t fwd = 0, rc = 0, run = 0, dr;
    uint32_t keycount = 0;
    size_t p = 0;
    while (p < len) {
        uint8_t ch = buf[p++];
        int b = base_code(ch);
        if (b < 0) {
            if (ch == '\n' || ch == '\r') continue;          /* skipped, no break */
            if (ch == '>') {                                   /* header: skip to end of line */
                /* the reference leaves pos on the byte it examines; the '>' itself is examined
                 * first, then following bytes until a '\n' is seen; EOF inside is an error. */
                while (p < len && buf[p] != '\n') p++;
                if (p >= len) return -2;
                p++;
            }
            run = 0;                                           /* any other byte breaks */
            continue;
        }
        fwd = ((fwd << 2) | (uint64_t)b) & c->tupmask;
        rc = (rc >> 2) + (((uint64_t)b ^ 3ULL) << c->crvsaddmove);
        if (++run < (uint64_t)c->TL) continue;
        if (!sample_kmer(c, fwd, rc, &dr)) continue;
        for (uint32_t i = 0; i < H; i++) {
            uint32_t n = probe(dr, i, H);
            if (co[n] == 0) {                 /* note: dr == 0 re-enters here forever (dropped) */
                co[n] = dr;
                if (++keycount > c->hashlimit) return -1;
                break;
            }
            if (!uniq) {
                if (co[n] == dr) break;
            } else if ((co[n] | HIBIT) == (dr | HIBIT)) {
                co[n] |= HIBIT;
                break;
            }
        }
    }
    return 0;
}

/* iseq2comem.c:78-186 reads2mco (--byread): FASTA-formatted reads, one sketch per '>' record, NO hash table --
 * every sampled k-mer is emitted at once, duplicates and code 0 included (:170-172).  Emission i carries the id
 * (drtuple >> comp_code_bits), its component (drtuple % component_num) and the record counter `readn` at that
 * moment (0 = before the first header); the Python side turns read_of into the inclusive per-component index
 * the reference writes at :175-180.  Returns the number of emissions, -2 on a header that runs into EOF
 * (:152), -4 if cap is too small. */
long orc_reads2mco(const orc_ctx_t *c, const uint8_t *buf, size_t len, uint32_t *ids, int32_t *comp,
                   uint64_t *read_of, size_t cap, uint64_t *n_reads)
{
    uint64_t fwd = 0, rc = 0, run = 0, dr, readn = 0;
    size_t p = 0, n = 0;
    while (p < len) {
        uint8_t ch = buf[p++];
        int b = base_code(ch);
        if (b < 0) {
            if (ch == '\n' || ch == '\r') continue;
            if (ch == '>') {
                readn++;                                        /* :140 */
                while (p < len && buf[p] != '\n') p++;
                if (p >= len) return -2;
                p++;
            }
            run = 0;
            continue;
        }
        fwd = ((fwd << 2) | (uint64_t)b) & c->tupmask;
        rc = (rc >> 2) + (((uint64_t)b ^ 3ULL) << c->crvsaddmove);
        if (++run < (uint64_t)c->TL) continue;
        if (!sample_kmer(c, fwd, rc, &dr)) continue;
        if (n >= cap) return -4;
        ids[n] = (uint32_t)(dr >> c->comp_code_bits);
        comp[n] = (int32_t)(dr % (uint64_t)c->component_num);
        read_of[n] = readn;
        n++;
    }
    if (n_reads) *n_reads = readn;
    return (long)n;
}

/* ---- fgets emulation over a memory buffer (for the fastq readers) ---- */
typedef struct { const uint8_t *b; size_t n, p; int eof; } memf_t;

/* Reads at most cap-1 bytes through the next '\n'. Returns length, or -1 when nothing could be
 * read (fgets -> NULL, destination untouched). Sets eof exactly when stdio would. */
static long mem_fgets(memf_t *f, const uint8_t **line, int cap)
{
    if (f->p >= f->n) { f->eof = 1; return -1; }
    size_t start = f->p, lim = f->p + (size_t)(cap - 1);
    while (f->p < f->n && f->p < lim) {
        if (f->b[f->p++] == '\n') { *line = f->b + start; return (long)(f->p - start); }
    }
    if (f->p >= f->n && f->p < lim) f->eof = 1;   /* hit end of data before cap or newline */
    *line = f->b + start;
    return (long)(f->p - start);
}

int orc_fastq2co(const orc_ctx_t *c, const uint8_t *buf, size_t len, int Q, int M, uint64_t *co,
                 int *reads_detected)
{
    enum { LEN = 20000, CT_BIT = 4 };
    const uint64_t CT_MAX = 0xfULL;
    const uint32_t H = c->hashsize;
    memset(co, 0, (size_t)H * sizeof(uint64_t));
    if (reads_detected) *reads_detected = 0;
    if (M >= (int)CT_MAX) return -3;
    memf_t f = {buf, len, 0, 0};
    const uint8_t *seq = NULL, *qual = NULL, *tmp;
    long sl = 0, ql = 0, r;
    /* first record: the reference ignores fgets failures; stale buffers keep their content */
    if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { seq = tmp; sl = r; }
    if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { seq = tmp; sl = r; }
    if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { qual = tmp; ql = r; }
    if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { qual = tmp; ql = r; }
    uint64_t fwd = 0, rc = 0, run = 0, dr;
    int line_num = 0;
    if (seq == NULL) return 0;
    for (long pos = 0; pos < sl; pos++) {
        uint8_t ch = seq[pos];
        if (ch == '\n') {
            if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { seq = tmp; sl = r; }
            if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { seq = tmp; sl = r; }
            if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { qual = tmp; ql = r; }
            if ((r = mem_fgets(&f, &tmp, LEN)) >= 0) { qual = tmp; ql = r; }
            line_num += 4;
            if (f.eof) break;          /* last record is dropped when its 4th line has no '\n' */
            run = 0;
            pos = -1;
            continue;
        }
        int b = base_code(ch);
        /* qual[pos] is a signed char compared to int Q, raw ASCII (no -33) */
        int q = (qual && pos < ql) ? (int)(int8_t)qual[pos] : 0;
        if (b < 0 || q < Q) { run = 0; continue; }
        fwd = ((fwd << 2) | (uint64_t)b) & c->tupmask;
        rc = (rc >> 2) + (((uint64_t)b ^ 3ULL) << c->crvsaddmove);
        if (++run < (uint64_t)c->TL) continue;
        if (!sample_kmer(c, fwd, rc, &dr)) continue;
        for (uint32_t i = 0; i < H; i++) {
            uint32_t n = probe(dr, i, H);
            if (co[n] == 0) {
                co[n] = (M == 1) ? ((dr << CT_BIT) | CT_MAX) : ((dr << CT_BIT) + 1ULL);
                break;                 /* keycount is never incremented in the reference */
            }
            if ((co[n] >> CT_BIT) == dr) {
                if ((co[n] & CT_MAX) == CT_MAX) break;
                co[n] += 1ULL;
                if (!((int)(co[n] & CT_MAX) < M)) co[n] |= CT_MAX;
                break;
            }
        }
    }
    if (reads_detected) *reads_detected = line_num;
    return 0;
}

int orc_shortreads2koc(const orc_ctx_t *c, const uint8_t *buf, size_t len, uint64_t *co)
{
    enum { FQ_LEN = 4096, OCCRC_BIT = 16 };
    const uint64_t OCCRC_MAX = 0xffffULL;
    const uint32_t H = c->hashsize;
    memset(co, 0, (size_t)H * sizeof(uint64_t));
    memf_t f = {buf, len, 0, 0};
    uint32_t keycount = 0;
    for (;;) {
        const uint8_t *t, *seq;
        long sl;
        if (mem_fgets(&f, &t, FQ_LEN) < 0) break;
        if ((sl = mem_fgets(&f, &seq, FQ_LEN)) < 0) break;
        if (mem_fgets(&f, &t, FQ_LEN) < 0) break;
        if (mem_fgets(&f, &t, FQ_LEN) < 0) break;
        uint64_t fwd = 0, rc = 0, run = 0, dr;
        for (long pos = 0; pos < sl && seq[pos] != '\n'; pos++) {
            int b = base_code(seq[pos]);
            if (b < 0) { run = 0; continue; }
            fwd = ((fwd << 2) | (uint64_t)b) & c->tupmask;
            rc = (rc >> 2) + (((uint64_t)b ^ 3ULL) << c->crvsaddmove);
            if (++run < (uint64_t)c->TL) continue;
            if (!sample_kmer(c, fwd, rc, &dr)) continue;
            for (uint32_t i = 0; i < H; i++) {
                uint32_t n = probe(dr, i, H);
                if (co[n] == 0) {
                    co[n] = (dr << OCCRC_BIT) + 1ULL;
                    if (++keycount > c->hashlimit) return -1;
                    break;
                }
                if ((co[n] >> OCCRC_BIT) == dr) {
                    if ((co[n] & OCCRC_MAX) < OCCRC_MAX) co[n] += 1ULL;
                    break;
                }
            }
        }
    }
    return 0;
}

size_t orc_write_co(const orc_ctx_t *c, const uint64_t *co, int mode, uint32_t *ids, int32_t *comp,
                    uint16_t *abund)
{
    size_t w = 0;
    const uint64_t ncomp = (uint64_t)c->component_num;
    for (uint32_t i = 0; i < c->hashsize; i++) {
        uint64_t v = co[i];
        if (mode == 0) {
            if (v == 0 || v >= HIBIT) continue;
            ids[w] = (uint32_t)(v >> c->comp_code_bits);
            comp[w] = (int32_t)(v % ncomp);
        } else if (mode == 1) {
            if ((v & 0xfULL) != 0xfULL) continue;
            ids[w] = (uint32_t)(v >> (c->comp_code_bits + 4));
            comp[w] = (int32_t)((v >> 4) % ncomp);
        } else {
            if (v == 0) continue;
            ids[w] = (uint32_t)(v >> (c->comp_code_bits + 16));
            comp[w] = (int32_t)((v >> 16) % ncomp);
            if (abund) abund[w] = (uint16_t)(v & 0xffffULL);
        }
        w++;
    }
    return w;
}

void orc_combco2mco(const uint32_t *combco, const uint64_t *cbdcoindex, int cofnum, int component_sz,
                    uint64_t *dense_incl, uint32_t *mco)
{
    const size_t comp_sz = (size_t)1 << (4 * component_sz);
    const size_t total = cbdcoindex[cofnum];
    uint64_t *cnt = dense_incl ? dense_incl : (uint64_t *)calloc(comp_sz, sizeof(uint64_t));
    if (dense_incl) memset(cnt, 0, comp_sz * sizeof(uint64_t));
    for (size_t i = 0; i < total; i++) cnt[combco[i]]++;
    for (size_t n = 1; n < comp_sz; n++) cnt[n] += cnt[n - 1];        /* inclusive, co2mco.c:57 */
    /* fill postings back to front so that gids end up ascending within a code */
    uint64_t *cur = (uint64_t *)malloc(comp_sz * sizeof(uint64_t));
    memcpy(cur, cnt, comp_sz * sizeof(uint64_t));
    for (int j = cofnum - 1; j >= 0; j--)
        for (size_t i = cbdcoindex[j + 1]; i-- > cbdcoindex[j];)
            mco[--cur[combco[i]]] = (uint32_t)j;
    free(cur);
    if (!dense_incl) free(cnt);
}

void orc_dist_counts_dense(const uint32_t *qcodes, const uint64_t *qindex, int qnum,
                           const uint64_t *dense_incl, const uint32_t *mco, int refnum,
                           uint32_t *ct, int nthreads)
{
#pragma omp parallel for num_threads(nthreads) schedule(guided)
    for (int q = 0; q < qnum; q++) {
        uint32_t *row = ct + (size_t)q * (size_t)refnum;
        for (uint64_t n = qindex[q]; n < qindex[q + 1]; n++) {
            uint32_t ind = qcodes[n];
            uint64_t s = ind > 0 ? dense_incl[ind - 1] : 0;
            for (uint64_t g = s; g < dense_incl[ind]; g++) row[mco[g]]++;
        }
    }
}

void orc_dist_counts_csr(const uint32_t *qcodes, const uint64_t *qindex, int qnum,
                         const uint32_t *ucodes, const uint64_t *uoff, size_t nuniq,
                         const uint32_t *mco, int refnum, uint32_t *ct, int nthreads)
{
#pragma omp parallel for num_threads(nthreads) schedule(guided)
    for (int q = 0; q < qnum; q++) {
        uint32_t *row = ct + (size_t)q * (size_t)refnum;
        for (uint64_t n = qindex[q]; n < qindex[q + 1]; n++) {
            uint32_t ind = qcodes[n];
            size_t lo = 0, hi = nuniq;
            while (lo < hi) { size_t mid = (lo + hi) >> 1; if (ucodes[mid] < ind) lo = mid + 1; else hi = mid; }
            if (lo == nuniq || ucodes[lo] != ind) continue;
            for (uint64_t g = uoff[lo]; g < uoff[lo + 1]; g++) row[mco[g]]++;
        }
    }
}

int orc_output_ctrl(uint32_t X, uint32_t Y, uint32_t I, int metric_kind, int correction,
                    int kmerlen, int dim_reduct_len, double dthreshold, uint64_t cmprsn_num,
                    double out[9])
{
    const double alp = 4.0;
    double rs = 0;
    if (correction) {
        uint32_t xo = X - I, yo = Y - I;
        double px = 1 - pow(1 - 1 / pow(alp, (double)(kmerlen - dim_reduct_len)), (double)xo);
        double py = 1 - pow(1 - 1 / pow(alp, (double)(kmerlen - dim_reduct_len)), (double)yo);
        rs = px * py * (double)(xo + yo) / (px + py - 2 * px * py);
    }
    uint32_t tmp = metric_kind == 0 ? X + Y - I : (X < Y ? X : Y);
    double m = ((double)I - rs) / tmp;
#define ORC_GM(y) (metric_kind == 0 ? 1 / (2 * (y)) + 0.5 : 1 / (y))
    double dist = log(ORC_GM(m)) / kmerlen;
    if (dist > 1) dist = 1;
    for (int i = 0; i < 9; i++) out[i] = 0;
    out[0] = m; out[1] = dist; out[8] = rs;
    if (dist > dthreshold) return 0;
    double sd = pow(m * (1 - m) / tmp, 0.5);
    double pv = 0.5 * erfc(m / sd * pow(0.5, 0.5));
    out[2] = pv;
    out[3] = pv * (double)cmprsn_num;
    double c1 = m - 1.96 * sd, c2 = m + 1.96 * sd;
    out[4] = c1; out[5] = c2;
    out[6] = log(ORC_GM(c2)) / kmerlen;
    out[7] = log(ORC_GM(c1)) / kmerlen;
    return 1;
}


/* ---- kssd set (command_set.c) ---- */
/* sketch_union (:226-293) / uniq_sketch_union (:374-443) for ONE component: the codes of combco.<c> that occur
 * (uniq: that occur exactly once in the whole file), ascending.  code_bits = 4*COMPONENT_SZ.  Returns the count. */
size_t orc_set_union(const uint32_t *combco, size_t n, int uniq, int code_bits, uint32_t *out)
{
    const size_t words = ((size_t)1 << code_bits) / 64;
    uint64_t *seen = (uint64_t *)calloc(words, 8), *once = NULL;
    if (uniq) { once = (uint64_t *)malloc(words * 8); memset(once, 0xff, words * 8); }
    for (size_t i = 0; i < n; i++) {
        const uint64_t bit = 0x8000000000000000ULL >> (combco[i] % 64);
        if (uniq && (seen[combco[i] / 64] & bit)) once[combco[i] / 64] &= ~bit;
        seen[combco[i] / 64] |= bit;
    }
    size_t k = 0;
    for (size_t w = 0; w < words; w++) {
        const uint64_t v = uniq ? (seen[w] & once[w]) : seen[w];
        if (!v) continue;
        for (int b = 0; b < 64; b++)
            if ((0x8000000000000000ULL >> b) & v) out[k++] = (uint32_t)(64 * w + b);
    }
    free(seen);
    free(once);
    return k;
}

/* sketch_operate (:294-373) for ONE component: every genome keeps, in order, the codes whose membership in the pan
 * sketch equals `intersect` (1 = -i, 0 = -s); out_index is the rebuilt combco.index (n_genomes + 1). */
void orc_set_operate(const uint32_t *combco, const uint64_t *index, int n_genomes, const uint32_t *pan, size_t n_pan,
                     int intersect, int code_bits, uint32_t *out, uint64_t *out_index)
{
    const size_t words = ((size_t)1 << code_bits) / 64;
    uint64_t *dict = (uint64_t *)calloc(words, 8);
    for (size_t i = 0; i < n_pan; i++) dict[pan[i] / 64] |= 0x8000000000000000ULL >> (pan[i] % 64);
    out_index[0] = 0;
    for (int g = 0; g < n_genomes; g++) {
        out_index[g + 1] = out_index[g];
        for (uint64_t i = index[g]; i < index[g + 1]; i++) {
            const int in = (dict[combco[i] / 64] & (0x8000000000000000ULL >> (combco[i] % 64))) != 0;
            if (in == intersect) out[out_index[g + 1]++] = combco[i];
        }
    }
    free(dict);
}
