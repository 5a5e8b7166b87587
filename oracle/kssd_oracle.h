/*
 * kssd_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the Kssd sketch -> index -> compare hot path,
 * used ONLY as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under public_kssd_b200/ includes, links or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py runs the unmodified reference binary
 * (oracle/_ref/kssd, built by oracle/Makefile from the reference C sources) on seeded inputs and
 * compares every function below with the files the reference writes; the resulting vectors are
 * committed under tests/golden/ so the pin also holds on boxes without /root/reference.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef KSSD_ORACLE_H
#define KSSD_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx {
    int k, s, L;              /* half_ctx_len, half_subctx_len, drlevel            */
    int component_sz;         /* COMPONENT_SZ (7 in the Makefile build)            */
    int TL;                   /* 2k                                                */
    int out;                  /* k - s                                             */
    int crvsaddmove;          /* 4k-2                                              */
    uint64_t tupmask, domask, undomask;
    uint32_t dim_end;         /* max(16^(s-L), 4096)                               */
    int component_num;
    int comp_code_bits;
    uint32_t hashsize, hashlimit;
    const int32_t *shuf;      /* 16^s entries, borrowed                            */
} orc_ctx_t;

/* iseq2comem.c:54-77 + command_dist.c:217-236. Returns 0, or -1 if primer index out of range. */
int orc_ctx_init(orc_ctx_t *c, int k, int s, int L, int component_sz, const int32_t *shuf);
size_t orc_ctx_sizeof(void);

/* iseq2comem.c:188-273 (uniq=0) / :616-703 (uniq=1).  `co` has hashsize slots, is cleared here.
 * Returns 0, -1 on "context space too crowd", -2 on a header that runs into EOF. */
int orc_fasta2co(const orc_ctx_t *c, const uint8_t *buf, size_t len, int uniq, uint64_t *co);

/* iseq2comem.c:277-356.  Returns 0 / -1; *reads_detected = line_num as printed at :351. */
int orc_fastq2co(const orc_ctx_t *c, const uint8_t *buf, size_t len, int Q, int M, uint64_t *co,
                 int *reads_detected);

/* iseq2comem.c:554-615 with p == 1 (the only deterministic setting). */
int orc_shortreads2koc(const orc_ctx_t *c, const uint8_t *buf, size_t len, uint64_t *co);

/* iseq2comem.c:78-186 reads2mco (--byread): every sampled k-mer of a FASTA-formatted read file in stream order,
 * with its component and the '>' record counter at emission.  Returns the count, -2 / -4 on error. */
long orc_reads2mco(const orc_ctx_t *c, const uint8_t *buf, size_t len, uint32_t *ids, int32_t *comp,
                   uint64_t *read_of, size_t cap, uint64_t *n_reads);

/* Writers, slot order.  mode 0: wrt_co2cmpn_use_inn_subctx (iseq2comem.c:525-551);
 * mode 1: write_fqco2file (:499-524); mode 2: write_fqkoc2files (:435-471, fills abund).
 * Outputs ids[i], comp[i] (component of entry i), abund[i] (mode 2, may be NULL otherwise).
 * Returns the number written (the reference's return value / ctx_ct_list[i]). */
size_t orc_write_co(const orc_ctx_t *c, const uint64_t *co, int mode, uint32_t *ids,
                    int32_t *comp, uint16_t *abund);

/* co2mco.c:25-77 for ONE component: combco codes + size_t[n+1] index -> postings (gid ascending
 * per code, codes ascending) and the DENSE inclusive prefix table (16^component_sz entries).
 * dense may be NULL (then only postings + sparse CSR are produced by the Python side). */
void orc_combco2mco(const uint32_t *combco, const uint64_t *cbdcoindex, int cofnum, int component_sz,
                    uint64_t *dense_incl, uint32_t *mco);

/* command_dist.c:774-784 for one component with a dense inclusive table; accumulates into ct. */
void orc_dist_counts_dense(const uint32_t *qcodes, const uint64_t *qindex, int qnum,
                           const uint64_t *dense_incl, const uint32_t *mco, int refnum,
                           uint32_t *ct, int nthreads);
/* Same with a sparse CSR (sorted unique codes + exclusive offsets[nuniq+1]); used for timing. */
void orc_dist_counts_csr(const uint32_t *qcodes, const uint64_t *qindex, int qnum,
                         const uint32_t *ucodes, const uint64_t *uoff, size_t nuniq,
                         const uint32_t *mco, int refnum, uint32_t *ct, int nthreads);

/* command_dist.c:1251-1287 output_ctrl, numbers only.  out[0..8] =
 * metric, dist, pv, fdr, ci_m1, ci_m2, ci_d1, ci_d2, rs ; returns 0 if the row is suppressed
 * (dist > dthreshold), else 1. metric_kind 0 = Jaccard, 1 = Containment. */
int orc_output_ctrl(uint32_t X, uint32_t Y, uint32_t I, int metric_kind, int correction,
                    int kmerlen, int dim_reduct_len, double dthreshold, uint64_t cmprsn_num,
                    double out[9]);

/* kssd set, one component (command_set.c:226-293 union, :374-443 uniq union, :294-373 intersect / subtract) */
size_t orc_set_union(const uint32_t *combco, size_t n, int uniq, int code_bits, uint32_t *out);
void orc_set_operate(const uint32_t *combco, const uint64_t *index, int n_genomes, const uint32_t *pan, size_t n_pan,
                     int intersect, int code_bits, uint32_t *out, uint64_t *out_index);

#ifdef __cplusplus
}
#endif
#endif
