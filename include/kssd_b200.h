/*
 * kssd_b200.h -- C-ABI of libkssd_b200.so: the B200 (sm_100a) implementation of Kssd's
 * sketch -> index -> compare hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (yhg926/public_kssd) has no FFI or plugin API: its hot path is three in-process
 * C seams driven by file-scope globals.  Each entry point below names the seam
 * (reference file:line, paths relative to the reference tree) it replaces; INTEGRATION.md shows
 * the calls a maintainer would put at those lines.
 *
 * Conventions
 *   - every function returns 0 on success or a negative KSSD_E_* code; kssd_last_error() gives
 *     the message for the calling thread.  (The reference calls err(errno, ...) and exits; the
 *     codes map 1:1 onto those exits so a host can keep that behaviour.)
 *   - "host" buffers are ordinary (ideally pinned) host memory; "dev" buffers are device memory
 *     on the context's device.  Caller owns every buffer it passes in; results live in opaque
 *     handles until copied out and freed.
 *   - one context = one device = one CUDA stream; contexts are independent (one per GPU rank).
 *   - there is NO CPU fallback: without a usable sm_100 device kssd_ctx_create fails.
 */
#ifndef KSSD_B200_H
#define KSSD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSSD_OK 0
#define KSSD_E_INVAL (-1)      /* bad argument                                                   */
#define KSSD_E_CUDA (-2)       /* CUDA runtime failure / no device                               */
#define KSSD_E_PRIMER (-3)     /* get_hashsz(): primer_ind out of range  (command_dist.c:221-232) */
#define KSSD_E_CROWD (-4)      /* "the context space is too crowd"       (iseq2comem.c:262-263)   */
#define KSSD_E_HEADER_EOF (-5) /* "can not find seqences head start from '>'" (iseq2comem.c:233)  */
#define KSSD_E_MISMATCH (-6)   /* query args not match ref args          (command_dist.c:701-706) */
#define KSSD_E_NOMEM (-7)
#define KSSD_E_NNEIGH (-8)     /* neighborN_max > NREF or > ref_num      (command_dist.c:1196-98) */
#define KSSD_E_LONGLINE (-9)   /* a FASTQ line exceeds the reference's fgets buffer (iseq2comem.c:274,553):
                                  the reference mis-frames every later record; not imitated                */

const char *kssd_last_error(void);
const char *kssd_version(void);
/* number of kernels launched by this library in the calling process since load (bench.py) */
uint64_t kssd_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------ *
 * Context: replaces read_dim_shuffle_file (command_shuffle.c:192-207), get_hashsz
 * (command_dist.c:217-236) and seq2co_global_var_initial (iseq2comem.c:54-77) -- the globals
 * `dim_shuffle`, `hashsize`, `hashlimit`, `component_num` become fields of the context.
 * shuf_table: the .shuf payload, int32[16^subk] on the host (copied; may be freed afterwards).
 * ------------------------------------------------------------------------------------------ */
typedef struct kssd_ctx kssd_ctx_t;

typedef struct kssd_ctx_info {
    int32_t k, subk, drlevel, component_sz;
    int32_t component_num, comp_code_bits;
    uint32_t dim_end;              /* max(16^(subk-drlevel), 4096)                              */
    uint32_t hashsize, hashlimit;  /* primer[...] and hashsize*0.6, as in the reference         */
    uint32_t n_sampled;            /* |{d : shuf[d] < dim_end}|                                  */
    int32_t device, sm_count;
} kssd_ctx_info_t;

int kssd_ctx_create(kssd_ctx_t **out, int device, const int32_t *shuf_table, int k, int subk,
                    int drlevel, int component_sz);
void kssd_ctx_destroy(kssd_ctx_t *ctx);
int kssd_ctx_info(const kssd_ctx_t *ctx, kssd_ctx_info_t *info);
/* the CUDA stream (cudaStream_t) all work of this context is issued on, for event timing */
void *kssd_ctx_stream(const kssd_ctx_t *ctx);
int kssd_ctx_sync(const kssd_ctx_t *ctx);

/* ------------------------------------------------------------------------------------------ *
 * Stage I -- sequence -> sketch.  Replaces, per genome, the pair
 *     co = fasta2co(seqfname, CO[tid], pipecmd)            iseq2comem.c:188   (KSSD_MODE_FASTA)
 *     co = uniq_fasta2co(...)                              iseq2comem.c:616   (KSSD_MODE_FASTA_UNIQ)
 *     co = fastq2co(..., Q, M)                             iseq2comem.c:277   (KSSD_MODE_FASTQ)
 *     co = mt_shortreads2koc(...)                          iseq2comem.c:554   (KSSD_MODE_FASTQ_ABUND)
 *     n  = wrt_co2cmpn_use_inn_subctx / write_fqco2file / write_fqkoc2files(cofname, co)
 *                                                          iseq2comem.c:525 / :499 / :435
 * as called from run_stageI (command_dist.c:277-310), for a whole batch of genomes at once.
 * Input: the decompressed text of n genomes laid out in one byte buffer; genome g occupies
 * [goff[g], goff[g]+glen[g]); goff[g] must be a multiple of 16.
 * Output (per component c, exactly the content of combco.<c> / combco.index.<c> / combco.<c>.a,
 * command_dist.c:331-354) with each genome's ids in ASCENDING order (the reference emits hash-slot
 * order; the set is identical, and `ord` lets a host replay slot order byte-for-byte).
 * ------------------------------------------------------------------------------------------ */
enum { KSSD_MODE_FASTA = 0, KSSD_MODE_FASTA_UNIQ = 1, KSSD_MODE_FASTQ = 2, KSSD_MODE_FASTQ_ABUND = 3, KSSD_MODE_BYREAD = 4 };

typedef struct kssd_sketch_opts {
    int32_t mode;        /* KSSD_MODE_*                                                          */
    int32_t Q;           /* fastq: minimum raw quality byte (-Q), iseq2comem.c:312               */
    int32_t M;           /* fastq: minimum occurrence (-n, 1..14), iseq2comem.c:336-346          */
    int32_t want_ord;    /* also return first-occurrence byte offsets (for slot-order replay)    */
    uint32_t span_bytes; /* 0 = auto; work-unit size the batch is cut into                       */
    uint32_t reserved;
} kssd_sketch_opts_t;

typedef struct kssd_sketch kssd_sketch_t; /* result handle */

/* seq on the HOST: copies to the device inside the call (end-to-end path). */
int kssd_sketch_batch_host(kssd_ctx_t *ctx, const uint8_t *seq, size_t seq_bytes,
                           const uint64_t *goff, const uint64_t *glen, int n_genomes,
                           const kssd_sketch_opts_t *opts, kssd_sketch_t **out);
/* seq already on the DEVICE (kernel-only path; goff/glen stay on the host).  seq_dev must be 32-byte aligned and
 * readable up to the next 16-byte boundary past seq_bytes (any cudaMalloc'ed buffer is). */
int kssd_sketch_batch_dev(kssd_ctx_t *ctx, const uint8_t *seq_dev, size_t seq_bytes,
                          const uint64_t *goff, const uint64_t *glen, int n_genomes,
                          const kssd_sketch_opts_t *opts, kssd_sketch_t **out);
/* total ids in component c, or negative error */
int64_t kssd_sketch_count(const kssd_sketch_t *s, int comp);
/* per-genome status: 0 ok, KSSD_E_CROWD, KSSD_E_HEADER_EOF (the reference would have exited) */
int kssd_sketch_status(const kssd_sketch_t *s, int32_t *status_out /* n_genomes */);
/* copy component c to host: ids[count], index[n_genomes+1], optional abund[count] (mode 3),
 * optional ord[count] (want_ord). NULL pointers are skipped. */
int kssd_sketch_fetch(const kssd_sketch_t *s, int comp, uint32_t *ids, uint64_t *index,
                      uint16_t *abund, uint64_t *ord);
/* KSSD_MODE_BYREAD -- replaces reads2mco (iseq2comem.c:78-186, `kssd dist --byread`): every input "genome" is a file
 * of FASTA-formatted reads; NOTHING is deduplicated and code 0 is kept.  kssd_sketch_fetch() then returns, per
 * component, every sampled k-mer id in stream order (files concatenated, index[n_files+1] delimiting them) and in `ord`
 * the record number of each id (0 = before the first '>').  kssd_sketch_read_counts(): '>' records per file (the
 * reference's readn).  kssd_sketch_fetch_read_index(): the reference's combco.index.<comp> of ONE file -- n_reads+1
 * inclusive cumulative counts, entry r = ids of records 0..r (:175-180). */
int kssd_sketch_read_counts(const kssd_sketch_t *s, uint64_t *n_reads_out /* n_genomes */);
int kssd_sketch_fetch_read_index(const kssd_sketch_t *s, int comp, int file, uint64_t *index_out /* n_reads+1 */);
/* device pointers of component c (valid until free): ids (u32), per-genome index (u64[n+1]) */
int kssd_sketch_dev_ptrs(const kssd_sketch_t *s, int comp, const uint32_t **ids_dev,
                         const uint64_t **index_dev);
/* k-mers that passed sampling before dedup (all components), and device time of the scan kernel */
int kssd_sketch_stats(const kssd_sketch_t *s, uint64_t *n_occurrences, float *scan_kernel_ms);
void kssd_sketch_free(kssd_sketch_t *s);

/* Stage I straight from files -- replaces the file loop of run_stageI (command_dist.c:277-312) together with the
 * popen("zcat -fc") decode inside fasta2co / fastq2co (iseq2comem.c:187-200, :283-290).  n_threads host threads read
 * plain files straight into pinned staging memory and inflate .gz files with zlib; batches of about batch_bytes are
 * copied and sketched on the context stream while the readers fill the second staging buffer.  The result is what
 * kssd_sketch_fetch would give for all files in input order (combco.<c>, combco.index.<c>, combco.<c>.a).
 * n_threads <= 0: all hardware threads; batch_bytes 0: 1 GiB.  Modes: every KSSD_MODE_* but BYREAD.
 * With 640 or more .gz files in the call (KSSD_GZ_GPU=1 / 0 forces / forbids it) the files are copied to the device as
 * they are and inflated THERE, one file per thread (csrc/inflate.cuh; ISIZE and CRC-32 of every member checked): the
 * compressed bytes cross PCIe and the host cores only read.  Batches are then sized by decoded bytes (16 GiB, or
 * KSSD_GZ_BATCH_BYTES) and batch_bytes is ignored.  A file that does not decode into its ISIZE bytes (several gzip
 * members, damage) sends the whole call through zlib on the host, which reports damage as before. */
typedef struct kssd_stage1 kssd_stage1_t;
int kssd_stage1_files(kssd_ctx_t *ctx, const char *const *paths, int n_files, const kssd_sketch_opts_t *opts,
                      int n_threads, size_t batch_bytes, kssd_stage1_t **out);
/* same with the reference's -P <cmd> (iseq2comem.c:195-199): every file is read from the stdout of "<cmd> <file>" */
int kssd_stage1_files_ex(kssd_ctx_t *ctx, const char *const *paths, int n_files, const kssd_sketch_opts_t *opts,
                         int n_threads, size_t batch_bytes, const char *pipecmd, kssd_stage1_t **out);
int64_t kssd_stage1_count(const kssd_stage1_t *s, int comp);
int kssd_stage1_fetch(const kssd_stage1_t *s, int comp, uint32_t *ids, uint64_t *index /* n_files+1 */, uint16_t *abund);
int kssd_stage1_status(const kssd_stage1_t *s, int32_t *status_out /* n_files */);
/* average busy seconds per reader thread, seconds inside GPU calls, wall seconds, decoded bytes, batches */
int kssd_stage1_timing(const kssd_stage1_t *s, double *read_s, double *gpu_s, double *total_s, uint64_t *bytes, int *batches);
/* whether the .gz files were inflated on the GPU, and the seconds of H2D + inflate (part of gpu_s) */
int kssd_stage1_gz_info(const kssd_stage1_t *s, int *on_gpu, double *gz_gpu_s);
/* diagnostic, host only: the GPU's gzip decoder run on the host (0, or -1 header / -2 data / -3 output full /
 * -4 truncated / -5 CRC / -6 ISIZE); the CPU test suite checks it against zlib */
int kssd_gunzip_host(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len);
void kssd_stage1_free(kssd_stage1_t *s);

/* ------------------------------------------------------------------------------------------ *
 * Stage II -- sketch -> inverted index.  Replaces combco2mco (co2mco.c:25-77) for one component:
 * input = combco.<c> (u32 codes) + combco.index.<c> (u64[n+1]); output = the postings of mco.<c>
 * (gids ascending within a code, codes ascending) as a CSR over the codes that occur, plus an
 * expander that writes the reference's dense inclusive table mco.index.<c> (16^component_sz u64).
 * ------------------------------------------------------------------------------------------ */
typedef struct kssd_index kssd_index_t;

int kssd_index_build_host(kssd_ctx_t *ctx, const uint32_t *combco, const uint64_t *cbdcoindex,
                          int n_genomes, kssd_index_t **out);
int kssd_index_build_dev(kssd_ctx_t *ctx, const uint32_t *combco_dev, const uint64_t *cbdcoindex_dev,
                         int n_genomes, uint64_t n_codes, kssd_index_t **out);
int kssd_index_sizes(const kssd_index_t *ix, uint64_t *n_unique, uint64_t *n_postings, int *n_genomes);
/* CSR to host: ucodes[n_unique], uoff[n_unique+1] (exclusive), gids[n_postings] (= mco.<c>) */
int kssd_index_fetch(const kssd_index_t *ix, uint32_t *ucodes, uint64_t *uoff, uint32_t *gids);
/* dense inclusive prefix table exactly as fwrite'n at co2mco.c:57-61; dense_out has 16^component_sz
 * entries on the host. */
int kssd_index_fetch_dense(const kssd_index_t *ix, uint64_t *dense_out);
/* load an index back from the reference's files (mco.<c> postings + dense mco.index.<c>) */
int kssd_index_from_dense_host(kssd_ctx_t *ctx, const uint64_t *dense_incl, const uint32_t *gids,
                               uint64_t n_postings, int n_genomes, kssd_index_t **out);
void kssd_index_free(kssd_index_t *ix);

/* ------------------------------------------------------------------------------------------ *
 * Stage III -- query x reference shared k-mer counts and statistics.  Replaces the hot loop of
 * mco_cbdco_nobin_dist (command_dist.c:763-790) and the numeric part of output_ctrl
 * (command_dist.c:1251-1287) / top-N selection of dist_print_nobin (command_dist.c:1212-1227).
 * ------------------------------------------------------------------------------------------ */
typedef struct kssd_dist kssd_dist_t; /* a Q x R job: accumulates over components */

/* ref_ctx_ct / qry_ctx_ct: per-genome sketch sizes (cofiles.stat / mcofiles.stat ctx_ct lists) */
int kssd_dist_create(kssd_ctx_t *ctx, int n_qry, int n_ref, const uint32_t *qry_ctx_ct,
                     const uint32_t *ref_ctx_ct, kssd_dist_t **out);
/* Same, but the Q x R uint32 count matrix lives in CALLER-owned device memory (e.g. a buffer that takes part in
 * an NCCL reduce-scatter when the reference index is sharded by code range across GPUs).  already_filled != 0:
 * the buffer already holds counts (statistics only / further components accumulate on top). */
int kssd_dist_create_ext(kssd_ctx_t *ctx, int n_qry, int n_ref, const uint32_t *qry_ctx_ct,
                         const uint32_t *ref_ctx_ct, uint32_t *ct_dev, int already_filled, kssd_dist_t **out);
/* add one component: ct[q][r] += |{codes of q} n postings| (command_dist.c:779-784).
 * Query sketch given as combco.<c>/combco.index.<c> content. */
int kssd_dist_accumulate_host(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes,
                              const uint64_t *qindex);
int kssd_dist_accumulate_dev(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes_dev,
                             const uint64_t *qindex_dev, uint64_t n_qcodes);
/* Fused count + reduction over peer memory (one process per GPU, reference index sharded by code range): every rank
 * walks the posting lists of ITS code range for ALL queries and adds straight into the count rows of the rank that
 * owns each query block -- row_blocks[o] is the device address (local, or a peer mapping opened with kssd_ipc_open)
 * of owner o's uint32[rows_per_block][n_ref] block.  The NVLink traffic is the non-zero increments only; no partial
 * matrix exists and no reduce-scatter runs.  Owners zero their block first; callers put a barrier before and after. */
int kssd_dist_accumulate_peer(kssd_ctx_t *ctx, const kssd_index_t *ref_ix, const uint32_t *qcodes_dev,
                              const uint64_t *qindex_dev, int n_qry, int n_ref, uint32_t *const *row_blocks,
                              int rows_per_block, int world);
/* plain cudaMalloc'ed (IPC-exportable, zeroed) device memory and CUDA IPC plumbing for the call above */
int kssd_dev_alloc(kssd_ctx_t *ctx, size_t bytes, void **ptr);
int kssd_dev_zero(kssd_ctx_t *ctx, void *ptr, size_t bytes);
void kssd_dev_free(kssd_ctx_t *ctx, void *ptr);
int kssd_ipc_export(kssd_ctx_t *ctx, void *ptr, uint8_t handle[64]);
int kssd_ipc_open(kssd_ctx_t *ctx, const uint8_t handle[64], void **ptr);
int kssd_ipc_close(kssd_ctx_t *ctx, void *ptr);

/* the sharedk_ct.dat matrix, uint32[Q][R] row-major (command_dist.c:708-748) */
int kssd_dist_fetch_counts(const kssd_dist_t *d, uint32_t *ct_out);
const uint32_t *kssd_dist_counts_dev(const kssd_dist_t *d);

enum { KSSD_METRIC_JACCARD = 0, KSSD_METRIC_CONTAINMENT = 1 };

typedef struct kssd_stat_opts {
    int32_t metric;      /* -M: KSSD_METRIC_*                                                    */
    int32_t correction;  /* --correction 0/1                                                     */
    int32_t kmerlen;     /* 2k   (co_dstat.kmerlen)                                              */
    int32_t dim_rd_len;  /* 2L   (co_dstat.dim_rd_len)                                           */
    double dthreshold;   /* -D: rows with dist > dthreshold are suppressed                       */
    int32_t n_neighbors; /* -N: 0 = all refs, else best N per query by raw metric                */
    int32_t skip_zero;   /* extension: 1 = also suppress rows with shared == 0                   */
    uint64_t cmprsn_num; /* 0 = (uint32)(ref_num*qry_num) as at command_dist.c:1186; else this value
                            (a rank that owns a block of query rows passes the whole job's number)  */
} kssd_stat_opts_t;

/* one output row; the doubles are exactly the values output_ctrl formats with %lf / %E */
typedef struct kssd_stat_row {
    uint32_t qry, ref;
    uint32_t shared, rs_u; /* XnY_size, (unsigned)rs                                             */
    uint32_t ref_size, qry_size;
    double metric, dist, pvalue, fdr;
    double ci_metric_lo, ci_metric_hi, ci_dist_lo, ci_dist_hi;
} kssd_stat_row_t;

/* Sparse job: the same search without the Q x R matrix, for runs that can never print a zero-shared cell (skip_zero,
 * or -D < 1 without --correction: output_ctrl gives such a cell dist = 1 > D, command_dist.c:1265-1267).  Components
 * are REGISTERED (the device pointers must stay valid until kssd_dist_stats returns; the host variant keeps its own
 * copy) and kssd_dist_stats counts, filters and lists in one kernel: shared counts live in a per-query shared-memory
 * hash table, work is proportional to the postings touched instead of Q x R.  Rows are identical to the dense job's.
 * Queries that touch more than 6144 references do not fit the shared-memory table: they are counted through a small
 * dense sub-job (n_over x R) and merged back in print order.  -N never lists a reference that shares nothing
 * (command_dist.c:1212-1227 inserts on metric > 0 only), so its best n are picked from the cells the kernel touched
 * (topn_sparse_kernel); if some query overflowed the table, -N goes through the matrix.  If the options do print zero
 * cells (-D >= 1 without skip_zero, --correction, empty sketches whose cells are NaN), most queries overflow, or
 * counts are fetched, the whole job falls back to the matrix transparently. */
int kssd_dist_create_sparse(kssd_ctx_t *ctx, int n_qry, int n_ref, const uint32_t *qry_ctx_ct,
                            const uint32_t *ref_ctx_ct, kssd_dist_t **out);
int kssd_dist_sparse_add_dev(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes_dev,
                             const uint64_t *qindex_dev, uint64_t n_qcodes);
int kssd_dist_sparse_add_host(kssd_dist_t *d, const kssd_index_t *ref_ix, const uint32_t *qcodes,
                              const uint64_t *qindex);

/* Runs the fused statistics kernel over the count matrix.  Rows come out query-major, refs
 * ascending (or best-first for -N), as dist_print_nobin writes them.  Two-call pattern:
 * kssd_dist_stats() computes on the device and returns the number of rows; kssd_dist_fetch_stats
 * copies them out. */
int64_t kssd_dist_stats(kssd_dist_t *d, const kssd_stat_opts_t *opts);
int kssd_dist_fetch_stats(const kssd_dist_t *d, kssd_stat_row_t *rows_out);
/* The same search without a host round trip in the middle (sparse jobs): kssd_dist_stats_async() queues count + list +
 * statistics and returns; kssd_dist_stats_wait() returns the number of rows (and redoes through kssd_dist_stats() whatever
 * the fast path could not finish).  Several jobs of one context may be in flight: a host that feeds query batch after query
 * batch -- the loop of mco_cbdco_nobin_dist, command_dist.c:763-790 -- keeps the GPU busy without waiting on it. */
int kssd_dist_stats_async(kssd_dist_t *d, const kssd_stat_opts_t *opts);
int64_t kssd_dist_stats_wait(kssd_dist_t *d);
void kssd_dist_free(kssd_dist_t *d);

/* distance.out text (host side, multi-threaded): the header line dist_print_nobin writes
 * (command_dist.c:1188-1195) and one line per statistics row exactly as output_ctrl prints it
 * (command_dist.c:1267-1285: "%s\t%s\t%u-%u|%u|%u\t%.6lf\t%.6lf", then "\t%E\t%E" for
 * outfields >= 1 and "\t[%.6lf,%.6lf]\t[%.6lf,%.6lf]" for outfields >= 2).  Names are the
 * name blocks of cofiles.stat / mcofiles.stat: NUL-terminated strings `name_stride` bytes apart
 * (256 in the reference).  The text is allocated by the library (free with kssd_host_free);
 * with_header != 0 prepends the header.  n_threads <= 0: all hardware threads. */
int kssd_format_distance_rows(const kssd_stat_row_t *rows, size_t n_rows, const char *qry_names, const char *ref_names,
                              size_t name_stride, int metric, int outfields, int with_header, int n_threads,
                              char **text_out, size_t *text_len);
/* The same text produced on the GPU from the rows kssd_dist_stats left on the device (the rows never travel: the host
 * receives distance.out's bytes).  "%.6lf" and "%E" are evaluated with integer arithmetic that reproduces glibc's
 * correctly rounded output (csrc/fmt_exact.cuh); a value that arithmetic cannot decide (within 2^-47 of a rounding tie,
 * |x| >= 2^40 under "%.6lf") sends the whole call through kssd_format_distance_rows instead, so the bytes are the same
 * either way.  Names as above: n_qry and n_ref records of `name_stride` bytes.  Free the text with kssd_host_free. */
int kssd_dist_format_text(kssd_dist_t *d, const char *qry_names, const char *ref_names, size_t name_stride, int metric,
                          int outfields, int with_header, char **text_out, size_t *text_len);
/* ... into pinned host memory owned by the context (one DMA at link speed, no allocation once the buffer has grown):
 * *text stays valid until the next kssd_dist_text call on this context or kssd_ctx_destroy; do not free it. */
int kssd_dist_text(kssd_dist_t *d, const char *qry_names, const char *ref_names, size_t name_stride, int metric,
                   int outfields, int with_header, const char **text, size_t *text_len);
/* ... and for rows that are on the host (they are copied to the device first): n_qry / n_ref = the number of name records */
int kssd_format_distance_rows_gpu(kssd_ctx_t *ctx, const kssd_stat_row_t *rows, size_t n_rows, int n_qry, int n_ref,
                                  const char *qry_names, const char *ref_names, size_t name_stride, int metric,
                                  int outfields, int with_header, char **text_out, size_t *text_len);
/* diagnostic, host only: the integer formatter against snprintf("%.6lf") / snprintf("%E") on 5 n + 25 values; returns
 * the number of differing strings (0 expected), *handed_back (may be NULL) = values the formatter declined. */
int64_t kssd_format_selftest(uint64_t n, uint64_t seed, uint64_t *handed_back);
void kssd_host_free(void *p);

/* ------------------------------------------------------------------------------------------ *
 * kssd set (reference command_set.c), one component per call (the caller loops over combco.<c>):
 *   union      -u  sketch_union      (:226-293)  pan.<c>      = every code that occurs, ascending
 *   uniq union -q  uniq_sketch_union (:374-443)  uniq_pan.<c> = codes that occur exactly once in combco.<c>
 *   operate    -i / -s <pan>  sketch_operate (:294-373): every genome keeps, in order, the codes whose membership in
 *              the pan sketch equals `intersect`; index_out is the rebuilt combco.index.<c>, the per-genome counts
 *              (cofiles.stat ctx_ct) are its differences summed over the components.
 * pan_out / combco_out need room for n_codes entries.
 * ------------------------------------------------------------------------------------------ */
int kssd_set_union_host(kssd_ctx_t *ctx, const uint32_t *combco, uint64_t n_codes, int uniq, uint32_t *pan_out, uint64_t *n_out);
int kssd_set_union_dev(kssd_ctx_t *ctx, const uint32_t *combco_dev, uint64_t n_codes, int uniq, uint32_t *pan_dev,
                       uint64_t pan_cap, uint64_t *n_out);
/* kssd set -g <grouping file> (grouping_genomes, command_set.c:698-790): the pan sketch of every group of genomes.
 * member_gids[group_index[g] .. group_index[g+1]) are the genomes of group g in the order the reference walks them (hostfmt.organize_taxf
 * replays its taxon table).  Out: per group the DISTINCT codes of its members in the order of their first occurrence -- what the slot
 * layout of the reference's per-group hash table depends on (hostfmt.group_slot_order turns it into the bytes of combco.<c>) -- and
 * index_out[n_groups + 1].  codes_out holds up to the sum of the members' sketch sizes. */
int kssd_set_group_host(kssd_ctx_t *ctx, const uint32_t *combco, const uint64_t *index, int n_genomes, const uint32_t *member_gids,
                        const uint64_t *group_index, int n_groups, uint32_t *codes_out, uint64_t *index_out);
int kssd_set_operate_host(kssd_ctx_t *ctx, const uint32_t *combco, const uint64_t *index, int n_genomes,
                          const uint32_t *pan, uint64_t n_pan, int intersect, uint32_t *combco_out, uint64_t *index_out);
int kssd_set_operate_dev(kssd_ctx_t *ctx, const uint32_t *combco_dev, const uint64_t *index_dev, int n_genomes,
                         uint64_t n_codes, const uint32_t *pan_dev, uint64_t n_pan, int intersect,
                         uint32_t *combco_out_dev, uint64_t *index_out_dev);

/* ------------------------------------------------------------------------------------------ *
 * kssd composite -- replaces get_species_abundance (command_composite.c:389-547): for every query sketch with
 * abundances (`-A`: combco.<c> + combco.<c>.a) and every reference, the abundances of the k-mers they share, summarised
 * as the reference prints them (:531): kmer_num, mean = (float)sum/kmer_num, pct = mean of the sorted abundances
 * a[floor(0.98 k)] .. a[n <= 0.99 k] (1-based), median = a[k/2], max = a[k].  Rows come in print order: queries
 * ascending, references with >= min_kmers (MIN_KM_S = 6 when <= 0) shared k-mers, most shared first, ties in reference
 * order.  The reference side is the inverted index of the reference sketches, one per component (kssd_index_build_*).
 * Rows are allocated by the library (kssd_host_free). */
typedef struct kssd_comp_row {
    uint32_t qry, ref, kmer_num, median, max;
    float mean, pct;
    uint32_t reserved;
} kssd_comp_row_t;
int kssd_composite_host(kssd_ctx_t *ctx, int n_comp, const kssd_index_t *const *ref_ix, const uint32_t *const *qcodes,
                        const uint64_t *const *qindex, const uint16_t *const *qabund, int n_qry, int min_kmers,
                        kssd_comp_row_t **rows_out, uint64_t *n_rows);

/* device time (ms, CUDA events on the context stream) of the last scan / index / count / stats
 * kernel sequence issued through this context; which = 0 sketch scan, 1 sketch total,
 * 2 index build, 3 dist counts, 4 dist stats, 5 distance.out text kernels */
float kssd_ctx_last_ms(const kssd_ctx_t *ctx, int which);

#ifdef __cplusplus
}
#endif
#endif /* KSSD_B200_H */
