/*
 * kssd_b200_dist.c -- a C host for libkssd_b200.so in the reference's own language: the three stages of `kssd dist`
 * driven through the C-ABI (include/kssd_b200.h), reading and writing the reference's on-disk formats, so that the
 * directories it leaves behind are interchangeable with the reference's (SURVEY.md s8b "Files").
 *
 *   kssd_b200_dist sketch <file.shuf> <outdir> [-u | -q <Q> -n <M> | -A] <seqfile>...   run_stageI   (command_dist.c:258-380)
 *   kssd_b200_dist index  <sketchdir>                                                 run_stageII  (command_dist.c:381-417)
 *   kssd_b200_dist dist   <refdir> <qrydir> <outdir> [-M 0|1] [-O 0|1|2] [-D <d>] [-N <n>] [--correction]
 *                                                                                      mco_cbdco_nobin_dist + dist_print_nobin
 *
 * This is not the reference's CLI (mode inference, argp, lists and the rest stay with the reference); it is the smallest
 * C program that exercises every seam the way run_stageI / run_stageII / mco_cbdco_nobin_dist would after the edits in
 * INTEGRATION.md.  The `.shuf` parameters travel from stage to stage in a side file `kssd_b200.ctx` (the reference keeps
 * them in globals of one process or asks for -L again).
 *
 * Build: gcc -std=c11 -O2 -Iinclude host/kssd_b200_dist.c -Lpublic_kssd_b200 -lkssd_b200 -Wl,-rpath,'$ORIGIN/../public_kssd_b200' -o host/kssd_b200_dist
 */
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "kssd_b200.h"

#define PATHLEN 256      /* global_basic.h: names are 256-byte NUL-padded */
#define COMPONENT_SZ 7

typedef struct {         /* co_dstat_t, global_basic.h:94-103 (x86-64 layout, 32 bytes) */
    uint32_t shuf_id;
    uint8_t koc, pad[3];
    int32_t kmerlen, dim_rd_len, comp_num, infile_num;
    uint64_t all_ctx_ct;
} co_dstat_t;

typedef struct {         /* mco_dstat_t, command_dist.h:57-64 (20 bytes) */
    uint32_t shuf_id;
    int32_t kmerlen, dim_rd_len, comp_num, infile_num;
} mco_dstat_t;

typedef struct { int32_t id, k, subk, drlevel; } shuf_hdr_t;   /* dim_shuffle_stat_t, command_shuffle.h:17-28 */

static void die(const char *what)
{
    fprintf(stderr, "kssd_b200_dist: %s: %s\n", what, kssd_last_error());
    exit(1);
}

static void die_io(const char *path)
{
    fprintf(stderr, "kssd_b200_dist: %s: %s\n", path, strerror(errno));
    exit(1);
}

static void *slurp(const char *path, size_t *bytes)
{
    FILE *f = fopen(path, "rb");
    if (!f) die_io(path);
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = malloc(n > 0 ? (size_t)n : 1);
    if (!p || (n > 0 && fread(p, 1, (size_t)n, f) != (size_t)n)) die_io(path);
    fclose(f);
    *bytes = (size_t)n;
    return p;
}

static void spill(const char *dir, const char *name, int comp, const char *suffix, const void *p, size_t bytes)
{
    char path[1024];
    if (comp >= 0) snprintf(path, sizeof path, "%s/%s.%d%s", dir, name, comp, suffix);
    else snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "wb");
    if (!f || (bytes && fwrite(p, 1, bytes, f) != bytes)) die_io(path);
    fclose(f);
}

static kssd_ctx_t *ctx_from_shuf(const char *shuf_path, shuf_hdr_t *hdr)
{
    size_t n;
    uint8_t *raw = slurp(shuf_path, &n);
    if (n < sizeof *hdr) { fprintf(stderr, "kssd_b200_dist: %s is not a .shuf file\n", shuf_path); exit(1); }
    memcpy(hdr, raw, sizeof *hdr);
    if (n != sizeof *hdr + ((size_t)4 << (4 * hdr->subk))) { fprintf(stderr, "kssd_b200_dist: %s: size does not match subk\n", shuf_path); exit(1); }
    kssd_ctx_t *ctx;
    if (kssd_ctx_create(&ctx, 0, (const int32_t *)(raw + sizeof *hdr), hdr->k, hdr->subk, hdr->drlevel, COMPONENT_SZ)) die("kssd_ctx_create");
    free(raw);
    return ctx;
}

/* the .shuf the sketches were made with, remembered beside them */
static void remember_shuf(const char *dir, const char *shuf_path)
{
    spill(dir, "kssd_b200.ctx", -1, "", shuf_path, strlen(shuf_path) + 1);
}

static kssd_ctx_t *ctx_of_dir(const char *dir, shuf_hdr_t *hdr)
{
    char path[1024];
    size_t n;
    snprintf(path, sizeof path, "%s/kssd_b200.ctx", dir);
    char *shuf_path = slurp(path, &n);
    kssd_ctx_t *ctx = ctx_from_shuf(shuf_path, hdr);
    free(shuf_path);
    return ctx;
}

/* ---- Stage I ---- */
static int cmd_sketch(int argc, char **argv)
{
    if (argc < 3) return 2;
    const char *shuf_path = argv[0], *outdir = argv[1];
    kssd_sketch_opts_t o = {KSSD_MODE_FASTA, 0, 1, 0, 0, 0};
    int a = 2;
    for (; a < argc && argv[a][0] == '-'; a++) {
        if (!strcmp(argv[a], "-u")) o.mode = KSSD_MODE_FASTA_UNIQ;
        else if (!strcmp(argv[a], "-A")) o.mode = KSSD_MODE_FASTQ_ABUND;
        else if (!strcmp(argv[a], "-q") && a + 1 < argc) { o.mode = KSSD_MODE_FASTQ; o.Q = atoi(argv[++a]); }
        else if (!strcmp(argv[a], "-n") && a + 1 < argc) { o.mode = KSSD_MODE_FASTQ; o.M = atoi(argv[++a]); }
        else return 2;
    }
    const int n = argc - a;
    if (n <= 0) return 2;
    shuf_hdr_t hdr;
    kssd_ctx_t *ctx = ctx_from_shuf(shuf_path, &hdr);
    kssd_ctx_info_t info;
    kssd_ctx_info(ctx, &info);
    kssd_stage1_t *s1;
    if (kssd_stage1_files(ctx, (const char *const *)(argv + a), n, &o, 0, 0, &s1)) die("kssd_stage1_files");
    int32_t *status = malloc(sizeof(int32_t) * n);
    kssd_stage1_status(s1, status);
    for (int i = 0; i < n; i++)
        if (status[i]) {           /* where fasta2co would have err()'d */
            fprintf(stderr, "kssd_b200_dist: %s: %s\n", argv[a + i],
                    status[i] == KSSD_E_CROWD ? "the context space is too crowd, try rerun the program using a larger -k"
                                              : "can not find seqences head start from '>'");
            return 1;
        }
    mkdir(outdir, 0777);
    uint32_t *ctx_ct = calloc(n, sizeof(uint32_t));
    uint64_t all = 0;
    for (int c = 0; c < info.component_num; c++) {
        const int64_t cnt = kssd_stage1_count(s1, c);
        uint32_t *ids = malloc(cnt > 0 ? (size_t)cnt * 4 : 4);
        uint16_t *ab = malloc(cnt > 0 ? (size_t)cnt * 2 : 2);
        uint64_t *index = malloc(sizeof(uint64_t) * (n + 1));
        if (kssd_stage1_fetch(s1, c, ids, index, o.mode == KSSD_MODE_FASTQ_ABUND ? ab : NULL)) die("kssd_stage1_fetch");
        spill(outdir, "combco", c, "", ids, (size_t)cnt * 4);                                  /* command_dist.c:331-354 */
        spill(outdir, "combco.index", c, "", index, sizeof(uint64_t) * (n + 1));
        if (o.mode == KSSD_MODE_FASTQ_ABUND) spill(outdir, "combco", c, ".a", ab, (size_t)cnt * 2);
        for (int i = 0; i < n; i++) ctx_ct[i] += (uint32_t)(index[i + 1] - index[i]);
        all += (uint64_t)cnt;
        free(ids); free(ab); free(index);
    }
    /* cofiles.stat, command_dist.c:361-377 */
    const size_t bytes = sizeof(co_dstat_t) + (size_t)n * 4 + (size_t)n * PATHLEN;
    uint8_t *st = calloc(1, bytes);
    co_dstat_t h = {(uint32_t)hdr.id, o.mode == KSSD_MODE_FASTQ_ABUND, {0, 0, 0}, 2 * hdr.k, 2 * hdr.drlevel, info.component_num, n, all};
    memcpy(st, &h, sizeof h);
    memcpy(st + sizeof h, ctx_ct, (size_t)n * 4);
    for (int i = 0; i < n; i++) strncpy((char *)st + sizeof h + (size_t)n * 4 + (size_t)i * PATHLEN, argv[a + i], PATHLEN - 1);
    spill(outdir, "cofiles.stat", -1, "", st, bytes);
    remember_shuf(outdir, shuf_path);
    double rs, gs, ts;
    uint64_t by;
    int nb;
    kssd_stage1_timing(s1, &rs, &gs, &ts, &by, &nb);
    printf("sketched %d files, %llu bytes, %llu codes in %.3f s (%d batches)\n", n, (unsigned long long)by, (unsigned long long)all, ts, nb);
    kssd_stage1_free(s1);
    kssd_ctx_destroy(ctx);
    free(st); free(ctx_ct); free(status);
    return 0;
}

typedef struct { co_dstat_t h; uint32_t *ctx_ct; char *names; uint8_t *raw; } sketch_dir_t;

static sketch_dir_t read_cofiles_stat(const char *dir)
{
    char path[1024];
    size_t n;
    snprintf(path, sizeof path, "%s/cofiles.stat", dir);
    sketch_dir_t d;
    d.raw = slurp(path, &n);
    memcpy(&d.h, d.raw, sizeof d.h);
    d.ctx_ct = (uint32_t *)(d.raw + sizeof d.h);
    d.names = (char *)(d.raw + sizeof d.h + (size_t)d.h.infile_num * 4);
    return d;
}

/* ---- Stage II ---- */
static int cmd_index(int argc, char **argv)
{
    if (argc != 1) return 2;
    const char *dir = argv[0];
    shuf_hdr_t hdr;
    kssd_ctx_t *ctx = ctx_of_dir(dir, &hdr);
    sketch_dir_t sd = read_cofiles_stat(dir);
    const int n = sd.h.infile_num;
    const size_t dense_n = (size_t)1 << (4 * COMPONENT_SZ);
    uint64_t *dense = malloc(dense_n * 8);
    for (int c = 0; c < sd.h.comp_num; c++) {
        char path[1024];
        size_t cb, ib;
        snprintf(path, sizeof path, "%s/combco.%d", dir, c);
        uint32_t *codes = slurp(path, &cb);
        snprintf(path, sizeof path, "%s/combco.index.%d", dir, c);
        uint64_t *index = slurp(path, &ib);
        kssd_index_t *ix;
        if (kssd_index_build_host(ctx, codes, index, n, &ix)) die("kssd_index_build_host");       /* combco2mco, co2mco.c:25-77 */
        uint64_t nu, np;
        int ng;
        kssd_index_sizes(ix, &nu, &np, &ng);
        uint32_t *ucodes = malloc(nu ? nu * 4 : 4), *gids = malloc(np ? np * 4 : 4);
        uint64_t *uoff = malloc((nu + 1) * 8);
        if (kssd_index_fetch(ix, ucodes, uoff, gids) || kssd_index_fetch_dense(ix, dense)) die("kssd_index_fetch");
        spill(dir, "mco", c, "", gids, np * 4);                                                    /* co2mco.c:66-71 */
        spill(dir, "mco.index", c, "", dense, dense_n * 8);                                        /* co2mco.c:57-61 */
        kssd_index_free(ix);
        free(codes); free(index); free(ucodes); free(gids); free(uoff);
    }
    /* mcofiles.stat, command_dist.c:397-409 */
    const size_t bytes = sizeof(mco_dstat_t) + (size_t)n * 4 + (size_t)n * PATHLEN;
    uint8_t *st = calloc(1, bytes);
    mco_dstat_t h = {sd.h.shuf_id, sd.h.kmerlen, sd.h.dim_rd_len, sd.h.comp_num, n};
    memcpy(st, &h, sizeof h);
    memcpy(st + sizeof h, sd.ctx_ct, (size_t)n * 4);
    memcpy(st + sizeof h + (size_t)n * 4, sd.names, (size_t)n * PATHLEN);
    spill(dir, "mcofiles.stat", -1, "", st, bytes);
    printf("indexed %d sketches, %d components\n", n, sd.h.comp_num);
    kssd_ctx_destroy(ctx);
    free(st); free(dense); free(sd.raw);
    return 0;
}

/* ---- Stage III ---- */
static int cmd_dist(int argc, char **argv)
{
    if (argc < 3) return 2;
    const char *refdir = argv[0], *qrydir = argv[1], *outdir = argv[2];
    int metric = 0, outfields = 2, correction = 0, nn = 0;
    double dthr = 1.0;
    for (int a = 3; a < argc; a++) {
        if (!strcmp(argv[a], "-M") && a + 1 < argc) metric = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-O") && a + 1 < argc) outfields = atoi(argv[++a]);
        else if (!strcmp(argv[a], "-D") && a + 1 < argc) dthr = atof(argv[++a]);
        else if (!strcmp(argv[a], "-N") && a + 1 < argc) nn = atoi(argv[++a]);
        else if (!strcmp(argv[a], "--correction")) correction = 1;
        else return 2;
    }
    shuf_hdr_t hdr;
    kssd_ctx_t *ctx = ctx_of_dir(refdir, &hdr);
    char path[1024];
    size_t n;
    snprintf(path, sizeof path, "%s/mcofiles.stat", refdir);
    uint8_t *mraw = slurp(path, &n);
    mco_dstat_t mh;
    memcpy(&mh, mraw, sizeof mh);
    const uint32_t *ref_ct = (const uint32_t *)(mraw + sizeof mh);
    const char *ref_names = (const char *)(mraw + sizeof mh + (size_t)mh.infile_num * 4);
    sketch_dir_t q = read_cofiles_stat(qrydir);
    if (q.h.comp_num != mh.comp_num || q.h.shuf_id != mh.shuf_id) {                                /* command_dist.c:701-706 */
        fprintf(stderr, "kssd_b200_dist: query args not match ref args: comp_num %d vs %d, shuf_id %u vs %u\n", q.h.comp_num, mh.comp_num,
                q.h.shuf_id, mh.shuf_id);
        return 1;
    }
    kssd_dist_t *job;
    if (kssd_dist_create(ctx, q.h.infile_num, mh.infile_num, q.ctx_ct, ref_ct, &job)) die("kssd_dist_create");
    for (int c = 0; c < mh.comp_num; c++) {                                                          /* command_dist.c:753-790 */
        size_t gb, db, cb, ib;
        snprintf(path, sizeof path, "%s/mco.%d", refdir, c);
        uint32_t *gids = slurp(path, &gb);
        snprintf(path, sizeof path, "%s/mco.index.%d", refdir, c);
        uint64_t *dense = slurp(path, &db);
        snprintf(path, sizeof path, "%s/combco.%d", qrydir, c);
        uint32_t *codes = slurp(path, &cb);
        snprintf(path, sizeof path, "%s/combco.index.%d", qrydir, c);
        uint64_t *index = slurp(path, &ib);
        kssd_index_t *ix;
        if (kssd_index_from_dense_host(ctx, dense, gids, gb / 4, mh.infile_num, &ix)) die("kssd_index_from_dense_host");
        if (kssd_dist_accumulate_host(job, ix, codes, index)) die("kssd_dist_accumulate_host");
        kssd_index_free(ix);
        free(gids); free(dense); free(codes); free(index);
    }
    mkdir(outdir, 0777);
    const size_t cells = (size_t)q.h.infile_num * mh.infile_num;
    uint32_t *ct = malloc(cells * 4);
    if (kssd_dist_fetch_counts(job, ct)) die("kssd_dist_fetch_counts");
    spill(outdir, "sharedk_ct.dat", -1, "", ct, cells * 4);                                          /* command_dist.c:748 */
    kssd_stat_opts_t so = {metric, correction, mh.kmerlen, mh.dim_rd_len, dthr, nn, 0, 0};
    const int64_t nrows = kssd_dist_stats(job, &so);
    if (nrows < 0) die("kssd_dist_stats");
    const char *text;                                        /* written by the GPU into the context's pinned buffer: the rows stay on the device */
    size_t len;
    if (kssd_dist_text(job, q.names, ref_names, PATHLEN, metric, outfields, 1, &text, &len)) die("kssd_dist_text");
    spill(outdir, "distance.out", -1, "", text, len);                                                /* dist_print_nobin */
    printf("%d x %d pairs, %lld rows\n", q.h.infile_num, mh.infile_num, (long long)nrows);
    kssd_dist_free(job);
    kssd_ctx_destroy(ctx);
    free(ct); free(mraw); free(q.raw);
    return 0;
}

int main(int argc, char **argv)
{
    int rc = 2;
    if (argc >= 2) {
        if (!strcmp(argv[1], "sketch")) rc = cmd_sketch(argc - 2, argv + 2);
        else if (!strcmp(argv[1], "index")) rc = cmd_index(argc - 2, argv + 2);
        else if (!strcmp(argv[1], "dist")) rc = cmd_dist(argc - 2, argv + 2);
    }
    if (rc == 2)
        fprintf(stderr, "usage: kssd_b200_dist sketch <file.shuf> <outdir> [-u | -q <Q> -n <M> | -A] <seqfile>...\n"
                        "       kssd_b200_dist index  <sketchdir>\n"
                        "       kssd_b200_dist dist   <refdir> <qrydir> <outdir> [-M 0|1] [-O 0|1|2] [-D <d>] [-N <n>] [--correction]\n");
    return rc;
}
