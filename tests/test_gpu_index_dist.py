"""GPU parity: Stage II (inverted index) and Stage III (shared counts, statistics) vs the CPU oracle."""
import numpy as np
import pytest

from public_kssd_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sketches():
    rc, ri = synth.synth_sketches(300, 400, seed=5, cluster_size=20)
    qc, qi = synth.synth_sketches(70, 400, seed=5, cluster_size=7)     # same ancestors -> shared codes
    return rc, ri, qc, qi


def test_index_matches_oracle(gpu_ctx_l3k10, oracle_mod, sketches):
    rc, ri, _, _ = sketches
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    uc, uo, gids = ix.csr()
    euc, euo, egids = oracle_mod.csr_from_combco(rc, ri)
    assert np.array_equal(uc, euc) and np.array_equal(uo, euo) and np.array_equal(gids, egids)
    # mco.<c> exactly as the reference writes it
    mco, _ = oracle_mod.combco2mco(rc, ri)
    assert np.array_equal(gids, mco)
    ix.close()


def test_index_dense_table(gpu_ctx_l3k10, oracle_mod):
    rc, ri = synth.synth_sketches(40, 300, seed=9)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    dense = ix.dense()
    mco, edense = oracle_mod.combco2mco(rc, ri, dense=True)
    assert dense.shape == edense.shape and np.array_equal(dense, edense)
    # round trip through the reference's file content
    ix2 = gpu_ctx_l3k10.index_from_dense(dense, mco, 40)
    a, b = ix.csr(), ix2.csr()
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    ix.close(); ix2.close()


def test_index_edge_cases(gpu_ctx_l3k10, oracle_mod):
    # empty genomes in the middle, code 0 and the largest 28-bit code, one genome only
    codes = np.array([0, 5, (1 << 28) - 1, 5, 7, (1 << 28) - 1, 0], dtype=np.uint32)
    index = np.array([0, 3, 3, 6, 6, 7], dtype=np.uint64)
    ix = gpu_ctx_l3k10.combco2mco(codes, index)
    uc, uo, gids = ix.csr()
    euc, euo, egids = oracle_mod.csr_from_combco(codes, index)
    assert np.array_equal(uc, euc) and np.array_equal(uo, euo) and np.array_equal(gids, egids)
    ix.close()


def test_dist_counts_match_oracle(gpu_ctx_l3k10, oracle_mod, sketches):
    from public_kssd_b200 import kssd
    rc, ri, qc, qi = sketches
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    job = kssd.DistJob(gpu_ctx_l3k10, np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32))
    job.accumulate(ix, qc, qi)
    ct = job.counts()
    euc, euo, egids = oracle_mod.csr_from_combco(rc, ri)
    exp = oracle_mod.dist_counts(qc, qi, euc, euo, egids, len(ri) - 1)
    assert np.array_equal(ct, exp)
    assert ct.sum() > 0
    # brute force on a few pairs
    for q, r in [(0, 0), (3, 2), (69, 299), (10, 21)]:
        a = qc[int(qi[q]):int(qi[q + 1])]; b = rc[int(ri[r]):int(ri[r + 1])]
        assert ct[q, r] == np.intersect1d(a, b).size
    # a second component accumulates on top (command_dist.c:769-785)
    job.accumulate(ix, qc, qi)
    assert np.array_equal(job.counts(), 2 * exp)
    job.close(); ix.close()


def test_dist_counts_wide_and_big_query(gpu_ctx_l3k10, oracle_mod):
    """More refs than one shared-memory strip holds, and a query sketch >= 65536 codes (32-bit strip path)."""
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(60_000, 24, seed=3, cluster_size=50)
    qc, qi = synth.synth_sketches(8, 24, seed=3, cluster_size=2)
    big = np.unique(np.concatenate([rc[:200_000:2], synth.synth_sketches(1, 70_000, seed=77)[0]]))
    qc2 = np.concatenate([qc, big]).astype(np.uint32)
    qi2 = np.concatenate([qi, [qi[-1] + big.size]]).astype(np.uint64)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    euc, euo, egids = oracle_mod.csr_from_combco(rc, ri)
    for codes, index in [(qc, qi), (qc2, qi2)]:
        job = kssd.DistJob(gpu_ctx_l3k10, np.diff(index).astype(np.uint32), np.diff(ri).astype(np.uint32))
        job.accumulate(ix, codes, index)
        exp = oracle_mod.dist_counts(codes, index, euc, euo, egids, len(ri) - 1, nthreads=8)
        assert np.array_equal(job.counts(), exp)
        job.close()
    ix.close()


def _check_rows(rows, ct, qsz, rsz, oracle_mod, metric, correction, dthr, kmerlen=20, dim_rd_len=6, skip_zero=0):
    Q, R = ct.shape
    cmprsn = (Q * R) & 0xFFFFFFFF
    exp = []
    for q in range(Q):
        for r in range(R):
            keep, v = oracle_mod.output_ctrl(rsz[r], qsz[q], ct[q, r], metric, correction, kmerlen, dim_rd_len, dthr, cmprsn)
            if keep and not (skip_zero and ct[q, r] == 0):
                exp.append((q, r, v))
    assert len(rows) == len(exp)
    for row, (q, r, v) in zip(rows, exp):
        assert row["qry"] == q and row["ref"] == r and row["shared"] == ct[q, r]
        assert row["ref_size"] == rsz[r] and row["qry_size"] == qsz[q]
        assert row["rs_u"] == np.uint32(v[8]) if np.isfinite(v[8]) else True
        got = [row[n] for n in ("metric", "dist", "pvalue", "fdr", "ci_metric_lo", "ci_metric_hi", "ci_dist_lo", "ci_dist_hi")]
        for g, e in zip(got, v[:8]):
            if np.isnan(e):
                assert np.isnan(g)
            elif np.isinf(e):
                assert g == e
            else:
                assert abs(g - e) <= 1e-6 * max(abs(e), 1e-300), (q, r, g, e)     # north_star: 1e-6 relative


@pytest.mark.parametrize("metric,correction,dthr,skip_zero", [(0, 0, 1.0, 0), (1, 0, 1.0, 0), (0, 1, 1.0, 0), (1, 1, 0.2, 0), (0, 0, 0.05, 1)])
def test_stats_match_output_ctrl(gpu_ctx_l3k10, oracle_mod, metric, correction, dthr, skip_zero):
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(45, 500, seed=5, cluster_size=9)
    qc, qi = synth.synth_sketches(12, 500, seed=5, cluster_size=4)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    job = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz)
    job.accumulate(ix, qc, qi)
    ct = job.counts()
    rows = job.stats(metric=metric, correction=correction, dthreshold=dthr, skip_zero=skip_zero)
    _check_rows(rows, ct, qsz, rsz, oracle_mod, metric, correction, dthr, skip_zero=skip_zero)
    job.close(); ix.close()


def test_topn_neighbors(gpu_ctx_l3k10, oracle_mod):
    """-N: best n refs per query by the raw metric, insertion semantics of command_dist.c:1212-1227."""
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(60, 300, seed=5, cluster_size=12)
    qc, qi = synth.synth_sketches(9, 300, seed=5, cluster_size=3)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    job = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz)
    job.accumulate(ix, qc, qi)
    ct = job.counts()
    for metric in (0, 1):
        N = 5
        rows = job.stats(metric=metric, n_neighbors=N)
        exp = []
        for q in range(len(qsz)):
            best = [(0.0, -1)] * (N + 1)
            for r in range(len(rsz)):
                X, Y, I = int(rsz[r]), int(qsz[q]), int(ct[q, r])
                m = I / min(X, Y) if metric == 1 else I / (X + Y - I)
                i = N - 1
                while i >= 0 and m > best[i][0]:
                    best[i + 1] = best[i]
                    best[i] = (m, r)
                    i -= 1
            exp += [(q, r) for (m, r) in best[:N] if r != -1]
        assert [(int(a["qry"]), int(a["ref"])) for a in rows] == exp
    with pytest.raises(kssd.KssdError):
        job.stats(n_neighbors=61)
    job.close(); ix.close()


@pytest.mark.parametrize("opts", [dict(skip_zero=1), dict(dthreshold=0.05), dict(metric=1, dthreshold=0.2), dict(metric=1, skip_zero=1, correction=1),
                                  dict(dthreshold=1.0), dict(n_neighbors=3), dict(dthreshold=0.3, correction=1),
                                  dict(n_neighbors=20, metric=1), dict(n_neighbors=7, dthreshold=0.1, correction=1)])
def test_sparse_job_rows_identical_to_dense(gpu_ctx_l3k10, opts):
    """kssd_dist_create_sparse: fused count + filter + listing without the Q x R matrix gives byte-identical rows; -N picks
    its rows from the cells the sparse kernel touched (a reference sharing nothing is never listed); options that print
    zero-shared cells (-D >= 1, --correction with -D) fall back to the matrix transparently."""
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(700, 400, seed=8, cluster_size=25)
    qc, qi = synth.synth_sketches(67, 400, seed=8, cluster_size=5)
    # an empty query and an empty reference in the middle
    qi = np.concatenate([qi[:10], qi[10:11], qi[10:]])
    ri = np.concatenate([ri[:100], ri[100:101], ri[100:]])
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    dense = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz)
    dense.accumulate(ix, qc, qi)
    want = dense.stats(**opts)
    sp = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz, sparse=True)
    sp.accumulate(ix, qc, qi)
    got = sp.stats(**opts)
    assert len(want) > 0 and got.tobytes() == want.tobytes()
    # distance.out written by the GPU from the rows on the device == the host formatter on the fetched rows
    from public_kssd_b200 import hostfmt
    qn, rn = [f"q{i}.fa" for i in range(len(qsz))], [f"dir/r{i}.fna" for i in range(len(rsz))]
    assert sp.distance_out(qn, rn, opts.get("metric", 0), 2) == hostfmt.format_distance_out(got, qn, rn, opts.get("metric", 0), 2)
    assert bytes(sp.distance_out_view(qn, rn, opts.get("metric", 0), 1)) == hostfmt.format_distance_out(got, qn, rn, opts.get("metric", 0), 1)
    assert np.array_equal(sp.counts(), dense.counts())          # counts on request: the matrix is built then
    dense.close(); sp.close(); ix.close()


def test_sparse_job_unpacked_table(gpu_ctx_l3k10, monkeypatch):
    """The two-array table (used when ref ids and counts do not fit one 32-bit word) gives the same rows as the packed one."""
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(900, 300, seed=9, cluster_size=30)
    qc, qi = synth.synth_sketches(40, 300, seed=9, cluster_size=4)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    out = []
    for unpacked in (False, True):
        if unpacked:
            monkeypatch.setenv("KSSD_SPARSE_UNPACKED", "1")
        sp = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz, sparse=True)
        sp.accumulate(ix, qc, qi)
        out.append(sp.stats(dthreshold=0.2).tobytes() + sp.stats(skip_zero=1).tobytes())
        sp.close()
    assert out[0] == out[1] and len(out[0]) > 0
    ix.close()


def test_sparse_job_many_refs_per_query_falls_back(gpu_ctx_l3k10):
    """A query that touches more references than the shared-memory table holds sends the job through the matrix."""
    from public_kssd_b200 import kssd
    n_ref = 9000
    base = np.arange(50, dtype=np.uint32) * 7919 + 13
    rc = np.concatenate([np.sort(np.concatenate([base[:3], np.array([100000 + g], dtype=np.uint32)])) for g in range(n_ref)])
    ri = np.arange(n_ref + 1, dtype=np.uint64) * 4
    qc = np.sort(np.concatenate([base, np.array([100000 + 5, 100000 + 77], dtype=np.uint32)]))
    qi = np.array([0, len(qc)], dtype=np.uint64)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    dense = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz)
    dense.accumulate(ix, qc, qi)
    want = dense.stats(skip_zero=1)
    sp = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz, sparse=True)
    sp.accumulate(ix, qc, qi)
    got = sp.stats(skip_zero=1)
    assert len(want) == n_ref and got.tobytes() == want.tobytes()
    dense.close(); sp.close(); ix.close()


@pytest.mark.parametrize("opts", [dict(skip_zero=1), dict(dthreshold=0.4), dict(metric=1, dthreshold=0.3), dict(n_neighbors=5)])
def test_sparse_job_heavy_queries_take_the_dense_sub_job(gpu_ctx_l3k10, opts):
    """Queries that touch more references than the shared-memory table holds are counted through a small dense sub-job
    and merged back in print order; the others stay sparse.  Rows identical to the all-dense job."""
    from public_kssd_b200 import kssd
    n_ref = 9000
    base = np.arange(50, dtype=np.uint32) * 7919 + 13
    rc = np.concatenate([np.sort(np.concatenate([base[:3], np.array([100000 + g, 200000 + g // 7], dtype=np.uint32)])) for g in range(n_ref)])
    ri = np.arange(n_ref + 1, dtype=np.uint64) * 5
    light = [np.sort(np.array([100000 + 11 * i, 100000 + 11 * i + 1, 200000 + i, 777 + i], dtype=np.uint32)) for i in range(9)]
    heavy = np.sort(np.concatenate([base, np.array([100000 + 5, 200000 + 3], dtype=np.uint32)]))
    queries = light[:3] + [heavy] + light[3:6] + [np.zeros(0, np.uint32), heavy[:40]] + light[6:]
    qc = np.concatenate(queries)
    qi = np.concatenate([[0], np.cumsum([len(q) for q in queries])]).astype(np.uint64)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    dense = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz)
    dense.accumulate(ix, qc, qi)
    want = dense.stats(**opts)
    sp = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz, sparse=True)
    sp.accumulate(ix, qc, qi)
    got = sp.stats(**opts)
    assert len(want) > 2 * n_ref - 100 or "dthreshold" in opts or "n_neighbors" in opts
    assert got.tobytes() == want.tobytes()
    got2 = sp.stats(**opts)                                   # the job is reusable
    assert got2.tobytes() == want.tobytes()
    dense.close(); sp.close(); ix.close()


def test_sparse_job_multi_component(shuf_l3k10):
    """K11: 16 components add into the same per-query table."""
    from public_kssd_b200 import kssd
    ctx = kssd.Context(11, 6, 3, shuf_l3k10)
    try:
        parts_r = [synth.synth_sketches(120, 90, seed=20 + c, cluster_size=10) for c in range(16)]
        parts_q = [synth.synth_sketches(15, 90, seed=20 + c, cluster_size=3) for c in range(16)]
        rsz = sum(np.diff(ri).astype(np.uint32) for _, ri in parts_r)
        qsz = sum(np.diff(qi).astype(np.uint32) for _, qi in parts_q)
        idx = [ctx.combco2mco(rc, ri) for rc, ri in parts_r]
        dense = kssd.DistJob(ctx, qsz, rsz)
        sp = kssd.DistJob(ctx, qsz, rsz, sparse=True)
        for ix, (qc, qi) in zip(idx, parts_q):
            dense.accumulate(ix, qc, qi)
            sp.accumulate(ix, qc, qi)
        want = dense.stats(metric=1, dthreshold=0.5)
        got = sp.stats(metric=1, dthreshold=0.5)
        assert len(want) > 0 and got.tobytes() == want.tobytes()
        dense.close(); sp.close()
        for ix in idx:
            ix.close()
    finally:
        ctx.close()


@pytest.mark.parametrize("sparse,opts", [(False, dict()), (False, dict(metric=1, dthreshold=0.3)), (True, dict(skip_zero=1)), (False, dict(n_neighbors=4)), (True, dict(n_neighbors=4))])
def test_query_batches_equal_one_job(gpu_ctx_l3k10, sparse, opts):
    """The reference's num_cof_batch loop (command_dist.c:731-734, :763-790): searching the queries a batch at a time gives the
    rows and the counts of one job, FDR column included (cmprsn_num is the whole search's)."""
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(500, 200, seed=3, cluster_size=10)
    qc, qi = synth.synth_sketches(77, 200, seed=3, cluster_size=7)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    one = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz, sparse=sparse)
    one.accumulate(ix, qc, qi)
    want = one.stats(**opts)
    want_ct = one.counts()
    got, cts = [], []
    for lo, ct, rows in kssd.batched_search(gpu_ctx_l3k10, [ix], [qc], [qi], qsz, rsz, rows_per_batch=16, sparse=sparse, fetch_counts=True, **opts):
        assert lo % 16 == 0
        got.append(rows)
        cts.append(ct)
    assert np.concatenate(got).tobytes() == want.tobytes()
    assert np.array_equal(np.concatenate(cts), want_ct)
    assert kssd.num_cof_batch(8 << 30, 100_000) == 5 * 4096       # -m 8 with 100k references: 20480 query rows per batch
    with pytest.raises(kssd.KssdError):
        kssd.num_cof_batch(1 << 20, 100_000)
    one.close(); ix.close()


@pytest.mark.parametrize("opts", [dict(skip_zero=1), dict(dthreshold=0.3), dict(metric=1, dthreshold=0.2), dict(dthreshold=1.0)])
def test_async_sparse_search_equals_sync(gpu_ctx_l3k10, opts):
    """kssd_dist_stats_async / _wait: several searches in flight on one context give the rows of the synchronous call;
    options the fast path cannot serve (zero cells print) fall back inside _wait."""
    from public_kssd_b200 import kssd
    rc, ri = synth.synth_sketches(800, 300, seed=4, cluster_size=20)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    rsz = np.diff(ri).astype(np.uint32)
    batches = [synth.synth_sketches(60, 300, seed=4, cluster_size=5, member_seed=50 + b) for b in range(4)]
    want = []
    for qc, qi in batches:
        j = kssd.DistJob(gpu_ctx_l3k10, np.diff(qi).astype(np.uint32), rsz, sparse=True)
        j.accumulate(ix, qc, qi)
        want.append(j.stats(**opts).tobytes())
        j.close()
    jobs = []
    for qc, qi in batches:
        j = kssd.DistJob(gpu_ctx_l3k10, np.diff(qi).astype(np.uint32), rsz, sparse=True)
        j.accumulate(ix, qc, qi)
        j.stats_async(**opts)
        jobs.append(j)
    got = [j.stats_wait(fetch=True).tobytes() for j in jobs]
    assert got == want and len(want[0]) > 0
    for j in jobs:
        j.close()
    ix.close()
