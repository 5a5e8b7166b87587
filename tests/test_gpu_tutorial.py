"""GPU: BASELINE.json configs[0] -- the reference's README tutorial on its own test_fna fixtures (20 references,
11 queries, 5.4 Mbp each, gzipped), L3K10: sketch both sets, index, search, and compare with what the UNMODIFIED
reference wrote for the same .shuf (tests/golden/tutorial_test_fna_l3k10.npz): every sketch set, the 11 x 20
sharedk_ct.dat matrix and the distance.out text.  The fixtures travel as oracle/_ref/test_fna (data, git-ignored)."""
import gzip
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
FNA = ROOT / "oracle" / "_ref" / "test_fna"
GOLD = ROOT / "tests" / "golden" / "tutorial_test_fna_l3k10.npz"


def test_readme_tutorial_matches_reference(gpu_ctx_l3k10):
    if not FNA.is_dir():
        pytest.skip("oracle/_ref/test_fna not present (make -C oracle ref)")
    from public_kssd_b200 import hostfmt, kssd
    g = np.load(GOLD)
    ctx = gpu_ctx_l3k10
    all_r, all_q = [str(n) for n in g["ref_names"]], [str(n) for n in g["qry_names"]]
    # the box carries a subset of the 31 fixture files; the golden holds the full 11 x 20 run
    rnames = [n for n in all_r if (FNA / "seqs1" / n).exists()]
    qnames = [n for n in all_q if (FNA / "seqs2" / n).exists()]
    assert len(rnames) >= 2 and len(qnames) >= 2
    ri, qi = [all_r.index(n) for n in rnames], [all_q.index(n) for n in qnames]
    refs = [np.frombuffer(gzip.open(FNA / "seqs1" / n).read(), dtype=np.uint8) for n in rnames]
    qrys = [np.frombuffer(gzip.open(FNA / "seqs2" / n).read(), dtype=np.uint8) for n in qnames]
    rs, qs = ctx.sketch(refs), ctx.sketch(qrys)
    for i, n in enumerate(rnames):
        assert np.array_equal(rs.genome_sets()[i][0], g[f"ref.{n}"]), n
    for i, n in enumerate(qnames):
        assert np.array_equal(qs.genome_sets()[i][0], g[f"qry.{n}"]), n
    assert np.array_equal(rs.ctx_ct(), g["ref_ctx_ct"][ri]) and np.array_equal(qs.ctx_ct(), g["qry_ctx_ct"][qi])
    ix = ctx.combco2mco(rs.ids[0], rs.index[0])
    job = kssd.DistJob(ctx, qs.ctx_ct(), rs.ctx_ct())
    job.accumulate(ix, qs.ids[0], qs.index[0])
    assert np.array_equal(job.counts(), g["sharedk_ct"][np.ix_(qi, ri)])
    # FDR = p * (ref_num * qry_num) of the reference's full 11 x 20 run
    rows = job.stats(cmprsn_num=len(all_r) * len(all_q))
    mine = hostfmt.distance_out_header(0, 2) + hostfmt.format_stat_rows(rows, qnames, rnames, 0, 2)

    def norm(t):
        out = []
        for ln in t.splitlines():
            f = ln.split("\t")
            if f[0] != "Qry":
                f[0], f[1] = Path(f[0]).name, Path(f[1]).name
            out.append("\t".join(f))
        return out
    want = [ln for ln in norm(g["distance_out"].tobytes().decode())
            if ln.startswith("Qry") or (ln.split("\t")[0] in qnames and ln.split("\t")[1] in rnames)]
    # the reference lists queries / refs in its own (shuffled) order: compare as ordered by (query, ref) name
    key = lambda ln: (ln.split("\t")[0], ln.split("\t")[1])
    assert norm(mine)[0] == want[0] and sorted(norm(mine)[1:], key=key) == sorted(want[1:], key=key)
    job.close(); ix.close()
