"""kssd_format_distance_rows (host-only C-ABI entry point) against the plain-Python statement of the reference's
printf formats (command_dist.c:1267-1285), on enough rows that the thread split is exercised."""
import numpy as np
import pytest

from public_kssd_b200 import hostfmt

ROW = np.dtype([("qry", "<u4"), ("ref", "<u4"), ("shared", "<u4"), ("rs_u", "<u4"), ("ref_size", "<u4"), ("qry_size", "<u4"),
                ("metric", "<f8"), ("dist", "<f8"), ("pvalue", "<f8"), ("fdr", "<f8"), ("ci_metric_lo", "<f8"),
                ("ci_metric_hi", "<f8"), ("ci_dist_lo", "<f8"), ("ci_dist_hi", "<f8")])


@pytest.mark.parametrize("metric,outfields,threads", [(0, 2, 4), (1, 1, 0), (0, 0, 1)])
def test_native_formatter_matches_python(metric, outfields, threads):
    rng = np.random.default_rng(5)
    n = 40_000
    rows = np.zeros(n, dtype=ROW)
    nq, nr = 37, 91
    rows["qry"] = np.sort(rng.integers(0, nq, n))
    rows["ref"] = rng.integers(0, nr, n)
    for f in ("shared", "rs_u", "ref_size", "qry_size"):
        rows[f] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    for f in ROW.names[6:]:
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-12, 6, n)
        special = rng.integers(0, 40, n)
        v[special == 0] = np.nan
        v[special == 1] = -np.nan
        v[special == 2] = np.inf
        v[special == 3] = -np.inf
        v[special == 4] = 0.0
        v[special == 5] = -0.0
        v[special == 6] = 0.9999995          # %.6lf rounding boundary
        rows[f] = v
    qn = [f"qry/dir/genome_{i}.fna" for i in range(nq)]
    rn = [f"ref_{i}" + "x" * (i % 50) for i in range(nr)]
    want = hostfmt.distance_out_header(metric, outfields) + hostfmt.format_stat_rows(rows, qn, rn, metric, outfields)
    got = hostfmt.format_distance_out(rows, qn, rn, metric, outfields, header=True, threads=threads).decode()
    assert got == want
    assert hostfmt.format_distance_out(rows[:0], qn, rn, metric, outfields, header=False) == b""


def test_list_file_reader(tmp_path):
    """-l <list>: one path per line, blank lines ignored."""
    lst = tmp_path / "in.list"
    lst.write_text("a/b.fna\n\n  c d/e.fa.gz  \nlast.fq")
    assert hostfmt.read_list_file(lst) == ["a/b.fna", "c d/e.fa.gz", "last.fq"]
