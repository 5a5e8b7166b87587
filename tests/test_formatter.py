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


def test_integer_formatter_matches_snprintf():
    """csrc/fmt_exact.cuh (the code the GPU text path runs) on the host against glibc's "%.6lf" / "%E": 1.5 M values of five
    families plus the special cases; it may decline a value (|x| >= 2^40 under %.6lf, a rounding tie) but never differ."""
    import ctypes as C
    from public_kssd_b200 import capi
    handed = C.c_uint64()
    assert capi.lib().kssd_format_selftest(300_000, 11, C.byref(handed)) == 0
    assert handed.value < 300_000            # the any-bit-pattern family is about half too large for %.6lf; nothing else is declined


def _random_rows(n, nq, nr, seed):
    rng = np.random.default_rng(seed)
    rows = np.zeros(n, dtype=ROW)
    rows["qry"] = np.sort(rng.integers(0, nq, n))
    rows["ref"] = rng.integers(0, nr, n)
    for f in ("shared", "rs_u", "ref_size", "qry_size"):
        rows[f] = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    for f in ROW.names[6:]:
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-12, 6, n)
        if f in ("pvalue", "fdr"):
            v = np.abs(v) * 10.0 ** rng.integers(-290, 10, n)
        special = rng.integers(0, 40, n)
        v[special == 0] = np.nan
        v[special == 1] = -np.nan
        v[special == 2] = np.inf
        v[special == 3] = -np.inf
        v[special == 4] = 0.0
        v[special == 5] = -0.0
        v[special == 6] = 0.9999995
        v[special == 7] = 5e-324
        rows[f] = v
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("metric,outfields", [(0, 2), (1, 1), (0, 0)])
def test_gpu_text_matches_host_formatter(gpu_ctx_l3k10, metric, outfields, monkeypatch):
    """kssd_format_distance_rows_gpu: the lines written by the GPU are the bytes snprintf writes -- NaN / inf / signed zero /
    subnormal values, names of 0 .. 255 bytes, a row count that is no multiple of the CTA; a value the integer formatter
    declines (|x| >= 2^40 under %.6lf) sends the call through the host formatter with the same result."""
    import ctypes as C
    from public_kssd_b200 import capi
    nq, nr, n = 37, 91, 100_003
    rows = _random_rows(n, nq, nr, 6)
    qn = [f"qry/dir/genome_{i}.fna" for i in range(nq)]
    qn[3] = ""
    rn = [f"ref_{i}" + "x" * (i * 3 % 250) for i in range(nr)]
    rn[5] = "y" * 255

    def gpu_text(r):
        text, ln = C.c_void_p(), C.c_size_t()
        capi.check(capi.lib().kssd_format_distance_rows_gpu(gpu_ctx_l3k10._h, r.ctypes.data_as(C.c_void_p), len(r), nq, nr, hostfmt._names_block(qn),
                                                            hostfmt._names_block(rn), 256, metric, outfields, 1, C.byref(text), C.byref(ln)))
        try:
            return C.string_at(text, ln.value)
        finally:
            capi.lib().kssd_host_free(text)
    want = hostfmt.format_distance_out(rows, qn, rn, metric, outfields, header=True)
    assert gpu_text(rows) == want
    assert gpu_text(rows[:0]) == hostfmt.distance_out_header(metric, outfields).encode()
    big = rows[:5000].copy()
    big["metric"][17] = 3.5e15                                   # declined by the integer formatter -> host formatter
    assert gpu_text(big) == hostfmt.format_distance_out(big, qn, rn, metric, outfields, header=True)
    monkeypatch.setenv("KSSD_TEXT_ON_HOST", "1")
    assert gpu_text(rows[:3000]) == hostfmt.format_distance_out(rows[:3000], qn, rn, metric, outfields, header=True)
    monkeypatch.delenv("KSSD_TEXT_ON_HOST")
    bad = rows[:100].copy()
    bad["ref"][7] = nr
    with pytest.raises(Exception):
        gpu_text(bad)


def test_list_file_reader(tmp_path):
    """-l <list>: one path per line, blank lines ignored."""
    lst = tmp_path / "in.list"
    lst.write_text("a/b.fna\n\n  c d/e.fa.gz  \nlast.fq")
    assert hostfmt.read_list_file(lst) == ["a/b.fna", "c d/e.fa.gz", "last.fq"]
