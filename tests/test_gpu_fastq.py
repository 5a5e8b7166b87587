"""GPU parity: FASTQ Stage I (fastq2co -Q/-n, mt_shortreads2koc -A) vs the oracle and the reference goldens."""
import sys
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx_l2k8(shuf_s5):
    from public_kssd_b200 import kssd
    c = kssd.Context(8, 5, 2, shuf_s5)
    yield c
    c.close()


def _edge_files():
    src = synth.random_bases(20_000, 301)
    g = dict(cases.fastq_inputs())
    g["crlf"] = np.frombuffer(synth.to_fastq(src, 300, 80, seed=302).tobytes().replace(b"\n", b"\r\n"), dtype=np.uint8)
    one = synth.to_fastq(src, 1, 120, seed=303)
    g["single_record"] = one
    g["single_no_nl"] = one[:-1]
    g["truncated_in_seq"] = synth.to_fastq(src, 5, 100, seed=304)[:-150]
    g["truncated_after_plus"] = synth.to_fastq(src, 5, 100, seed=305)[:-101]
    g["blank_tail"] = np.concatenate([synth.to_fastq(src, 20, 90, seed=306), np.frombuffer(b"\n\n", dtype=np.uint8)])
    g["lower_and_n"] = np.frombuffer(synth.to_fastq(src, 200, 100, seed=307, n_rate=0.02).tobytes().replace(b"ACG", b"acg"), dtype=np.uint8)
    return g


@pytest.mark.parametrize("Q,M", [(0, 1), (40, 2), (0, 3), (60, 1), (-128, 2)])
def test_fastq_matches_oracle(ctx_l2k8, shuf_s5, oracle_mod, Q, M):
    files = _edge_files()
    names = list(files)
    orc = oracle_mod.Ctx(8, 5, 2, shuf_s5)
    sk = ctx_l2k8.sketch_fastq([files[n] for n in names], Q=Q, M=M)
    sets = sk.genome_sets()
    for i, n in enumerate(names):
        ids, comp = orc.fastq(files[n], Q, M)
        assert np.array_equal(sets[i][0], np.sort(ids)), (n, Q, M, len(sets[i][0]), len(ids))


def test_fastq_abundance_matches_oracle(ctx_l2k8, shuf_s5, oracle_mod):
    files = _edge_files()
    names = list(files)
    orc = oracle_mod.Ctx(8, 5, 2, shuf_s5)
    sk = ctx_l2k8.sketch_fastq([files[n] for n in names], abundance=True)
    for i, n in enumerate(names):
        ids, comp, ab = orc.fastq_abund(files[n])
        o = np.argsort(ids, kind="stable")
        lo, hi = int(sk.index[0][i]), int(sk.index[0][i + 1])
        assert np.array_equal(sk.ids[0][lo:hi], ids[o]), n
        assert np.array_equal(sk.abund[0][lo:hi], ab[o]), n


@pytest.mark.parametrize("tag,k,s,L,Q,M", [("fastq_l2k8_q0n1", 8, 5, 2, 0, 1), ("fastq_l2k8_q40n2", 8, 5, 2, 40, 2),
                                           ("fastq_l2k8_q0n3", 8, 5, 2, 0, 3), ("fastq_l3k11_q0n2", 11, 6, 3, 0, 2)])
def test_fastq_matches_reference_golden(tag, k, s, L, Q, M):
    """Straight against the files the unmodified reference wrote (set equality per genome and component)."""
    from public_kssd_b200 import kssd
    g = np.load(GOLD / f"{tag}.npz")
    fq = cases.fastq_inputs()
    tab = synth.make_shuf_table(s, cases.SHUF_SEED_S6 if s == 6 else cases.SHUF_SEED_S5)
    ctx = kssd.Context(k, s, L, tab)
    try:
        names = [str(n) for n in g["names"]]
        sk = ctx.sketch_fastq([fq[n] for n in names], Q=Q, M=M)
        sets = sk.genome_sets()
        for i, n in enumerate(names):
            for c in range(ctx.component_num):
                assert np.array_equal(sets[i][c], np.sort(g[f"{n}.{c}"])), (tag, n, c)
    finally:
        ctx.close()


def test_fastq_abundance_matches_reference_golden(ctx_l2k8):
    g = np.load(GOLD / "fastq_abund_l2k8.npz")
    fq = cases.fastq_inputs()
    names = [str(n) for n in g["names"]]
    sk = ctx_l2k8.sketch_fastq([fq[n] for n in names], abundance=True)
    for i, n in enumerate(names):
        ids, ab = g[f"{n}.0"], g[f"{n}.0.a"]
        o = np.argsort(ids, kind="stable")
        lo, hi = int(sk.index[0][i]), int(sk.index[0][i + 1])
        assert np.array_equal(sk.ids[0][lo:hi], ids[o]) and np.array_equal(sk.abund[0][lo:hi], ab[o])


def test_fasta_matches_reference_golden():
    """FASTA goldens straight from the reference, all configs incl. 16 components."""
    from public_kssd_b200 import kssd
    fa = cases.fasta_inputs()
    for tag, k, s, L, uniq in [("fasta_l3k10", 10, 6, 3, False), ("fasta_uniq_l3k10", 10, 6, 3, True), ("fasta_l2k8", 8, 5, 2, False),
                               ("fasta_l3k11", 11, 6, 3, False)]:
        g = np.load(GOLD / f"{tag}.npz")
        tab = synth.make_shuf_table(s, cases.SHUF_SEED_S6 if s == 6 else cases.SHUF_SEED_S5)
        ctx = kssd.Context(k, s, L, tab)
        try:
            names = [str(n) for n in g["names"]]
            sk = ctx.sketch([fa[n] for n in names], uniq=uniq)
            sets = sk.genome_sets()
            for i, n in enumerate(names):
                for c in range(ctx.component_num):
                    assert np.array_equal(sets[i][c], np.sort(g[f"{n}.{c}"])), (tag, n, c)
        finally:
            ctx.close()


def test_index_dist_match_reference_golden(gpu_ctx_l3k10):
    """mco.0, the dense mco.index.0, sharedk_ct.dat and the distance.out text of the reference, from the GPU."""
    from public_kssd_b200 import hostfmt, kssd
    g = np.load(GOLD / "index_dist_l3k10.npz")
    ix = gpu_ctx_l3k10.combco2mco(g["ref_combco"], g["ref_combco_index"])
    uc, uo, gids = ix.csr()
    assert np.array_equal(gids, g["mco"]) and np.array_equal(uc, g["dense_nonzero_codes"])
    dense = ix.dense()
    assert np.array_equal(dense[g["dense_nonzero_codes"]], g["dense_values_at_nonzero"]) and dense[-1] == g["dense_last"][0]
    job = kssd.DistJob(gpu_ctx_l3k10, g["qry_ctx_ct"], g["ref_ctx_ct"])
    job.accumulate(ix, g["qry_combco"], g["qry_combco_index"])
    assert np.array_equal(job.counts(), g["sharedk_ct"])
    qn, rn = [str(x) for x in g["qry_names"]], [str(x) for x in g["ref_names"]]

    def norm(t):
        out = []
        for ln in t.splitlines():
            f = ln.split("\t")
            if f[0] != "Qry":
                f[0] = Path(f[0]).name.rsplit(".", 1)[0]
                f[1] = Path(f[1]).name.rsplit(".", 1)[0]
            out.append("\t".join(f))
        return out
    for tag, metric, outfields, corr, dthr, nmax in [("default", 0, 2, 0, 1.0, 0), ("M1_O1", 1, 1, 0, 1.0, 0), ("corr_O2", 0, 2, 1, 1.0, 0),
                                                     ("N2_M1", 1, 2, 0, 1.0, 2), ("D0.1", 0, 2, 0, 0.1, 0), ("O0", 0, 0, 0, 1.0, 0)]:
        rows = job.stats(metric=metric, correction=corr, dthreshold=dthr, n_neighbors=nmax)
        mine = hostfmt.distance_out_header(metric, outfields) + hostfmt.format_stat_rows(rows, qn, rn, metric, outfields)
        ref = g[f"distance_out.{tag}"].tobytes().decode()
        assert norm(mine) == norm(ref), tag
        assert job.distance_out(qn, rn, metric, outfields).decode() == mine, tag          # the same text written by the GPU
    job.close(); ix.close()


def _tiny_reads_fastq(n: int, seed: int) -> np.ndarray:
    """reads of 1 .. 6 bases with empty headers: ~3 bytes per line, so the single-pass line index outgrows its buffer
    (one entry per 8 bytes of text) and the library has to fall back to the two-pass index"""
    r = synth._stream(seed, 2 * n, salt=41)
    parts = []
    for i in range(n):
        ln = 1 + int(r[2 * i] % np.uint64(6))
        seq = bytes(b"ACGT"[int((r[2 * i + 1] >> np.uint64(2 * j)) & np.uint64(3))] for j in range(ln))
        parts.append(b"@\n" + seq + b"\n+\n" + b"I" * ln + b"\n")
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def test_fastq_line_index_paths_agree(ctx_l2k8, shuf_s5, oracle_mod, monkeypatch):
    """single-pass line index (default), the two-pass index (KSSD_FASTQ_TWO_PASS) and the overflow fallback give the oracle's sets"""
    files = _edge_files()
    src = synth.random_bases(400_000, 311)
    files["many_reads"] = synth.to_fastq(src, 12_000, 150, seed=312)               # > 100 blocks of the index: look-back across blocks
    files["tiny_reads"] = _tiny_reads_fastq(30_000, 313)                           # index overflow -> two-pass fallback for the whole call
    hb = synth.to_fastq(src, 3_000, 120, seed=314).tobytes().replace(b"<", b"\xbc")   # quality bytes with the high bit set: negative as
    files["highbit_qual"] = np.frombuffer(hb, dtype=np.uint8)                      # signed char, they fail -Q 0 (the walk must read the quality lines)
    names = list(files)
    orc = oracle_mod.Ctx(8, 5, 2, shuf_s5)
    want = [np.sort(orc.fastq(files[n], 0, 1)[0]) for n in names]
    for two_pass in (False, True):
        if two_pass:
            monkeypatch.setenv("KSSD_FASTQ_TWO_PASS", "1")
        else:
            monkeypatch.delenv("KSSD_FASTQ_TWO_PASS", raising=False)
        sk = ctx_l2k8.sketch_fastq([files[n] for n in names], Q=0, M=1)
        sets = sk.genome_sets()
        for i, n in enumerate(names):
            assert np.array_equal(sets[i][0], want[i]), (n, two_pass, len(sets[i][0]), len(want[i]))
    # without the overflowing file the single-pass index serves every file
    monkeypatch.delenv("KSSD_FASTQ_TWO_PASS", raising=False)
    keep = [n for n in names if n != "tiny_reads"]
    sk = ctx_l2k8.sketch_fastq([files[n] for n in keep], Q=0, M=1)
    sets = sk.genome_sets()
    for i, n in enumerate(keep):
        assert np.array_equal(sets[i][0], want[names.index(n)]), n
    sk = ctx_l2k8.sketch_fastq([files[n] for n in keep], abundance=True)
    for i, n in enumerate(keep):
        ids, comp, ab = orc.fastq_abund(files[n])
        o = np.argsort(ids, kind="stable")
        lo, hi = int(sk.index[0][i]), int(sk.index[0][i + 1])
        assert np.array_equal(sk.ids[0][lo:hi], ids[o]), n
        assert np.array_equal(sk.abund[0][lo:hi], ab[o]), n


def _mixed_length_fastq(n: int, seed: int) -> np.ndarray:
    """reads of 1 .. 400 bases in random order (the lanes of the warp walk get reads of different piece counts), a few N, lower case,
    headers of different lengths so that the sequence lines start at every alignment"""
    src = synth.random_bases(50_000, seed)
    r = synth._stream(seed, 3 * n, salt=43)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    parts = []
    for i in range(n):
        ln = 1 + int(r[3 * i] % np.uint64(400))
        st = int(r[3 * i + 1] % np.uint64(src.size - ln))
        seq = acgt[src[st:st + ln]].copy()
        f = int(r[3 * i + 2])
        if f & 7 == 0 and ln > 30:
            seq[ln // 2] = ord("N")
        if f & 24 == 8:
            seq[: ln // 3] |= 0x20
        parts.append(b"@r" + b"x" * ((f >> 8) % 37) + b"\n" + seq.tobytes() + b"\n+\n" + b"I" * ln + b"\n")
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


@pytest.mark.parametrize("abund", [False, True])
def test_fastq_mixed_read_lengths(ctx_l2k8, shuf_s5, oracle_mod, abund, monkeypatch):
    """warp walk (lane per read, rounds of pieces) on reads of every length and alignment, against the oracle and the thread walk"""
    files = [_mixed_length_fastq(5000, 321), _mixed_length_fastq(77, 322)[:-1]]
    orc = oracle_mod.Ctx(8, 5, 2, shuf_s5)
    res = {}
    for walk in ("warp", "thread"):
        if walk == "thread":
            monkeypatch.setenv("KSSD_FASTQ_THREAD_WALK", "1")
        else:
            monkeypatch.delenv("KSSD_FASTQ_THREAD_WALK", raising=False)
        res[walk] = ctx_l2k8.sketch_fastq(files, Q=0, M=1, abundance=abund)
    for i, f in enumerate(files):
        for walk, sk in res.items():
            lo, hi = int(sk.index[0][i]), int(sk.index[0][i + 1])
            if abund:
                ids, comp, ab = orc.fastq_abund(f)
                o = np.argsort(ids, kind="stable")
                assert np.array_equal(sk.ids[0][lo:hi], ids[o]), (walk, i)
                assert np.array_equal(sk.abund[0][lo:hi], ab[o]), (walk, i)
            else:
                ids, comp = orc.fastq(f, 0, 1)
                assert np.array_equal(sk.ids[0][lo:hi], np.sort(ids)), (walk, i, hi - lo, len(ids))
