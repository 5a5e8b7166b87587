"""GPU parity on the shapes BASELINE.json's configs name and the corners the lazy scan (csrc/sketch_scan3.cuh) opens:
configs[3] (L4K10 -> auto subk 7, many contigs, N runs, soft-masking), configs[4] (L3K11 FASTQ -> 16-component index ->
containment search, oracle-checked through Stage II / III), a >= 10^7-pair search compared cell by cell, bytes that the
scan takes at face value and only rejects at the end, and the reference's "context space is too crowd" accounting."""
import numpy as np
import pytest

from public_kssd_b200 import synth

pytestmark = pytest.mark.gpu


def _compare_sets(sk, orc, genomes, uniq=False):
    sets = sk.genome_sets()
    for i, g in enumerate(genomes):
        ids, comp = orc.fasta(g, uniq=uniq)
        for c in range(orc.component_num):
            exp = np.sort(ids[comp == c])
            assert np.array_equal(sets[i][c], exp), f"genome {i} comp {c}: got {len(sets[i][c])} expected {len(exp)}"


# ---- configs[3]: `kssd dist -L 4 -k 10` generates subk = 7 (command_shuffle.c:154-160): 1 GiB table, 14-base inner window
@pytest.fixture(scope="module")
def l4k10():
    from oracle import oracle as O
    from public_kssd_b200 import kssd
    O.build()
    tab = synth.make_shuf_table_affine(7, 1)
    ctx = kssd.Context(10, 7, 4, tab)
    yield ctx, O.Ctx(10, 7, 4, tab)
    ctx.close()


def test_l4k10_subk7_many_contig_genome(l4k10):
    """>= 50 Mbp, ~500 contigs (a header every ~100 KB, each with a run of ACGT in it), ~1 % N in runs up to 5000,
    soft-masked blocks, 60-column lines; plus a CRLF copy of a smaller one and a 70-column messy file."""
    ctx, orc = l4k10
    genomes = [synth.contig_genome(50_000_000, 5), synth.contig_genome(3_000_000, 6, contig_len=20_000, width=61, crlf=True),
               synth.messy_fasta(2_000_000, 9, ncontigs=40, width=70)]
    sk = ctx.sketch(genomes)
    assert ctx.info.hashsize == 131071
    _compare_sets(sk, orc, genomes)
    assert len(sk.genome_sets()[0][0]) > 500


@pytest.mark.parametrize("span", [1024, 65536])
def test_l4k10_span_sizes(l4k10, span):
    ctx, orc = l4k10
    genomes = [synth.contig_genome(4_000_000, 7, contig_len=30_000), synth.to_fasta(synth.random_bases(3_000_000, 3), "one", 0)]
    sk = ctx.sketch(genomes, span_bytes=span)
    _compare_sets(sk, orc, genomes)


# ---- bytes the lazy scan carries along as fake bases / fake skips, and the ones that must send it to the exact path
def _sprinkle(text: np.ndarray, alphabet: bytes, every: int, seed: int) -> np.ndarray:
    out = text.copy()
    r = synth._stream(seed, out.size // every + 1, salt=41)
    pos = (np.arange(r.size, dtype=np.int64) * every + (r % np.uint64(every)).astype(np.int64))
    pos = pos[pos < out.size]
    al = np.frombuffer(alphabet, dtype=np.uint8)
    out[pos] = al[(r[: pos.size] >> np.uint64(24)).astype(np.int64) % al.size]
    return out


FAKE_BASES = b"RSWBDVrswbdv01234567 !\"#$%&\x00\x01\x07\x10\x81\x97\xf3@PQpq"           # bit 3 clear: read as a base
FAKE_SKIPS = b"\x08\x09\x0b\x0c\x0e\x0f"                                                  # bit 3 set, high nibble 0: read as a line end
DIRTY = b"NnYyKkMmHh>89*-.:;<=?\x1a\x8a\xff\x7f"                                          # bit 3 set, high nibble set: exact path


@pytest.mark.parametrize("name,alphabet,every", [("fake_bases", FAKE_BASES, 700), ("fake_skips", FAKE_SKIPS, 500), ("dirty", DIRTY, 5000),
                                                 ("mixed", FAKE_BASES + FAKE_SKIPS + DIRTY, 300), ("dense_fakes", FAKE_BASES + FAKE_SKIPS, 23)])
def test_lazy_validation_never_accepts_a_fake(gpu_ctx_l3k10, shuf_l3k10, oracle_mod, name, alphabet, every):
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    genomes = []
    for i, (width, crlf) in enumerate([(80, False), (60, True), (0, False)]):
        body = synth.to_fasta(synth.random_bases(1_500_000, 100 + i), f"g{i}", width, crlf=crlf)
        genomes.append(np.concatenate([body[:8], _sprinkle(body[8:], alphabet, every, 7 * i + 1)]))
    sk = gpu_ctx_l3k10.sketch(genomes, strict=False)
    _compare_sets(sk, orc, genomes)
    if every >= 300 and b">" not in alphabet:          # (a '>' swallows the rest of its line: the one-line genome ends there)
        assert all(len(s[0]) > 50 for s in sk.genome_sets())


def test_headers_with_sequence_like_text(gpu_ctx_l3k10, shuf_l3k10, oracle_mod):
    """Header lines made of ACGT, longer than one 1 KiB iteration, straddling iteration boundaries, '>' in mid-line,
    a header as the very last line; sequence right up against them."""
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    seq = synth._ACGT[synth.random_bases(400_000, 55)].tobytes()
    parts, pos = [], 0
    for i in range(60):
        hdr_len = [5, 40, 1000, 1030, 2500, 31, 32, 33][i % 8]
        parts.append(b">" + seq[pos:pos + hdr_len] + (b"\r\n" if i % 5 == 0 else b"\n"))
        pos += hdr_len
        ln = 3000 + 37 * i
        body = seq[pos:pos + ln]
        pos += ln
        if i % 7 == 3:
            body = body[:500] + b">" + body[500:900] + b"\n" + body[900:]       # a header opening in the middle of a line
        parts.append(b"\n".join(body[j:j + 70] for j in range(0, len(body), 70)) + b"\n")
    text = np.frombuffer(b"".join(parts), dtype=np.uint8)
    genomes = [text, np.concatenate([text, np.frombuffer(b">" + seq[:300] + b"\n", dtype=np.uint8)])]
    sk = gpu_ctx_l3k10.sketch(genomes)
    _compare_sets(sk, orc, genomes)
    for span in (512, 4096):
        _compare_sets(gpu_ctx_l3k10.sketch(genomes, span_bytes=span), orc, genomes)


def test_repeated_kmer_million_times(gpu_ctx_l3k10, shuf_l3k10, oracle_mod):
    """The post-pass must not walk a run serially: one sampled k-mer repeated ~10^6 times (FASTA, -u, and -A counts)."""
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    # find a sampled 20-mer, then tile it (period 20 -> the same 20 rotations over and over)
    probe = synth.to_fasta(synth.random_bases(400_000, 77), "p", 0)
    sk = gpu_ctx_l3k10.sketch([probe])
    lo = int(sk.ord[0][0])                       # byte offset of the last base of a sampled k-mer
    unit = probe[lo - 19: lo + 1]
    tiled = np.concatenate([np.frombuffer(b">rep\n", dtype=np.uint8), np.tile(unit, 1_000_000), np.frombuffer(b"\n", dtype=np.uint8),
                            synth.to_fasta(synth.random_bases(100_000, 78), "tail", 80)])
    for uniq in (False, True):
        got = gpu_ctx_l3k10.sketch([tiled], uniq=uniq)
        _compare_sets(got, orc, [tiled], uniq=uniq)
    reads = np.frombuffer(b"".join(b"@r\n" + unit.tobytes() * 7 + b"\n+\n" + b"I" * 140 + b"\n" for _ in range(20000)), dtype=np.uint8)
    ab = gpu_ctx_l3k10.sketch_fastq([reads], abundance=True)
    ids, comp, cnt = orc.fastq_abund(reads)
    order = np.argsort(ids)
    assert np.array_equal(ab.ids[0], ids[order]) and np.array_equal(ab.abund[0], cnt[order])
    assert ab.abund[0].max() == 65535


def test_crowd_counts_every_distinct_key_and_code_zero(oracle_mod):
    """`-u`: a genome whose distinct keys exceed hashlimit dies in the reference even if few of them are unique
    (iseq2comem.c:685-691); FASTA: every occurrence of code 0 takes a slot again (:258-263)."""
    from public_kssd_b200 import capi, kssd
    tab = synth.make_shuf_table(5, 2)
    k, s, L = 8, 5, 4                               # hashsize = primer[1] = 509, hashlimit = 305
    ctx = kssd.Context(k, s, L, tab)
    orc = oracle_mod.Ctx(k, s, L, tab)
    try:
        assert ctx.info.hashlimit == 305
        once = synth.random_bases(150_000, 31)
        doubled = synth.to_fasta(np.concatenate([once, once]), "twice", 80)        # every k-mer at least twice
        few = synth.to_fasta(synth.random_bases(30_000, 32), "few", 80)
        for uniq in (False, True):
            sk = ctx.sketch([doubled, few], uniq=uniq, strict=False)
            for i, g in enumerate([doubled, few]):
                crowd = False
                try:
                    orc.fasta(g, uniq=uniq)
                except Exception:
                    crowd = True
                assert (sk.status[i] == capi.E_CROWD) == crowd, (uniq, i, sk.status[i], crowd)
            assert sk.status[0] == capi.E_CROWD          # ~580 distinct keys > 305, none of them unique
    finally:
        ctx.close()


# ---- configs[4]: FASTQ reads at L3K11 (-n 2) -> 16-component reference index -> containment search
def test_l3k11_fastq_index_containment_chain(shuf_l3k10, oracle_mod):
    from public_kssd_b200 import kssd
    ctx = kssd.Context(11, 6, 3, shuf_l3k10)
    orc = oracle_mod.Ctx(11, 6, 3, shuf_l3k10)
    try:
        NC = ctx.component_num
        assert NC == 16
        genomes = [b for _, b in synth.cluster_genomes(24, 300_000, seed=11, cluster_size=4)]
        fasta = [synth.to_fasta(b, f"g{i}", 80) for i, b in enumerate(genomes)]
        refs = ctx.sketch(fasta)
        # two read sets: a mixture of three source genomes, and one source alone
        mix = np.concatenate([synth.to_fastq(genomes[s], n, 150, seed=40 + s) for s, n in [(1, 30_000), (9, 12_000), (18, 6_000)]])
        solo = synth.to_fastq(genomes[5], 25_000, 150, seed=50)
        q = ctx.sketch_fastq([mix, solo], Q=0, M=2)
        # Stage I against the oracle, component by component
        for gi, txt in enumerate([mix, solo]):
            ids, comp = orc.fastq(txt, 0, 2)
            for c in range(NC):
                assert np.array_equal(q.genome_sets()[gi][c], np.sort(ids[comp == c])), (gi, c)
        # Stage II per component against the oracle, Stage III summed over components
        rsz = sum(np.diff(refs.index[c]).astype(np.uint32) for c in range(NC))
        qsz = sum(np.diff(q.index[c]).astype(np.uint32) for c in range(NC))
        exp = np.zeros((2, 24), dtype=np.uint32)
        dense = kssd.DistJob(ctx, qsz, rsz)
        sparse = kssd.DistJob(ctx, qsz, rsz, sparse=True)
        idx = []
        for c in range(NC):
            ix = ctx.combco2mco(refs.ids[c], refs.index[c])
            idx.append(ix)
            uc, uo, gids = ix.csr()
            euc, euo, egids = oracle_mod.csr_from_combco(refs.ids[c], refs.index[c])
            assert np.array_equal(uc, euc) and np.array_equal(uo, euo) and np.array_equal(gids, egids)
            exp += oracle_mod.dist_counts(q.ids[c], q.index[c], euc, euo, egids, 24)
            dense.accumulate(ix, q.ids[c], q.index[c])
            sparse.accumulate(ix, q.ids[c], q.index[c])
        ct = dense.counts()
        assert np.array_equal(ct, exp)
        for dthr in (1.0, 0.05):
            rows = dense.stats(metric=1, dthreshold=dthr)
            want = []
            for qi in range(2):
                for r in range(24):
                    keep, v = oracle_mod.output_ctrl(rsz[r], qsz[qi], ct[qi, r], 1, 0, 22, 6, dthr, 48)
                    if keep:
                        want.append((qi, r, v))
            assert len(rows) == len(want)
            for row, (qi, r, v) in zip(rows, want):
                assert (row["qry"], row["ref"], row["shared"]) == (qi, r, ct[qi, r])
                for g, e in zip([row["metric"], row["dist"], row["pvalue"]], v[:3]):
                    assert (np.isnan(g) and np.isnan(e)) or g == e or abs(g - e) <= 1e-6 * abs(e)
            if dthr < 1:
                assert sparse.stats(metric=1, dthreshold=dthr).tobytes() == rows.tobytes()
                hit = {(int(r["qry"]), int(r["ref"])) for r in rows}
                assert {(0, 1), (0, 9), (0, 18), (1, 5)} <= hit
        dense.close(); sparse.close()
        for ix in idx:
            ix.close()
    finally:
        ctx.close()


# ---- a search of >= 10^7 pairs compared with the oracle cell by cell (dense and sparse job)
def test_ten_million_pair_search_content(gpu_ctx_l3k10, oracle_mod):
    from public_kssd_b200 import kssd
    Q, R = 2500, 4000
    rc, ri = synth.synth_sketches(R, 300, seed=12, cluster_size=20)
    qc, qi = synth.synth_sketches(Q, 300, seed=12, cluster_size=10)
    ix = gpu_ctx_l3k10.combco2mco(rc, ri)
    euc, euo, egids = oracle_mod.csr_from_combco(rc, ri)
    exp = oracle_mod.dist_counts(qc, qi, euc, euo, egids, R, nthreads=8)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    dense = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz)
    dense.accumulate(ix, qc, qi)
    assert np.array_equal(dense.counts(), exp)
    sp = kssd.DistJob(gpu_ctx_l3k10, qsz, rsz, sparse=True)
    sp.accumulate(ix, qc, qi)
    rows = sp.stats(skip_zero=1)
    nzq, nzr = np.nonzero(exp)
    assert len(rows) == nzq.size
    assert np.array_equal(rows["qry"], nzq) and np.array_equal(rows["ref"], nzr) and np.array_equal(rows["shared"], exp[nzq, nzr])
    pick = np.random.default_rng(3).choice(len(rows), 1500, replace=False)
    cm = (Q * R) & 0xFFFFFFFF
    for i in pick:
        row = rows[i]
        keep, v = oracle_mod.output_ctrl(rsz[row["ref"]], qsz[row["qry"]], row["shared"], 0, 0, 20, 6, 1.0, cm)
        assert keep
        for g, e in zip([row["metric"], row["dist"], row["pvalue"], row["fdr"]], v[:4]):
            assert g == e or abs(g - e) <= 1e-6 * abs(e), (i, g, e)
    assert sp.stats(dthreshold=0.2).tobytes() == dense.stats(dthreshold=0.2).tobytes()
    dense.close(); sp.close(); ix.close()
