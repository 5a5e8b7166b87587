"""GPU parity: Stage I (sequence -> sketch) through the C-ABI vs the CPU oracle. Bit-exact sets."""
import numpy as np
import pytest

from public_kssd_b200 import synth

pytestmark = pytest.mark.gpu


def _inputs():
    g = {}
    g["plain80"] = synth.to_fasta(synth.random_bases(300_000, 11), "plain", 80)
    g["oneline"] = synth.to_fasta(synth.random_bases(200_000, 12), "one", 0)
    g["crlf70"] = synth.to_fasta(synth.random_bases(150_000, 13), "crlf", 70, crlf=True)
    g["messy"] = synth.messy_fasta(400_000, 3)
    g["messy_crlf"] = synth.messy_fasta(300_000, 4, crlf=True, width=70)
    g["short_lines"] = synth.to_fasta(synth.random_bases(60_000, 14), "w7", 7)
    g["tiny"] = np.frombuffer(b">t\nACGTACGTACGTACGTACGTACGTACGTAC\n", dtype=np.uint8)
    g["no_header"] = synth.to_fasta(synth.random_bases(100_000, 15), "x", 60)[3:]
    g["lower"] = np.frombuffer(synth.to_fasta(synth.random_bases(120_000, 16), "low", 80).tobytes().lower(), dtype=np.uint8)
    g["allN"] = np.frombuffer(b">n\n" + b"N" * 5000 + b"\n", dtype=np.uint8)
    g["nl_only"] = np.frombuffer(b">n\n" + b"\n" * 3000, dtype=np.uint8)
    return g


def _compare(sk, ctx_or, names, inputs, uniq=False):
    sets = sk.genome_sets()
    for i, name in enumerate(names):
        ids, comp = ctx_or.fasta(inputs[name], uniq=uniq)
        for c in range(ctx_or.component_num):
            exp = np.sort(ids[comp == c])
            got = sets[i][c]
            assert np.array_equal(got, exp), f"{name} comp {c}: got {len(got)} expected {len(exp)}"


def test_fasta_parity_l3k10(gpu_ctx_l3k10, shuf_l3k10, oracle_mod):
    inputs = _inputs()
    names = list(inputs)
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    sk = gpu_ctx_l3k10.sketch([inputs[n] for n in names])
    assert (sk.status == 0).all()
    _compare(sk, orc, names, inputs)
    # sorted ascending, duplicate free
    for g in sk.genome_sets():
        assert np.all(np.diff(g[0].astype(np.int64)) > 0)


@pytest.mark.parametrize("span", [512, 1024, 4096, 65536])
def test_fasta_parity_span_sizes(gpu_ctx_l3k10, shuf_l3k10, oracle_mod, span):
    """Work-unit size must not change the result (k-mer ownership across span boundaries)."""
    inputs = _inputs()
    names = ["plain80", "oneline", "messy", "short_lines", "crlf70"]
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    sk = gpu_ctx_l3k10.sketch([inputs[n] for n in names], span_bytes=span)
    _compare(sk, orc, names, inputs)


def test_fasta_uniq_parity(gpu_ctx_l3k10, shuf_l3k10, oracle_mod):
    b = synth.random_bases(200_000, 21)
    dup = np.concatenate([b, b[50_000:150_000], synth.random_bases(1000, 22)])
    inputs = {"dup": synth.to_fasta(dup, "dup", 80), "messy": synth.messy_fasta(200_000, 5)}
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    sk = gpu_ctx_l3k10.sketch(list(inputs.values()), uniq=True)
    _compare(sk, orc, list(inputs), inputs, uniq=True)


def test_first_occurrence_order_replays_slot_order(gpu_ctx_l3k10, shuf_l3k10, oracle_mod):
    """`ord` (first-occurrence offsets) lets the host rebuild the reference's hash-slot order byte for byte."""
    from public_kssd_b200 import hostfmt
    inputs = _inputs()
    names = ["plain80", "messy"]
    orc = oracle_mod.Ctx(10, 6, 3, shuf_l3k10)
    sk = gpu_ctx_l3k10.sketch([inputs[n] for n in names])
    for i, n in enumerate(names):
        ids, comp = orc.fasta(inputs[n])
        lo, hi = int(sk.index[0][i]), int(sk.index[0][i + 1])
        replay = hostfmt.slot_order(sk.ids[0][lo:hi], sk.ord[0][lo:hi], gpu_ctx_l3k10.info.hashsize)
        assert np.array_equal(replay, ids)


def test_header_eof_and_errors(gpu_ctx_l3k10):
    from public_kssd_b200 import capi, kssd
    bad = np.frombuffer(b">ok\nACGTACGTACGTACGTACGTACGTACGT\n>trailing header without newline", dtype=np.uint8)
    sk = gpu_ctx_l3k10.sketch([bad], strict=False)
    assert sk.status[0] == capi.E_HEADER_EOF
    with pytest.raises(kssd.KssdError):
        gpu_ctx_l3k10.sketch([bad])


def test_multi_config_parity(oracle_mod):
    """Other (k, subk, drlevel): the CLI default L2K8 (subk 5), K11 with 16 components, K9."""
    from public_kssd_b200 import kssd
    tab5 = synth.make_shuf_table(5, 2)
    tab6 = synth.make_shuf_table(6, 1)
    g = [synth.messy_fasta(300_000, 7), synth.to_fasta(synth.random_bases(400_000, 8), "p", 80)]
    for (k, s, L, tab) in [(8, 5, 2, tab5), (11, 6, 3, tab6), (9, 6, 3, tab6), (12, 6, 3, tab6)]:
        orc = oracle_mod.Ctx(k, s, L, tab)
        ctx = kssd.Context(k, s, L, tab)
        try:
            sk = ctx.sketch(g, strict=False)
            assert ctx.component_num == orc.component_num
            sets = sk.genome_sets()
            for i in range(len(g)):
                ids, comp = orc.fasta(g[i])
                for c in range(orc.component_num):
                    assert np.array_equal(sets[i][c], np.sort(ids[comp == c])), (k, s, L, i, c)
        finally:
            ctx.close()


def test_genomes_at_16_byte_starts_give_the_same_sketch(gpu_ctx_l3k10):
    """The C-ABI asks for genome starts at multiples of 16 only: a genome that starts in the middle of a 32-byte lane / a 128-byte chunk
    must give the ids AND the first-occurrence offsets of the 128-aligned layout (lane offsets below the genome start are negative)."""
    from public_kssd_b200 import capi, kssd
    gens = [synth.to_fasta(synth.random_bases(40_000 + 37 * i, 500 + i), f"g{i}", 60 + i) for i in range(7)]
    gens.append(synth.messy_fasta(60_000, 9))
    want = gpu_ctx_l3k10.sketch(gens)
    buf, goff, glen = kssd.pack_genomes(gens, align=16)
    assert any(int(o) % 32 for o in goff)
    h = gpu_ctx_l3k10.sketch_raw(buf, buf.size, goff, glen)
    got = gpu_ctx_l3k10.fetch_sketch(h, len(gens), want_ord=True)
    assert np.array_equal(got.index[0], want.index[0]) and np.array_equal(got.ids[0], want.ids[0])
    assert np.array_equal(got.ord[0], want.ord[0])
    # FASTQ through the same layout (warp walk: pieces that start before the file does)
    src = synth.random_bases(30_000, 77)
    fq = [synth.to_fastq(src, 400 + 3 * i, 100 + i, seed=600 + i) for i in range(5)]
    w = gpu_ctx_l3k10.sketch_fastq(fq, Q=0, M=1)
    buf, goff, glen = kssd.pack_genomes(fq, align=16)
    h = gpu_ctx_l3k10.sketch_raw(buf, buf.size, goff, glen, mode=capi.MODE_FASTQ, Q=0, M=1)
    g = gpu_ctx_l3k10.fetch_sketch(h, len(fq), want_ord=True)
    assert np.array_equal(g.index[0], w.index[0]) and np.array_equal(g.ids[0], w.ids[0]) and np.array_equal(g.ord[0], w.ord[0])
