"""CPU: the oracle restatement against the committed golden vectors (outputs of the UNMODIFIED reference on
the seeded inputs of tests/golden/cases.py, produced by tests/golden/make_golden.py).  This is the pin."""
import sys
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))
import cases  # noqa: E402

from public_kssd_b200 import hostfmt, synth  # noqa: E402


@pytest.fixture(scope="module")
def tables():
    return {6: synth.make_shuf_table(6, cases.SHUF_SEED_S6), 5: synth.make_shuf_table(5, cases.SHUF_SEED_S5)}


def _load(tag):
    return np.load(GOLD / f"{tag}.npz", allow_pickle=False)


@pytest.mark.parametrize("tag,k,s,L,uniq", [("fasta_l3k10", 10, 6, 3, False), ("fasta_uniq_l3k10", 10, 6, 3, True),
                                            ("fasta_l2k8", 8, 5, 2, False), ("fasta_l3k11", 11, 6, 3, False)])
def test_fasta_byte_identical_to_reference(oracle_mod, tables, tag, k, s, L, uniq):
    g = _load(tag)
    fa = cases.fasta_inputs()
    ctx = oracle_mod.Ctx(k, s, L, tables[s])
    assert ctx.component_num == int(g["comp_num"]) and 2 * k == int(g["kmerlen"]) and 2 * L == int(g["dim_rd_len"])
    for name in g["names"]:
        ids, comp = ctx.fasta(fa[str(name)], uniq=uniq)
        for c in range(ctx.component_num):
            assert np.array_equal(ids[comp == c], g[f"{name}.{c}"]), (tag, name, c)   # hash-slot order, byte for byte


@pytest.mark.parametrize("tag,k,s,L,Q,M", [("fastq_l2k8_q0n1", 8, 5, 2, 0, 1), ("fastq_l2k8_q40n2", 8, 5, 2, 40, 2),
                                           ("fastq_l2k8_q0n3", 8, 5, 2, 0, 3), ("fastq_l3k11_q0n2", 11, 6, 3, 0, 2)])
def test_fastq_byte_identical_to_reference(oracle_mod, tables, tag, k, s, L, Q, M):
    g = _load(tag)
    fq = cases.fastq_inputs()
    ctx = oracle_mod.Ctx(k, s, L, tables[s])
    for name in g["names"]:
        ids, comp = ctx.fastq(fq[str(name)], Q, M)
        for c in range(ctx.component_num):
            assert np.array_equal(ids[comp == c], g[f"{name}.{c}"]), (tag, name, c)


@pytest.mark.parametrize("tag,k,s,L", [("byread_l2k8", 8, 5, 2), ("byread_l3k11", 11, 6, 3)])
def test_byread_byte_identical_to_reference(oracle_mod, tables, tag, k, s, L):
    """--byread (reads2mco): combco.<c> in stream order with duplicates, combco.index.<c> inclusive over the records."""
    g = _load(tag)
    ctx = oracle_mod.Ctx(k, s, L, tables[s])
    for name, data in cases.byread_inputs().items():
        n_reads, out = ctx.byread(data)
        assert ctx.component_num == int(g[f"{name}.comp_num"])
        for c in range(ctx.component_num):
            assert np.array_equal(out[c][0], g[f"{name}.{c}"]), (tag, name, c)
            assert np.array_equal(out[c][1], g[f"{name}.{c}.index"]), (tag, name, c)
            assert len(out[c][1]) == n_reads + 1


@pytest.mark.parametrize("tag", ["set_l3k10", "set_l3k11"])
def test_set_operations_identical_to_reference(oracle_mod, tag):
    """kssd set -u / -q / -i / -s: pan, uniq_pan and the filtered combco + index + per-genome counts, per component."""
    g = _load(tag)
    comp = int(g["comp_num"])
    n = len(g["names"])
    ct_i, ct_s = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    for c in range(comp):
        pan = oracle_mod.set_union(g[f"in2.{c}"])
        assert np.array_equal(pan, g[f"u.{c}"])
        assert np.array_equal(oracle_mod.set_union(g[f"in.{c}"], uniq=True), g[f"q.{c}"])
        for key, inter, ct in (("i", True, ct_i), ("s", False, ct_s)):
            codes, ix = oracle_mod.set_operate(g[f"in.{c}"], g[f"in.index.{c}"], pan, inter)
            assert np.array_equal(codes, g[f"{key}.{c}"]) and np.array_equal(ix, g[f"{key}.index.{c}"])
            ct += np.diff(ix).astype(np.uint32)
    assert np.array_equal(ct_i, g["i.ctx_ct"]) and np.array_equal(ct_s, g["s.ctx_ct"])
    # the reference leaves all_ctx_ct in the header stale (command_set.c:314-315, 365-367)
    assert int(g["i.all_ctx_ct"]) == int(g["in.all_ctx_ct"]) == int(g["s.all_ctx_ct"])


@pytest.mark.parametrize("tag", ["composite_l2k8", "composite_l3k10"])
def test_composite_text_identical_to_reference(oracle_mod, tag):
    """kssd composite (get_species_abundance): shared-k-mer abundance statistics per (query, reference), text identical."""
    g = _load(tag)
    nc = int(g["comp_num"])
    rows = oracle_mod.composite([g[f"ref.{c}"] for c in range(nc)], [g[f"ref.index.{c}"] for c in range(nc)],
                                [g[f"qry.{c}"] for c in range(nc)], [g[f"qry.index.{c}"] for c in range(nc)],
                                [g[f"qry.a.{c}"] for c in range(nc)])
    ref_txt = g["stdout"].tobytes().decode()
    norm = ["\t".join([Path(f[0]).name, Path(f[1]).name] + f[2:]) for f in (ln.split("\t") for ln in ref_txt.splitlines())]
    mine = oracle_mod.composite_text(rows, [str(n) for n in g["qry_names"]], [str(n) for n in g["ref_names"]]).splitlines()
    assert mine == norm and len(mine) > 5


def test_fastq_abundance_identical_to_reference(oracle_mod, tables):
    g = _load("fastq_abund_l2k8")
    fq = cases.fastq_inputs()
    ctx = oracle_mod.Ctx(8, 5, 2, tables[5])
    for name in g["names"]:
        ids, comp, ab = ctx.fastq_abund(fq[str(name)])
        assert np.array_equal(ids, g[f"{name}.0"]) and np.array_equal(ab, g[f"{name}.0.a"])


def test_last_fastq_record_dropped_without_trailing_newline(oracle_mod, tables):
    """fastq2co drops the last record iff the file lacks a trailing newline."""
    fq = cases.fastq_inputs()
    ctx = oracle_mod.Ctx(8, 5, 2, tables[5])
    a = fq["b_cov5_nonl"]
    with_nl = np.concatenate([a, np.frombuffer(b"\n", dtype=np.uint8)])
    last_rec_start = a.tobytes().rfind(b"@r")
    without_last = a[:last_rec_start]
    i1, _ = ctx.fastq(a)
    i2, _ = ctx.fastq(without_last)
    i3, _ = ctx.fastq(with_nl)
    assert np.array_equal(np.sort(i1), np.sort(i2))
    assert set(i1.tolist()) <= set(i3.tolist())


def test_index_identical_to_reference(oracle_mod):
    g = _load("index_dist_l3k10")
    mco, dense = oracle_mod.combco2mco(g["ref_combco"], g["ref_combco_index"], dense=True)
    assert np.array_equal(mco, g["mco"])                                  # mco.0 byte for byte
    assert dense.size == 1 << 28 and dense[-1] == g["dense_last"][0]      # mco.index.0: 2 GiB, inclusive prefix
    nz = np.flatnonzero(np.diff(np.concatenate([[0], dense])))
    assert np.array_equal(nz.astype(np.uint32), g["dense_nonzero_codes"])
    assert np.array_equal(dense[nz], g["dense_values_at_nonzero"])
    uc, uo, gids = oracle_mod.csr_from_combco(g["ref_combco"], g["ref_combco_index"])
    assert np.array_equal(gids, g["mco"]) and np.array_equal(uc, g["dense_nonzero_codes"])


def test_shared_counts_identical_to_reference(oracle_mod):
    g = _load("index_dist_l3k10")
    uc, uo, gids = oracle_mod.csr_from_combco(g["ref_combco"], g["ref_combco_index"])
    ct = oracle_mod.dist_counts(g["qry_combco"], g["qry_combco_index"], uc, uo, gids, len(g["ref_names"]))
    assert np.array_equal(ct, g["sharedk_ct"])                             # sharedk_ct.dat
    _, dense = oracle_mod.combco2mco(g["ref_combco"], g["ref_combco_index"], dense=True)
    ct2 = oracle_mod.dist_counts_dense(g["qry_combco"], g["qry_combco_index"], dense, g["mco"], len(g["ref_names"]))
    assert np.array_equal(ct2, g["sharedk_ct"])


def _rows_from_oracle(oracle_mod, g, metric, correction, dthr, nmax=0):
    ct, X, Y = g["sharedk_ct"], g["ref_ctx_ct"], g["qry_ctx_ct"]
    Q, R = ct.shape
    cm = (Q * R) & 0xFFFFFFFF
    rows = []
    for q in range(Q):
        order = range(R)
        if nmax:
            best = [(0.0, -1)] * (nmax + 1)
            for r in range(R):
                m = ct[q, r] / min(X[r], Y[q]) if metric == 1 else ct[q, r] / (int(X[r]) + int(Y[q]) - int(ct[q, r]))
                i = nmax - 1
                while i >= 0 and m > best[i][0]:
                    best[i + 1] = best[i]
                    best[i] = (m, r)
                    i -= 1
            order = [r for (_, r) in best[:nmax] if r != -1]
        for r in order:
            keep, v = oracle_mod.output_ctrl(X[r], Y[q], ct[q, r], metric, correction, 20, 6, dthr, cm)
            if keep:
                rows.append((q, r, int(ct[q, r]), np.uint32(v[8]) if np.isfinite(v[8]) else 0, int(X[r]), int(Y[q]), *v[:8]))
    return np.array(rows, dtype=[("qry", "<u4"), ("ref", "<u4"), ("shared", "<u4"), ("rs_u", "<u4"), ("ref_size", "<u4"), ("qry_size", "<u4"),
                                 ("metric", "<f8"), ("dist", "<f8"), ("pvalue", "<f8"), ("fdr", "<f8"), ("ci_metric_lo", "<f8"),
                                 ("ci_metric_hi", "<f8"), ("ci_dist_lo", "<f8"), ("ci_dist_hi", "<f8")])


@pytest.mark.parametrize("tag,metric,outfields,correction,dthr,nmax", [("default", 0, 2, 0, 1.0, 0), ("M1_O1", 1, 1, 0, 1.0, 0), ("corr_O2", 0, 2, 1, 1.0, 0),
                                                                       ("N2_M1", 1, 2, 0, 1.0, 2), ("D0.1", 0, 2, 0, 0.1, 0), ("O0", 0, 0, 0, 1.0, 0)])
def test_distance_out_text_identical_to_reference(oracle_mod, tag, metric, outfields, correction, dthr, nmax):
    """output_ctrl numbers + the host formatter reproduce the reference's distance.out text exactly."""
    g = _load("index_dist_l3k10")
    ref_txt = g[f"distance_out.{tag}"].tobytes().decode()
    rows = _rows_from_oracle(oracle_mod, g, metric, correction, dthr, nmax)
    # the reference prints the path it was given; goldens were made in a scratch dir -> compare by basename
    qn = [str(n) for n in g["qry_names"]]
    rn = [str(n) for n in g["ref_names"]]
    mine = hostfmt.distance_out_header(metric, outfields) + hostfmt.format_stat_rows(rows, qn, rn, metric, outfields)

    def norm(t):
        out = []
        for ln in t.splitlines():
            f = ln.split("\t")
            if f[0] != "Qry":
                f[0] = Path(f[0]).name.rsplit(".", 1)[0]
                f[1] = Path(f[1]).name.rsplit(".", 1)[0]
            out.append("\t".join(f))
        return out
    assert norm(mine) == norm(ref_txt)
    # the library's native formatter (host-only entry point of the C-ABI) writes the same bytes
    native = hostfmt.format_distance_out(rows, qn, rn, metric, outfields, header=True, threads=3).decode()
    assert native == mine


def test_stat_files_roundtrip(tmp_path):
    g = _load("index_dist_l3k10")
    (tmp_path / "cofiles.stat").write_bytes(g["cofiles_stat"].tobytes())
    (tmp_path / "mcofiles.stat").write_bytes(g["mcofiles_stat"].tobytes())
    co = hostfmt.read_cofiles_stat(tmp_path)
    mco = hostfmt.read_mcofiles_stat(tmp_path)
    assert co["shuf_id"] == cases.SHUF_ID and co["kmerlen"] == 20 and co["dim_rd_len"] == 6 and co["comp_num"] == 1
    assert np.array_equal(co["ctx_ct"], g["ref_ctx_ct"]) and mco["names"] == co["names"]
    out = tmp_path / "w"
    out.mkdir()
    hostfmt.write_cofiles_stat(out, co["shuf_id"], co["koc"], co["kmerlen"], co["dim_rd_len"], co["comp_num"], co["ctx_ct"], co["names"])
    hostfmt.write_mcofiles_stat(out, mco["shuf_id"], mco["kmerlen"], mco["dim_rd_len"], mco["comp_num"], mco["ctx_ct"], mco["names"])
    mine, ref = bytearray((out / "cofiles.stat").read_bytes()), bytearray(g["cofiles_stat"].tobytes())
    # the reference leaves garbage in the struct padding after `bool koc` and after each name's NUL (it fwrites
    # char[256] buffers): compare sizes, the header + ctx_ct block, and the parsed names
    assert len(mine) == len(ref)
    mine[5:8] = ref[5:8] = b"\0\0\0"
    n = co["infile_num"]
    assert mine[: 32 + 4 * n] == ref[: 32 + 4 * n]
    m2, r2 = (out / "mcofiles.stat").read_bytes(), g["mcofiles_stat"].tobytes()
    assert len(m2) == len(r2) and m2[: 20 + 4 * n] == r2[: 20 + 4 * n]
    assert hostfmt.read_cofiles_stat(out)["names"] == co["names"] and hostfmt.read_mcofiles_stat(out)["names"] == mco["names"]


def test_slot_order_replay_matches_reference(oracle_mod, tables):
    """ids + first-occurrence offsets -> the reference's combco byte order (hostfmt.slot_order), with the
    occurrence order taken from an independent pure-Python scan of a small input."""
    g = _load("fasta_l2k8")
    fa = cases.fasta_inputs()
    k, s, L = 8, 5, 2
    tab = tables[5]
    data = fa["f_short_lines"].tobytes()
    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    TL, out = 2 * k, k - s
    mask = (1 << (4 * k)) - 1
    fwd = rc = run = 0
    first = {}
    hdr = False
    for pos, ch in enumerate(data):
        if hdr:
            if ch == 10:
                hdr = False
            continue
        if ch in code:
            b = code[ch]
            fwd = ((fwd << 2) | b) & mask
            rc = (rc >> 2) | ((b ^ 3) << (4 * k - 2))
            run += 1
            if run >= TL:
                u = min(fwd, rc)
                inner = (u >> (2 * out)) & ((1 << (4 * s)) - 1)
                pf = int(tab[inner])
                if pf < 4096:
                    left = u >> (2 * (k + s))
                    right = u & ((1 << (2 * out)) - 1)
                    dr = (((left << (2 * (k + s))) + (right << (4 * s))) >> (4 * L)) + pf
                    if dr != 0 and dr not in first:
                        first[dr] = pos
        elif ch in (10, 13):
            continue
        else:
            run = 0
            hdr = ch == 62
    ids = np.array(sorted(first), dtype=np.uint32)
    occ = np.array([first[int(i)] for i in ids], dtype=np.uint64)
    ctx = oracle_mod.Ctx(k, s, L, tab)
    replay = hostfmt.slot_order(ids, occ, ctx.hashsize)
    assert np.array_equal(replay, g["f_short_lines.0"])


@pytest.mark.parametrize("tag", ["setgroup_l3k10", "setgroup_l3k11"])
def test_set_grouping_oracle_matches_reference_golden(oracle_mod, tag):
    """kssd set -g: the oracle's restatement of grouping_genomes and the grouping-file replay (hostfmt.organize_taxf) against the
    files the unmodified reference wrote."""
    from public_kssd_b200 import hostfmt
    g = np.load(GOLD / f"{tag}.npz", allow_pickle=False)
    groups_all, n_lines = hostfmt.organize_taxf("\n".join(str(t) for t in g["tax"]) + "\n")
    groups = [x["gids"] for x in groups_all if x["taxid"] != 0]
    assert hostfmt.group_names(groups_all) == [str(n) for n in g["g.names"]]
    total = 0
    for c in range(int(g["comp_num"])):
        codes, ix = oracle_mod.set_group(g[f"in.{c}"], g[f"in.index.{c}"], groups)
        assert np.array_equal(codes, g[f"g.{c}"]) and np.array_equal(ix, g[f"g.index.{c}"]), (tag, c)
        total += codes.size
    assert total == int(g["g.all_ctx_ct"])


def test_combine_pans_is_concatenation():
    from public_kssd_b200 import hostfmt
    a, b = np.array([3, 9, 12], np.uint32), np.array([], np.uint32)
    codes, ix = hostfmt.combine_pans([a, b, a[:1]])
    assert codes.tolist() == [3, 9, 12, 3] and ix.tolist() == [0, 3, 3, 4]
