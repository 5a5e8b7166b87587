#!/usr/bin/env python
"""Regenerate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/kssd, built by
oracle/Makefile from /root/reference) on the seeded inputs of cases.py.  Run here (the container with
/root/reference); the outputs are committed so the pin travels to boxes without the reference.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))

from oracle import oracle as O  # noqa: E402
from public_kssd_b200 import synth  # noqa: E402
import cases  # noqa: E402

OUT = Path(__file__).resolve().parent


def by_name(d):
    """sketch dir -> {basename without extension: [per component id arrays (file order)]}, plus stat."""
    st = O.read_cofiles_stat(d)
    res = {}
    comps = [O.read_combco(d, c) for c in range(st["comp_num"])]
    for i, nm in enumerate(st["names"]):
        key = Path(nm).name.rsplit(".", 1)[0]
        res[key] = [(codes[int(ix[i]):int(ix[i + 1])], None if ab is None else ab[int(ix[i]):int(ix[i + 1])]) for codes, ix, ab in comps]
    return st, res


def sketch_case(tag, k, s, L, table, inputs, ext, extra=(), store_abund=False):
    rr = O.RefRun(k, s, L, table, shuf_id=cases.SHUF_ID)
    d = rr.dir / "in"
    d.mkdir()
    for n, b in inputs.items():
        (d / f"{n}.{ext}").write_bytes(b.tobytes())
    out = rr.sketch(d, "sk", extra=extra, p=1)
    st, res = by_name(out)
    pack = {"comp_num": np.int32(st["comp_num"]), "kmerlen": np.int32(st["kmerlen"]), "dim_rd_len": np.int32(st["dim_rd_len"]),
            "names": np.array(sorted(res))}
    for n in sorted(res):
        for c, (ids, ab) in enumerate(res[n]):
            pack[f"{n}.{c}"] = ids            # reference order (hash-slot order)
            if store_abund and ab is not None:
                pack[f"{n}.{c}.a"] = ab
    np.savez_compressed(OUT / f"{tag}.npz", **pack)
    print(tag, {n: [len(x[0]) for x in res[n]] for n in sorted(res)})
    return rr, out, st


def main():
    O.build()
    assert O.ref_available(), "reference binary missing: make -C oracle ref"
    t6 = synth.make_shuf_table(6, cases.SHUF_SEED_S6)
    t5 = synth.make_shuf_table(5, cases.SHUF_SEED_S5)
    fa, fq = cases.fasta_inputs(), cases.fastq_inputs()

    # ---- Stage I, FASTA ----
    rr, out, st = sketch_case("fasta_l3k10", 10, 6, 3, t6, fa, "fasta")
    # ---- Stage II + III on the same sketches: refs = all, queries = subset ----
    rr.index(out)
    mco = np.fromfile(out / "mco.0", dtype="<u4")
    dense = np.fromfile(out / "mco.index.0", dtype="<u8")
    nz = np.flatnonzero(np.diff(np.concatenate([[0], dense])))
    mst = O.read_mcofiles_stat(out)
    qd = rr.dir / "qin"
    qd.mkdir()
    for n in ("h_anc", "i_mut1", "j_mut5", "d_messy"):
        (qd / f"{n}.fasta").write_bytes(fa[n].tobytes())
    qout = rr.sketch(qd, "qsk", p=1)
    qst = O.read_cofiles_stat(qout)
    pack = {"ref_names": np.array([Path(n).name.rsplit(".", 1)[0] for n in mst["names"]]), "ref_ctx_ct": mst["ctx_ct"],
            "qry_names": np.array([Path(n).name.rsplit(".", 1)[0] for n in qst["names"]]), "qry_ctx_ct": qst["ctx_ct"],
            "mco": mco, "dense_nonzero_codes": nz.astype(np.uint32), "dense_values_at_nonzero": dense[nz], "dense_last": dense[-1:],
            "mcofiles_stat": np.frombuffer((out / "mcofiles.stat").read_bytes(), dtype=np.uint8),
            "cofiles_stat": np.frombuffer((out / "cofiles.stat").read_bytes(), dtype=np.uint8),
            "ref_combco": np.fromfile(out / "combco.0", dtype="<u4"), "ref_combco_index": np.fromfile(out / "combco.index.0", dtype="<u8"),
            "qry_combco": np.fromfile(qout / "combco.0", dtype="<u4"), "qry_combco_index": np.fromfile(qout / "combco.index.0", dtype="<u8")}
    variants = {"default": [], "M1_O1": ["-M", "1", "-O", "1"], "corr_O2": ["--correction", "1"], "N2_M1": ["-N", "2", "-M", "1"],
                "D0.1": ["-D", "0.1"], "O0": ["-O", "0"]}
    for tag, extra in variants.items():
        dout = rr.dist(out, qout, f"dist_{tag}", extra=["--keepskf"] + extra if tag == "default" else extra)
        pack[f"distance_out.{tag}"] = np.frombuffer((dout / "distance.out").read_bytes(), dtype=np.uint8)
        if tag == "default":
            pack["sharedk_ct"] = np.fromfile(dout / "sharedk_ct.dat", dtype="<u4").reshape(len(qst["names"]), len(mst["names"]))
    np.savez_compressed(OUT / "index_dist_l3k10.npz", **pack)
    print("index/dist: postings", mco.size, "unique", nz.size, "ct\n", pack["sharedk_ct"])
    rr.cleanup()

    # ---- Stage I variants ----
    rr, _, _ = sketch_case("fasta_uniq_l3k10", 10, 6, 3, t6, {n: fa[n] for n in ("g_dup", "d_messy", "a_plain80")}, "fasta", extra=["-u"])
    rr.cleanup()
    rr, _, _ = sketch_case("fasta_l2k8", 8, 5, 2, t5, {n: fa[n] for n in ("a_plain80", "d_messy", "f_short_lines")}, "fasta")
    rr.cleanup()
    rr, _, _ = sketch_case("fasta_l3k11", 11, 6, 3, t6, {n: fa[n] for n in ("a_plain80", "d_messy", "h_anc")}, "fasta")
    rr.cleanup()
    for tag, extra in {"fastq_l2k8_q0n1": [], "fastq_l2k8_q40n2": ["-Q", "40", "-n", "2"], "fastq_l2k8_q0n3": ["-n", "3"]}.items():
        rr, _, _ = sketch_case(tag, 8, 5, 2, t5, fq, "fastq", extra=extra)
        rr.cleanup()
    rr, _, _ = sketch_case("fastq_l3k11_q0n2", 11, 6, 3, t6, fq, "fastq", extra=["-n", "2"])
    rr.cleanup()
    rr, _, _ = sketch_case("fastq_abund_l2k8", 8, 5, 2, t5, fq, "fastq", extra=["-A"], store_abund=True)
    rr.cleanup()


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def tutorial_golden():
    """BASELINE.json configs[0]: the README quick tutorial on test_fna (seqs1 = refs, seqs2 = queries), L3K10."""
    import gzip
    t6 = synth.make_shuf_table(6, cases.SHUF_SEED_S6)
    rr = O.RefRun(10, 6, 3, t6, shuf_id=cases.SHUF_ID)
    src = Path("/root/reference/test_fna")
    ref = rr.sketch(src / "seqs1", "reference", p=4)
    rr.index(ref)
    qry = rr.sketch(src / "seqs2", "query", p=4)
    out = rr.dist(ref, qry, "distout", extra=["--keepskf"])
    rst, qst = O.read_mcofiles_stat(ref), O.read_cofiles_stat(qry)
    rc, ri, _ = O.read_combco(ref, 0)
    qc, qi, _ = O.read_combco(qry, 0)
    pack = {"ref_names": np.array([Path(n).name for n in rst["names"]]), "qry_names": np.array([Path(n).name for n in qst["names"]]),
            "ref_ctx_ct": rst["ctx_ct"], "qry_ctx_ct": qst["ctx_ct"],
            "sharedk_ct": np.fromfile(out / "sharedk_ct.dat", dtype="<u4").reshape(len(qst["names"]), len(rst["names"])),
            "distance_out": np.frombuffer((out / "distance.out").read_bytes(), dtype=np.uint8)}
    for i, n in enumerate(pack["ref_names"]):
        pack[f"ref.{n}"] = np.sort(rc[int(ri[i]):int(ri[i + 1])])
    for i, n in enumerate(pack["qry_names"]):
        pack[f"qry.{n}"] = np.sort(qc[int(qi[i]):int(qi[i + 1])])
    np.savez_compressed(OUT / "tutorial_test_fna_l3k10.npz", **pack)
    print("tutorial:", len(rc), "ref codes", len(qc), "query codes; ct sum", int(pack["sharedk_ct"].sum()))
    rr.cleanup()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "tutorial":
    tutorial_golden()


def byread_golden():
    """`kssd dist --byread` (reads2mco): one run per input file (each run rewrites combco.* of the output directory)."""
    O.build()
    t6 = synth.make_shuf_table(6, cases.SHUF_SEED_S6)
    t5 = synth.make_shuf_table(5, cases.SHUF_SEED_S5)
    for tag, (k, s, L, tab) in {"byread_l2k8": (8, 5, 2, t5), "byread_l3k11": (11, 6, 3, t6)}.items():
        pack = {}
        for n, b in cases.byread_inputs().items():
            rr = O.RefRun(k, s, L, tab, shuf_id=cases.SHUF_ID)
            d = rr.dir / "in"
            d.mkdir()
            (d / f"{n}.fasta").write_bytes(b.tobytes())
            out = rr.sketch(d, "sk", extra=["--byread"], p=1)
            comp = 0
            while (out / f"combco.{comp}").exists():
                pack[f"{n}.{comp}"] = np.fromfile(out / f"combco.{comp}", dtype="<u4")
                pack[f"{n}.{comp}.index"] = np.fromfile(out / f"combco.index.{comp}", dtype="<u8")
                comp += 1
            pack[f"{n}.comp_num"] = np.int32(comp)
            print(tag, n, "components", comp, "reads", len(pack[f"{n}.0.index"]) - 1, "codes", sum(len(pack[f"{n}.{c}"]) for c in range(comp)))
            rr.cleanup()
        np.savez_compressed(OUT / f"{tag}.npz", **pack)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "byread":
    byread_golden()


def set_golden():
    """`kssd set` on a Stage-I directory: -u (pan), -q (uniq_pan), -i / -s <pan> (filters), L3K10 and L3K11 (16 components)."""
    O.build()
    t6 = synth.make_shuf_table(6, cases.SHUF_SEED_S6)
    fa = cases.fasta_inputs()
    for tag, (k, s, L) in {"set_l3k10": (10, 6, 3), "set_l3k11": (11, 6, 3)}.items():
        rr = O.RefRun(k, s, L, t6, shuf_id=cases.SHUF_ID)
        d = rr.dir / "in"
        d.mkdir()
        for n in ("g_dup", "h_anc", "i_mut1", "j_mut5", "a_plain80"):
            (d / f"{n}.fasta").write_bytes(fa[n].tobytes())
        sk = rr.sketch(d, "sk", p=1)
        # the pan sketch is built from a SUBSET (another directory), so that intersect / subtract are both non-trivial
        d2 = rr.dir / "in2"
        d2.mkdir()
        for n in ("h_anc", "g_dup"):
            (d2 / f"{n}.fasta").write_bytes(fa[n].tobytes())
        sk2 = rr.sketch(d2, "sk2", p=1)
        st = O.read_cofiles_stat(sk)
        comp = st["comp_num"]
        pack = {"comp_num": np.int32(comp), "names": np.array([Path(n).name.rsplit(".", 1)[0] for n in st["names"]])}
        for c in range(comp):
            codes, ix, _ = O.read_combco(sk, c)
            pack[f"in.{c}"] = codes
            pack[f"in.index.{c}"] = ix
            c2, _, _ = O.read_combco(sk2, c)
            pack[f"in2.{c}"] = c2
        runs = {"u": (["-u"], sk2, "pan"), "q": (["-q"], sk, "uniq_pan")}
        for key, (flags, src, prefix) in runs.items():
            out = rr.dir / f"set_{key}"
            r = O.run_ref(["set", *flags, "-o", out, src], cwd=rr.dir)
            assert r.returncode == 0, r.stderr
            for c in range(comp):
                pack[f"{key}.{c}"] = np.fromfile(out / f"{prefix}.{c}", dtype="<u4")
        for key, flag in {"i": "-i", "s": "-s"}.items():
            out = rr.dir / f"set_{key}"
            r = O.run_ref(["set", flag, rr.dir / "set_u", "-o", out, sk], cwd=rr.dir)
            assert r.returncode == 0, r.stderr
            for c in range(comp):
                pack[f"{key}.{c}"] = np.fromfile(out / f"combco.{c}", dtype="<u4")
                pack[f"{key}.index.{c}"] = np.fromfile(out / f"combco.index.{c}", dtype="<u8")
            pack[f"{key}.ctx_ct"] = O.read_cofiles_stat(out)["ctx_ct"]
            pack[f"{key}.all_ctx_ct"] = np.uint64(O.read_cofiles_stat(out)["all_ctx_ct"])
        pack["in.all_ctx_ct"] = np.uint64(st["all_ctx_ct"])
        np.savez_compressed(OUT / f"{tag}.npz", **pack)
        print(tag, "components", comp, "input codes", sum(len(pack[f'in.{c}']) for c in range(comp)),
              "pan", sum(len(pack[f'u.{c}']) for c in range(comp)), "uniq", sum(len(pack[f'q.{c}']) for c in range(comp)),
              "intersect", sum(len(pack[f'i.{c}']) for c in range(comp)), "subtract", sum(len(pack[f's.{c}']) for c in range(comp)))
        rr.cleanup()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "set":
    set_golden()


def composite_golden():
    """`kssd composite -r <ref sketches> -q <-A query sketches>` (get_species_abundance, command_composite.c:389-547): the
    text it prints plus the sketch files it read, L2K8 and L3K10 (the reference aborts in -A sketching at K11)."""
    O.build()
    t6 = synth.make_shuf_table(6, cases.SHUF_SEED_S6)
    t5 = synth.make_shuf_table(5, cases.SHUF_SEED_S5)
    src = synth.random_bases(100_000, 41)            # the genome cases.fastq_inputs() draws its reads from
    refs = {"r0_src": synth.to_fasta(src, "src", 80), "r1_mut": synth.to_fasta(synth.mutate(src, 0.05, 7), "mut", 80),
            "r2_other": synth.to_fasta(synth.random_bases(100_000, 99), "oth", 80), "r3_half": synth.to_fasta(src[:50_000], "half", 80),
            "r4_half_again": synth.to_fasta(src[:50_000], "half2", 70), "r5_tail": synth.to_fasta(src[90_000:], "tail", 80)}
    fq = cases.fastq_inputs()
    for tag, (k, s, L, tab) in {"composite_l2k8": (8, 5, 2, t5), "composite_l3k10": (10, 6, 3, t6)}.items():
        rr = O.RefRun(k, s, L, tab, shuf_id=cases.SHUF_ID)
        d = rr.dir / "refs"
        d.mkdir()
        for n, b in refs.items():
            (d / f"{n}.fasta").write_bytes(b.tobytes())
        rs = rr.sketch(d, "rsk", p=1)
        q = rr.dir / "q"
        q.mkdir()
        for n, b in fq.items():
            (q / f"{n}.fastq").write_bytes(b.tobytes())
        qs = rr.sketch(q, "qsk", extra=["-A"], p=1)
        r = O.run_ref(["composite", "-r", rs, "-q", qs], cwd=rr.dir)
        assert r.returncode == 0, r.stderr
        rst, qst = O.read_cofiles_stat(rs), O.read_cofiles_stat(qs)
        pack = {"comp_num": np.int32(rst["comp_num"]), "ref_names": np.array([Path(n).name for n in rst["names"]]),
                "qry_names": np.array([Path(n).name for n in qst["names"]]),
                "stdout": np.frombuffer(r.stdout.encode(), dtype=np.uint8)}
        for c in range(rst["comp_num"]):
            rc, ri, _ = O.read_combco(rs, c)
            qc, qi, qa = O.read_combco(qs, c)
            pack[f"ref.{c}"], pack[f"ref.index.{c}"] = rc, ri
            pack[f"qry.{c}"], pack[f"qry.index.{c}"], pack[f"qry.a.{c}"] = qc, qi, qa
        np.savez_compressed(OUT / f"{tag}.npz", **pack)
        print(tag, "\n" + r.stdout[:900].replace(str(rr.dir), ""))
        rr.cleanup()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "composite":
    composite_golden()


def setgroup_golden():
    """`kssd set -g <grouping file>` (grouping_genomes): pan sketch per group of genomes, L3K10 and L3K11 (16 components)."""
    O.build()
    t6 = synth.make_shuf_table(6, cases.SHUF_SEED_S6)
    fa = cases.fasta_inputs()
    tax_lines = ["101\tcladeA", "202\tcladeB", "202\tcladeB", "0", "101\tcladeA", "7"]
    for tag, (k, s, L) in {"setgroup_l3k10": (10, 6, 3), "setgroup_l3k11": (11, 6, 3)}.items():
        rr = O.RefRun(k, s, L, t6, shuf_id=cases.SHUF_ID)
        d = rr.dir / "in"
        d.mkdir()
        for n in ("g_dup", "h_anc", "i_mut1", "j_mut5", "a_plain80", "d_messy"):
            (d / f"{n}.fasta").write_bytes(fa[n].tobytes())
        sk = rr.sketch(d, "sk", p=1)
        st = O.read_cofiles_stat(sk)
        comp = st["comp_num"]
        assert len(st["names"]) == len(tax_lines)
        tax = rr.dir / "groups.tsv"
        tax.write_text("\n".join(tax_lines) + "\n")
        out = rr.dir / "set_g"
        r = O.run_ref(["set", "-g", tax, "-o", out, sk], cwd=rr.dir)
        assert r.returncode == 0, r.stderr
        ost = O.read_cofiles_stat(out)
        pack = {"comp_num": np.int32(comp), "names": np.array([Path(n).name.rsplit(".", 1)[0] for n in st["names"]]), "tax": np.array(tax_lines),
                "g.ctx_ct": ost["ctx_ct"], "g.all_ctx_ct": np.uint64(ost["all_ctx_ct"]), "g.names": np.array([str(n) for n in ost["names"]]),
                "g.infile_num": np.int32(len(ost["names"]))}
        for c in range(comp):
            codes, ix, _ = O.read_combco(sk, c)
            pack[f"in.{c}"] = codes
            pack[f"in.index.{c}"] = ix
            pack[f"g.{c}"] = np.fromfile(out / f"combco.{c}", dtype="<u4")
            pack[f"g.index.{c}"] = np.fromfile(out / f"combco.index.{c}", dtype="<u8")
        np.savez_compressed(OUT / f"{tag}.npz", **pack)
        print(tag, "components", comp, "groups", len(ost["names"]), "group codes", sum(len(pack[f'g.{c}']) for c in range(comp)), list(pack["g.names"]))
        rr.cleanup()


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "setgroup":
    setgroup_golden()
