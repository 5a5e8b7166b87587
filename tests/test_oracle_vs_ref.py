"""CPU: live pin of the oracle restatement against the UNMODIFIED reference binary (oracle/_ref/kssd),
on inputs other than the golden ones.  Skipped where the binary is absent (it is built from
/root/reference by oracle/Makefile and travels to the GPU box as a prebuilt file)."""
import numpy as np
import pytest

from public_kssd_b200 import synth


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref/kssd not built")
    return oracle_mod


def _run_case(O, k, s, L, table, inputs, ext, extra, fn):
    rr = O.RefRun(k, s, L, table, shuf_id=777)
    try:
        d = rr.dir / "in"
        d.mkdir()
        for n, b in inputs.items():
            (d / f"{n}.{ext}").write_bytes(b.tobytes())
        out = rr.sketch(d, "sk", extra=extra, p=1)
        st = O.read_cofiles_stat(out)
        comps = [O.read_combco(out, c) for c in range(st["comp_num"])]
        ctx = O.Ctx(k, s, L, table)
        for i, nm in enumerate(st["names"]):
            key = nm.rsplit("/", 1)[-1].rsplit(".", 1)[0]
            res = fn(ctx, inputs[key])
            ids, comp = res[0], res[1]
            assert st["ctx_ct"][i] == ids.size
            for c, (codes, ix, ab) in enumerate(comps):
                assert np.array_equal(codes[int(ix[i]):int(ix[i + 1])], ids[comp == c]), (key, c)
                if ab is not None:
                    assert np.array_equal(ab[int(ix[i]):int(ix[i + 1])], res[2][comp == c])
    finally:
        rr.cleanup()


def test_fasta_live(ref, shuf_l3k10):
    inputs = {f"m{i}": synth.messy_fasta(150_000 + 7001 * i, 100 + i, ncontigs=5 + i, width=50 + 9 * i, crlf=bool(i & 1)) for i in range(4)}
    inputs["hdr_only"] = np.frombuffer(b">x\n>y\n>z\nACGT\n", dtype=np.uint8)
    inputs["gt_mid"] = np.frombuffer(b">x\n" + b"ACGTTGCAAGCTTGCATGCAAGGTCCATG" * 400 + b">mid ACGT\n" + b"TTGACCATGCATGGACTGACTGGT" * 300 + b"\n", dtype=np.uint8)
    _run_case(ref, 10, 6, 3, shuf_l3k10, inputs, "fasta", [], lambda c, b: c.fasta(b))
    _run_case(ref, 10, 6, 3, shuf_l3k10, inputs, "fasta", ["-u"], lambda c, b: c.fasta(b, uniq=True))


def test_fastq_live(ref, shuf_s5):
    src = synth.random_bases(40_000, 201)
    inputs = {"a": synth.to_fastq(src, 3000, 100, seed=202), "b": synth.to_fastq(src, 500, 151, seed=203, trailing_newline=False)}
    _run_case(ref, 8, 5, 2, shuf_s5, inputs, "fq", ["-Q", "45", "-n", "2"], lambda c, b: c.fastq(b, 45, 2))
    _run_case(ref, 8, 5, 2, shuf_s5, inputs, "fq", ["-A"], lambda c, b: c.fastq_abund(b))


def test_combine_queries_matches_reference(ref, shuf_s5, tmp_path):
    """hostfmt.combine_queries vs `kssd dist -o combined dirA dirB` (command_dist.c:1323-1475)."""
    from public_kssd_b200 import hostfmt
    O = ref
    rr = O.RefRun(8, 5, 2, shuf_s5, shuf_id=99)
    try:
        dirs = []
        for j, seeds in enumerate([(1, 2, 3), (4, 5)]):
            d = rr.dir / f"in{j}"
            d.mkdir()
            for s in seeds:
                (d / f"g{s}.fasta").write_bytes(synth.messy_fasta(60_000 + 999 * s, 300 + s).tobytes())
            dirs.append(rr.sketch(d, f"sk{j}", p=1))
        comb = rr.dir / "combined"
        r = O.run_ref(["dist", "-p", 1, "-o", comb, dirs[0], dirs[1]], cwd=rr.dir)
        assert r.returncode == 0 and (comb / "cofiles.stat").exists(), r.stderr[-300:]
        mine = tmp_path / "mine"
        st = hostfmt.combine_queries(dirs, mine)
        want = hostfmt.read_cofiles_stat(comb)
        assert st["infile_num"] == want["infile_num"] == 5 and st["all_ctx_ct"] == want["all_ctx_ct"]
        assert np.array_equal(st["ctx_ct"], want["ctx_ct"]) and st["names"] == want["names"]
        assert (mine / "combco.0").read_bytes() == (comb / "combco.0").read_bytes()
        assert (mine / "combco.index.0").read_bytes() == (comb / "combco.index.0").read_bytes()
    finally:
        rr.cleanup()
