"""GPU parity: `--byread` (reference reads2mco, iseq2comem.c:78-186) through the C-ABI (KSSD_MODE_BYREAD) against the
goldens written by the unmodified reference and against the oracle on extra edge cases.  Bit-exact: ids in stream
order with duplicates, the per-record inclusive index, record counts."""
import sys
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,k,s,L", [("byread_l2k8", 8, 5, 2), ("byread_l3k11", 11, 6, 3)])
def test_byread_matches_reference_golden(shuf_s5, shuf_l3k10, tag, k, s, L):
    from public_kssd_b200 import kssd
    g = np.load(GOLD / f"{tag}.npz", allow_pickle=False)
    ctx = kssd.Context(k, s, L, shuf_s5 if s == 5 else shuf_l3k10)
    try:
        files = cases.byread_inputs()
        names = sorted(files)
        res = ctx.reads2mco([files[n] for n in names])          # all files in ONE batch
        for n, rec in zip(names, res):
            assert ctx.component_num == int(g[f"{n}.comp_num"])
            for c in range(ctx.component_num):
                assert np.array_equal(rec["ids"][c], g[f"{n}.{c}"]), (tag, n, c)
                assert np.array_equal(rec["index"][c], g[f"{n}.{c}.index"]), (tag, n, c)
                assert len(rec["index"][c]) == rec["n_reads"] + 1
    finally:
        ctx.close()


def _edge_files():
    src = synth.random_bases(50_000, 401)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    g = {}
    g["no_header_at_all"] = np.concatenate([acgt[src[:3000]], np.frombuffer(b"\n", dtype=np.uint8)])
    g["gt_gt_same_line"] = np.frombuffer(b">a>b>c\n" + acgt[src[:500]].tobytes() + b"\n>>x\n" + acgt[src[500:900]].tobytes() + b"\n", dtype=np.uint8)
    g["header_only_records"] = np.frombuffer(b">r1\n>r2\n>r3\n" + acgt[src[:200]].tobytes() + b"\n>r4\n", dtype=np.uint8)
    g["gt_mid_sequence"] = np.frombuffer(b">r\n" + acgt[src[:300]].tobytes() + b">junk ACGTACGTACGTACGTACGT\n" + acgt[src[300:700]].tobytes() + b"\n",
                                         dtype=np.uint8)
    g["long_reads"] = synth.to_read_fasta(src, 40, seed=402, min_len=5000, max_len=20000, width=0)
    g["many_short"] = synth.to_read_fasta(src, 5000, seed=403, min_len=16, max_len=40, width=0, messy=True)
    g["empty_tail_record"] = np.frombuffer(b">r1\n" + acgt[src[:100]].tobytes() + b"\n>r2\n", dtype=np.uint8)
    return g


def test_byread_edge_cases_match_oracle(shuf_s5, oracle_mod):
    from public_kssd_b200 import kssd
    ctx = kssd.Context(8, 5, 2, shuf_s5)
    octx = oracle_mod.Ctx(8, 5, 2, shuf_s5)
    try:
        files = _edge_files()
        names = sorted(files)
        res = ctx.reads2mco([files[n] for n in names])
        total = 0
        for n, rec in zip(names, res):
            want_reads, want = octx.byread(files[n])
            assert rec["n_reads"] == want_reads, n
            assert np.array_equal(rec["ids"][0], want[0][0]), n
            assert np.array_equal(rec["index"][0], want[0][1]), n
            total += len(want[0][0])
        assert total > 1000
    finally:
        ctx.close()


def test_byread_header_into_eof_is_reported(shuf_s5):
    from public_kssd_b200 import kssd
    ctx = kssd.Context(8, 5, 2, shuf_s5)
    try:
        bad = np.frombuffer(b">r1\nACGTACGTACGTACGTACGTACGTACGT\n>r2 no newline", dtype=np.uint8)
        with pytest.raises(kssd.KssdError):
            ctx.reads2mco([bad])
    finally:
        ctx.close()
