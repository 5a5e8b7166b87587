"""GPU parity: `kssd composite` (get_species_abundance) through the C-ABI against the text the unmodified reference printed
(tests/golden/composite_*.npz) and against the oracle on a larger multi-component case.  Integers exact, the two float
columns bit-equal (they are single-precision divisions of integers)."""
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import hostfmt, synth

GOLD = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,k,s,L", [("composite_l2k8", 8, 5, 2), ("composite_l3k10", 10, 6, 3)])
def test_composite_text_identical_to_reference(shuf_s5, shuf_l3k10, tag, k, s, L):
    from public_kssd_b200 import kssd
    g = np.load(GOLD / f"{tag}.npz", allow_pickle=False)
    ctx = kssd.Context(k, s, L, shuf_s5 if s == 5 else shuf_l3k10)
    try:
        nc = int(g["comp_num"])
        idx = [ctx.combco2mco(g[f"ref.{c}"], g[f"ref.index.{c}"]) for c in range(nc)]
        rows = ctx.composite(idx, [g[f"qry.{c}"] for c in range(nc)], [g[f"qry.index.{c}"] for c in range(nc)],
                             [g[f"qry.a.{c}"] for c in range(nc)])
        mine = hostfmt.format_composite_rows(rows, [str(n) for n in g["qry_names"]], [str(n) for n in g["ref_names"]]).splitlines()
        ref_txt = g["stdout"].tobytes().decode()
        norm = ["\t".join([Path(f[0]).name, Path(f[1]).name] + f[2:]) for f in (ln.split("\t") for ln in ref_txt.splitlines())]
        assert mine == norm and len(mine) > 5
        for ix in idx:
            ix.close()
    finally:
        ctx.close()


def test_composite_multi_component_matches_oracle(shuf_l3k10, oracle_mod):
    from public_kssd_b200 import kssd
    ctx = kssd.Context(11, 6, 3, shuf_l3k10)          # 16 components
    try:
        rng = np.random.default_rng(8)
        refs = [synth.synth_sketches(300, 60, seed=40 + c, cluster_size=15) for c in range(16)]
        qrys = [synth.synth_sketches(7, 400, seed=40 + c, cluster_size=2) for c in range(16)]
        qab = [rng.integers(1, 300, len(qc)).astype(np.uint16) for qc, _ in qrys]
        qab[0][:5] = 65535
        idx = [ctx.combco2mco(rc, ri) for rc, ri in refs]
        rows = ctx.composite(idx, [q[0] for q in qrys], [q[1] for q in qrys], qab)
        want = oracle_mod.composite([r[0] for r in refs], [r[1] for r in refs], [q[0] for q in qrys], [q[1] for q in qrys], qab)
        assert len(rows) == len(want) > 20
        for r, w in zip(rows, want):
            assert (int(r["qry"]), int(r["ref"]), int(r["kmer_num"]), int(r["median"]), int(r["max"])) == (w[0], w[1], w[2], w[5], w[6])
            assert r["mean"] == w[3] and r["pct"] == w[4]
        assert len(ctx.composite(idx, [q[0] for q in qrys], [q[1] for q in qrys], qab, min_kmers=10**6)) == 0
        for ix in idx:
            ix.close()
    finally:
        ctx.close()
