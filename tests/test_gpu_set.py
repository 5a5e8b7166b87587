"""GPU parity: `kssd set` (-u / -q / -i / -s) through the C-ABI against the files the unmodified reference wrote
(tests/golden/set_*.npz) and against the oracle on larger random sketches.  Bit-exact, order included."""
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,k", [("set_l3k10", 10), ("set_l3k11", 11)])
def test_set_matches_reference_golden(shuf_l3k10, tag, k):
    from public_kssd_b200 import kssd
    g = np.load(GOLD / f"{tag}.npz", allow_pickle=False)
    ctx = kssd.Context(k, 6, 3, shuf_l3k10)
    try:
        assert ctx.component_num == int(g["comp_num"])
        n = len(g["names"])
        ct = {"i": np.zeros(n, np.uint32), "s": np.zeros(n, np.uint32)}
        for c in range(ctx.component_num):
            pan = ctx.set_union(g[f"in2.{c}"])
            assert np.array_equal(pan, g[f"u.{c}"])
            assert np.array_equal(ctx.set_union(g[f"in.{c}"], uniq=True), g[f"q.{c}"])
            for key, inter in (("i", True), ("s", False)):
                codes, ix = ctx.set_operate(g[f"in.{c}"], g[f"in.index.{c}"], pan, inter)
                assert np.array_equal(codes, g[f"{key}.{c}"]) and np.array_equal(ix, g[f"{key}.index.{c}"])
                ct[key] += np.diff(ix).astype(np.uint32)
        assert np.array_equal(ct["i"], g["i.ctx_ct"]) and np.array_equal(ct["s"], g["s.ctx_ct"])
    finally:
        ctx.close()


def test_set_large_random_matches_oracle(gpu_ctx_l3k10, oracle_mod):
    rc, ri = synth.synth_sketches(3000, 400, seed=12, cluster_size=15)
    rng = np.random.default_rng(3)
    rc = rc.copy()
    for g in range(len(ri) - 1):                    # the reference's order inside a genome is arbitrary: shuffle it
        a, b = int(ri[g]), int(ri[g + 1])
        rng.shuffle(rc[a:b])
    pan_src, _ = synth.synth_sketches(400, 400, seed=12, cluster_size=15)
    pan = gpu_ctx_l3k10.set_union(pan_src)
    assert np.array_equal(pan, oracle_mod.set_union(pan_src))
    assert np.array_equal(gpu_ctx_l3k10.set_union(rc, uniq=True), oracle_mod.set_union(rc, uniq=True))
    for inter in (True, False):
        codes, ix = gpu_ctx_l3k10.set_operate(rc, ri, pan, inter)
        wc, wi = oracle_mod.set_operate(rc, ri, pan, inter)
        assert np.array_equal(codes, wc) and np.array_equal(ix, wi)
    # edge cases: empty input, empty pan, everything / nothing kept
    e = np.zeros(0, np.uint32)
    assert gpu_ctx_l3k10.set_union(e).size == 0
    codes, ix = gpu_ctx_l3k10.set_operate(rc[:100], np.array([0, 40, 40, 100], np.uint64), e, False)
    assert np.array_equal(codes, rc[:100]) and np.array_equal(ix, [0, 40, 40, 100])
    codes, ix = gpu_ctx_l3k10.set_operate(rc[:100], np.array([0, 40, 40, 100], np.uint64), e, True)
    assert codes.size == 0 and np.array_equal(ix, [0, 0, 0, 0])


@pytest.mark.parametrize("tag,k", [("setgroup_l3k10", 10), ("setgroup_l3k11", 11)])
def test_set_grouping_matches_reference_golden(shuf_l3k10, tag, k):
    """`kssd set -g`: union per group on the GPU (distinct codes in first-occurrence order), hash-slot order and the grouping file on
    the host (hostfmt) -- combco.<c>, combco.index.<c>, ctx_ct and names equal to what the unmodified reference wrote."""
    from public_kssd_b200 import hostfmt, kssd
    g = np.load(GOLD / f"{tag}.npz", allow_pickle=False)
    ctx = kssd.Context(k, 6, 3, shuf_l3k10)
    try:
        groups_all, n_lines = hostfmt.organize_taxf("\n".join(str(t) for t in g["tax"]) + "\n")
        assert n_lines == len(g["names"])
        groups = [x for x in groups_all if x["taxid"] != 0]
        assert hostfmt.group_names(groups_all) == [str(n) for n in g["g.names"]]
        ctx_ct = np.zeros(len(groups), np.uint32)
        for c in range(ctx.component_num):
            codes, ix = g[f"in.{c}"], g[f"in.index.{c}"]
            first, fix = ctx.set_group(codes, ix, [x["gids"] for x in groups])
            out, oix = [], [0]
            for t, x in enumerate(groups):
                n_member = int(sum(int(ix[i + 1] - ix[i]) for i in x["gids"]))
                row = hostfmt.group_slot_order(first[int(fix[t]):int(fix[t + 1])], n_member)
                out.append(row)
                oix.append(oix[-1] + row.size)
                ctx_ct[t] += row.size
            assert np.array_equal(np.concatenate(out), g[f"g.{c}"]), (tag, c)
            assert np.array_equal(np.array(oix, np.uint64), g[f"g.index.{c}"]), (tag, c)
        assert np.array_equal(ctx_ct, g["g.ctx_ct"]) and int(ctx_ct.sum()) == int(g["g.all_ctx_ct"])
    finally:
        ctx.close()


def test_set_grouping_large_random_matches_oracle(gpu_ctx_l3k10, oracle_mod):
    from public_kssd_b200 import hostfmt
    rc, ri = synth.synth_sketches(600, 300, seed=21, cluster_size=12)
    rng = np.random.default_rng(5)
    rc = rc.copy()
    for gi in range(len(ri) - 1):                   # the reference's order inside a genome is arbitrary: shuffle it
        a, b = int(ri[gi]), int(ri[gi + 1])
        rng.shuffle(rc[a:b])
    rc[5] = 0                                        # the table's empty marker: never written
    perm = rng.permutation(600)
    groups = [perm[:50].tolist(), perm[50:51].tolist(), [], perm[51:400].tolist(), perm[380:420].tolist()]     # an empty group, an overlap
    first, fix = gpu_ctx_l3k10.set_group(rc, ri, groups)
    want_c, want_i = oracle_mod.set_group(rc, ri, groups)
    got = []
    for t, gids in enumerate(groups):
        n_member = int(sum(int(ri[i + 1] - ri[i]) for i in gids))
        seg = first[int(fix[t]):int(fix[t + 1])]
        # distinct, and in first-occurrence order
        seq = np.concatenate([rc[int(ri[i]):int(ri[i + 1])] for i in gids]) if gids else np.zeros(0, np.uint32)
        _, firsts = np.unique(seq, return_index=True)
        assert np.array_equal(seg, seq[np.sort(firsts)]), t
        got.append(hostfmt.group_slot_order(seg, n_member))
    assert np.array_equal(np.concatenate(got), want_c)
    assert np.array_equal(np.cumsum([0] + [x.size for x in got]).astype(np.uint64), want_i)
