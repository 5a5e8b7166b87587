"""CPU, world_size 2 over gloo: the multi-GPU host logic -- genome shards, code-range shards of the reference
index, query broadcast, reduce-scatter of partial count matrices, row ownership.  The per-rank compute is stood in
by the oracle (checker); on the box the same plan drives the CUDA library (tests/multigpu_check.py)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from public_kssd_b200 import parallel, synth  # noqa: E402


def test_plans_are_partitions():
    for n, w in [(1000, 8), (7, 4), (3, 8), (100000, 8)]:
        cover = [g for r in range(w) for g in parallel.genome_shard(n, w, r)]
        assert cover == list(range(n))
        rows = [parallel.row_block(n, w, r) for r in range(w)]
        assert rows[0][0] == 0 and rows[-1][1] == n and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    for w in (1, 2, 3, 8):
        rs = [parallel.code_range(r, w, 28) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == 1 << 28 and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    shards = parallel.balanced_genome_shards([5, 1, 1, 1, 4, 4, 2, 2], 2)
    assert sorted(sum(shards, [])) == list(range(8))
    assert abs(sum([5, 1, 1, 1, 4, 4, 2, 2][g] for g in shards[0]) - 10) <= 1


def test_filter_codes_to_range_keeps_structure():
    codes, index = synth.synth_sketches(50, 200, seed=4)
    parts = [parallel.filter_codes_to_range(codes, index, *parallel.code_range(r, 3, 28)) for r in range(3)]
    assert sum(len(c) for c, _ in parts) == len(codes)
    for g in range(50):
        merged = np.sort(np.concatenate([c[int(ix[g]):int(ix[g + 1])] for c, ix in parts]))
        assert np.array_equal(merged, codes[int(index[g]):int(index[g + 1])])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        rc, ri = synth.synth_sketches(120, 300, seed=5, cluster_size=12)
        qc, qi = synth.synth_sketches(9, 300, seed=5, cluster_size=3)
        R, Q = len(ri) - 1, len(qi) - 1
        # queries live on rank 0 and are broadcast
        meta = torch.tensor([Q, len(qc)] if rank == 0 else [0, 0], dtype=torch.int64)
        dist.broadcast(meta, 0)
        tq = torch.from_numpy(qc.view(np.int32).copy()) if rank == 0 else torch.empty(int(meta[1]), dtype=torch.int32)
        ti = torch.from_numpy(qi.view(np.int64).copy()) if rank == 0 else torch.empty(int(meta[0]) + 1, dtype=torch.int64)
        dist.broadcast(tq, 0)
        dist.broadcast(ti, 0)
        bq, bi = tq.numpy().view(np.uint32), ti.numpy().view(np.uint64)
        # this rank's slice of the reference index, partial counts by the checker
        lo, hi = parallel.code_range(rank, world, 28)
        c, ix = parallel.filter_codes_to_range(rc, ri, lo, hi)
        uc, uo, gids = O.csr_from_combco(c, ix)
        part = O.dist_counts(bq, bi, uc, uo, gids, R)
        per = (Q + world - 1) // world
        padded = np.zeros((per * world, R), dtype=np.int32)
        padded[:Q] = part.view(np.int32)
        mine = parallel.reduce_scatter_rows(torch.from_numpy(padded), world, rank)
        r0, r1 = parallel.row_block(Q, world, rank)
        fuc, fuo, fg = O.csr_from_combco(rc, ri)
        full = O.dist_counts(qc, qi, fuc, fuo, fg, R)
        ok = np.array_equal(mine.numpy()[: r1 - r0].view(np.uint32), full[r0:r1]) and part.sum() < full.sum()
        q.put((rank, bool(ok), int(part.sum()), int(full.sum())))
    finally:
        dist.destroy_process_group()


def test_sharded_dist_plan_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert sum(s for _, _, s, _ in res) == res[0][3]        # partial counts add up to the full matrix


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 41                                              # not a multiple of the world size
        rc, ri = synth.synth_sketches(n, 250, seed=9, cluster_size=7)
        g = parallel.genome_shard(n, world, rank)
        lc, li = parallel.genome_block(rc, ri, g.start, g.stop)
        c, ix = parallel.exchange_codes_by_range(lc, li, g.start, n, world, rank, 28)
        wc, wi = parallel.filter_codes_to_range(rc, ri, *parallel.code_range(rank, world, 28))
        ok = np.array_equal(c.numpy().view(np.uint32), wc) and np.array_equal(ix.numpy().astype(np.uint64), wi)
        q.put((rank, bool(ok), int(c.numel())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_codes_by_range(world):
    """Stage I shards (by genome) -> Stage II shards (by code range) through one all-to-all: every rank ends up with
    exactly the slice filter_codes_to_range would cut from the whole reference, genomes in order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(nc for _, _, nc in res) == len(synth.synth_sketches(41, 250, seed=9, cluster_size=7)[0])
