#!/usr/bin/env python
"""N-GPU check, launched with torchrun (NCCL):  python -m torch.distributed.run --nproc-per-node N tests/multigpu_check.py
  1. Stage I sharded by genome: every rank sketches its genomes, results gathered on rank 0 == single-GPU result.
  2. Stage III sharded by code range (NCCL reduce-scatter, and the peer-memory count kernel) == single-GPU count
     matrix and statistics, bit for bit.
  3. The whole path across ranks: Stage I shards -> all-to-all by code range -> per-rank index -> peer-memory search
     == single-GPU all-vs-all of the same genomes.
Not a pytest file: the default GPU tier has one device."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from public_kssd_b200 import kssd, parallel, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tab = synth.make_shuf_table(5, 2)
    ctx = kssd.Context(8, 5, 2, tab, device=local)
    # ---- 1. sketch sharded by genome ----
    genomes = [synth.to_fasta(b, n, 80) for n, b in synth.cluster_genomes(23, 150_000, seed=3, cluster_size=5)]
    mine = list(parallel.genome_shard(len(genomes), world, rank))
    sk = ctx.sketch([genomes[g] for g in mine]) if mine else None
    payload = [(g, sk.ids[0][int(sk.index[0][i]):int(sk.index[0][i + 1])]) for i, g in enumerate(mine)]
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    ok1 = True
    if rank == 0:
        full = ctx.sketch(genomes)
        sets = {g: ids for part in gathered for g, ids in part}
        ok1 = sorted(sets) == list(range(len(genomes))) and all(
            np.array_equal(sets[g], full.ids[0][int(full.index[0][g]):int(full.index[0][g + 1])]) for g in sets)
    # ---- 2. dist sharded by code range ----
    rc, ri = synth.synth_sketches(3000, 300, seed=5, cluster_size=20, code_bits=20)
    qc, qi = synth.synth_sketches(101, 300, seed=5, cluster_size=4, code_bits=20)
    opts = dict(metric=0, correction=0, dthreshold=1.0, skip_zero=1)
    results = {}
    for mode in ("code", "code_p2p"):
        sd = parallel.ShardedDist(ctx, world, rank, code_bits=20, mode=mode).build_reference(rc, ri)
        lo, hi, block, rows = sd.search(qc if rank == 0 else None, qi if rank == 0 else None, src=0, stats_opts=opts)
        got = [None] * world
        dist.all_gather_object(got, (lo, hi, block, rows))
        results[mode] = got
        sd.close()
    parts = results["code"]
    ok2 = True
    if rank == 0:
        ix = ctx.combco2mco(rc, ri)
        job = kssd.DistJob(ctx, np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32))
        job.accumulate(ix, qc, qi)
        ct = job.counts()
        ref_rows = job.stats(**opts)
        got = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0]) if p[1] > p[0]])
        got_rows = np.concatenate([p[3] for p in sorted(parts, key=lambda p: p[0]) if p[3] is not None and len(p[3])])
        ok2 = np.array_equal(got, ct) and got_rows.tobytes() == ref_rows.tobytes() and ct.sum() > 0
        p2p = results["code_p2p"]
        got2 = np.concatenate([p[2] for p in sorted(p2p, key=lambda p: p[0]) if p[1] > p[0]])
        rows2 = np.concatenate([p[3] for p in sorted(p2p, key=lambda p: p[0]) if p[3] is not None and len(p[3])])
        ok2 = ok2 and np.array_equal(got2, ct) and rows2.tobytes() == ref_rows.tobytes()
        print(f"multigpu_check world={world}: sketch_sharding={'ok' if ok1 else 'FAIL'} dist_code_range_reduce_scatter_and_p2p="
              f"{'ok' if ok2 else 'FAIL'} shared_total={int(ct.sum())} rows={len(ref_rows)}")
    # ---- 3. sketch shards -> exchange by code range -> index -> search, nothing gathered on one rank ----
    code_bits = 4 * (8 - 2)
    n_g = len(genomes)
    local_sizes = np.diff(sk.index[0]).astype(np.uint32) if mine else np.zeros(0, np.uint32)
    all_sizes = [None] * world
    dist.all_gather_object(all_sizes, local_sizes)
    ref_sizes = np.concatenate(all_sizes)
    sd = parallel.ShardedDist(ctx, world, rank, code_bits=code_bits, mode="code_p2p")
    sd.build_reference_exchanged(sk.ids[0] if mine else np.zeros(0, np.uint32), sk.index[0] if mine else np.zeros(1, np.uint64),
                                 mine[0] if mine else 0, n_g, ref_sizes)
    if rank == 0:
        qc3, qi3 = full.ids[0], full.index[0]
    lo, hi, block, rows = sd.search(qc3 if rank == 0 else None, qi3 if rank == 0 else None, src=0, stats_opts=opts)
    got3 = [None] * world
    dist.all_gather_object(got3, (lo, hi, block, rows))
    sd.close()
    ok3 = True
    if rank == 0:
        ix = ctx.combco2mco(full.ids[0], full.index[0])
        sz = np.diff(full.index[0]).astype(np.uint32)
        job = kssd.DistJob(ctx, sz, sz)
        job.accumulate(ix, full.ids[0], full.index[0])
        ct3 = job.counts()
        rows3 = job.stats(**opts)
        g3 = np.concatenate([p[2] for p in sorted(got3, key=lambda p: p[0]) if p[1] > p[0]])
        r3 = np.concatenate([p[3] for p in sorted(got3, key=lambda p: p[0]) if p[3] is not None and len(p[3])])
        ok3 = np.array_equal(ref_sizes, sz) and np.array_equal(g3, ct3) and r3.tobytes() == rows3.tobytes() and np.array_equal(np.diag(ct3), sz)
        print(f"multigpu_check world={world}: sketch->exchange->index->search pipeline={'ok' if ok3 else 'FAIL'} "
              f"shared_total={int(ct3.sum())} rows={len(rows3)}")
    flag = torch.tensor([int(ok1 and ok2 and ok3)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
