#!/usr/bin/env python
"""BASELINE.json configs[2] on N GPUs (torchrun): R reference sketches x ~1220 codes indexed across the ranks, Q
queries searched against them, both sharding schemes of parallel.ShardedDist.
usage: torchrun --nproc-per-node N profiles/dist_scale_multi.py [R] [Q]   (defaults 100000 10000)"""
import os
import sys
import time
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from public_kssd_b200 import kssd, parallel, synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rc, ri = synth.synth_sketches(R, 1220, seed=5, cluster_size=20)
qc, qi = synth.synth_sketches(Q, 1220, seed=5, cluster_size=2)
ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1), device=local)
dev = torch.device("cuda", local)
tq = torch.from_numpy(qc.view(np.int32)).to(dev) if rank == 0 else None      # queries start resident on rank 0
ti = torch.from_numpy(qi.view(np.int64)).to(dev) if rank == 0 else None
out = {}
for mode in ("code", "code_p2p", "genome"):
    sd = parallel.ShardedDist(ctx, world, rank, code_bits=28, mode=mode).build_reference(rc, ri)
    best = 1e9
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        lo, hi, block, rows = sd.search(tq, ti, src=0, stats_opts=dict(skip_zero=1), fetch_counts=(it == 2), fetch_stats=False)
        torch.cuda.synchronize(); dist.barrier()
        best = min(best, time.perf_counter() - t0) if it < 2 else best
    tot = torch.tensor([int(block.sum(dtype=np.uint64)) if block is not None else 0, int(rows) if rows is not None else 0], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    t = torch.tensor([best], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[mode] = (float(t.item()), int(tot[0]), int(tot[1]))
    sd.close()
if rank == 0:
    for mode, (t, shared, nrows) in out.items():
        print(f"world={world} mode={mode}: {Q}x{R} pairs in {t * 1e3:.2f} ms (queries broadcast + count"
              f"{' + reduce-scatter' if mode == 'code' else (' fused with the peer-memory reduction' if mode == 'code_p2p' else '')} + statistics, wall clock, max over ranks) = {Q * R / t:.3e} pairs/s; "
              f"shared total {shared}, rows {nrows}")
    assert out["code"][1:] == out["genome"][1:] == out["code_p2p"][1:], "the sharding schemes disagree"
dist.destroy_process_group()
