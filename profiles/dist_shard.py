#!/usr/bin/env python
"""One rank's work of the genome-sharded search, on one GPU: an index of the first R / N references, the full query batch.
usage: python profiles/dist_shard.py [N=8]   (KSSD_SPARSE_SHAPE=wide|narrow forces the CTA shape)"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from public_kssd_b200 import kssd, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_ref, n_qry = 100_000 // N, 10_000
ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1))
rc, ri = synth.synth_sketches(n_ref, 1220, seed=5, cluster_size=20)
qc, qi = synth.synth_sketches(n_qry, 1220, seed=5, cluster_size=2)
dev = torch.device("cuda", 0)
ix = ctx.combco2mco(rc, ri)
tq = torch.from_numpy(qc.view(np.int32)).to(dev)
ti = torch.from_numpy(qi.view(np.int64)).to(dev)
qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
for it in range(4):
    job = kssd.DistJob(ctx, qsz, rsz, sparse=True)
    job.accumulate_dev(ix, tq.data_ptr(), ti.data_ptr(), len(qc))
    n = job.stats(skip_zero=1, fetch=False)
    print(f"shard 1/{N}: {n_ref} refs, rows {n}, count+list {ctx.last_ms(3):.3f} ms, rows {ctx.last_ms(4):.3f} ms")
    job.close()
