#!/usr/bin/env python
"""Stage II at BASELINE.json configs[2] size: index of 100,000 reference sketches (122 M postings), device resident.
usage: python profiles/index_scale.py [n_ref]"""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from public_kssd_b200 import kssd, synth

n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1))
rc, ri = synth.synth_sketches(n_ref, 1220, seed=5, cluster_size=20)
dev = torch.device("cuda", 0)
t_rc = torch.from_numpy(rc.view(np.int32)).to(dev)
t_ri = torch.from_numpy(ri.view(np.int64)).to(dev)
ms = []
for it in range(5):
    torch.cuda.synchronize()
    ix = ctx.combco2mco_dev(t_rc.data_ptr(), t_ri.data_ptr(), n_ref, len(rc))
    ms.append(ctx.last_ms(2))
    nu, npost = ix.n_unique, ix.n_postings
    ix.close()
alg = 12 * npost + 8 * nu
print(f"index: {npost} postings, {nu} unique codes; build ms {['%.3f' % m for m in ms]}; best {min(ms):.3f} ms = {alg / min(ms) / 1e6:.0f} GB/s algorithmic "
      f"(12 P + 8 U = {alg / 1e9:.2f} GB) = {alg / min(ms) / 1e6 / 6545.3:.3f} of measured HBM peak")
