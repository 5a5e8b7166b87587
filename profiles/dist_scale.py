#!/usr/bin/env python
"""Stage II/III at BASELINE.json configs[2] scale on one GPU: R refs x ~1220 codes (28-bit), Q queries from the same
clusters.  usage: python profiles/dist_scale.py [R] [Q] [codes]   (defaults 100000 10000 1220)"""
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from public_kssd_b200 import kssd, synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 1220
t0 = time.time()
rc, ri = synth.synth_sketches(R, C, seed=5, cluster_size=20)
qc, qi = synth.synth_sketches(Q, C, seed=5, cluster_size=2)
print(f"synth {time.time() - t0:.1f}s refs {len(rc)} codes, queries {len(qc)} codes", flush=True)
ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1))
ix = ctx.combco2mco(rc, ri)
print(f"index: {ix.n_postings} postings, {ix.n_unique} unique, build {ctx.last_ms(2):.3f} ms", flush=True)
qs, rs = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
for it in range(3):
    job = kssd.DistJob(ctx, qs, rs)
    job.accumulate(ix, qc, qi)
    cms = ctx.last_ms(3)
    n = job.stats(skip_zero=1, fetch=False)
    sms = ctx.last_ms(4)
    if it == 2:
        ct = job.counts()
        P = int(ct.sum(dtype=np.uint64))
        byt = 4 * len(qc) + 8 * len(qc) + 4 * P + 4 * Q * R
        print(f"count {cms:.3f} ms  stats(skip_zero) {sms:.3f} ms rows {n}  P={P}  pairs/s {Q * R / ((cms + sms) * 1e-3):.3e}  "
              f"count-kernel {byt / cms / 1e6:.0f} GB/s = {byt / cms / 1e6 / 6545.3:.3f} of measured HBM peak", flush=True)
        # spot check against brute force
        for q, r in [(0, 0), (1, 1), (Q - 1, R - 1), (17, 170)]:
            a = qc[int(qi[q]):int(qi[q + 1])]; b = rc[int(ri[r]):int(ri[r + 1])]
            assert ct[q, r] == np.intersect1d(a, b).size
    job.close()
# the same search as a sparse job: no Q x R matrix, counts in a per-query shared-memory hash table
import torch
tq = torch.from_numpy(qc.view(np.int32)).cuda()
ti = torch.from_numpy(qi.view(np.int64)).cuda()
for opts, tag in [(dict(skip_zero=1), "skip_zero"), (dict(dthreshold=0.3), "-D 0.3")]:
    dense = kssd.DistJob(ctx, qs, rs)
    dense.accumulate_dev(ix, tq.data_ptr(), ti.data_ptr(), len(qc))
    cms = ctx.last_ms(3)
    want = dense.stats(**opts)
    dms = ctx.last_ms(4)
    dense.close()
    best = None
    for it in range(3):
        sp = kssd.DistJob(ctx, qs, rs, sparse=True)
        sp.accumulate_dev(ix, tq.data_ptr(), ti.data_ptr(), len(qc))
        n = sp.stats(fetch=False, **opts)
        t = (ctx.last_ms(3), ctx.last_ms(4))
        best = t if best is None or sum(t) < sum(best) else best
        if it == 2:
            rows = np.empty(n, dtype=want.dtype)
            from public_kssd_b200.capi import check, lib
            check(lib().kssd_dist_fetch_stats(sp._h, rows.ctypes.data_as(__import__("ctypes").c_void_p)))
            same = rows.tobytes() == want.tobytes()
        sp.close()
    print(f"sparse job ({tag}): count+list {best[0]:.3f} ms + rows {best[1]:.3f} ms = {sum(best):.3f} ms -> {Q * R / (sum(best) * 1e-3):.3e} pairs/s "
          f"(dense job: {cms:.3f} + {dms:.3f} ms), rows {n}, identical to the dense job: {same}", flush=True)
