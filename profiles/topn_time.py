"""-N (top-n neighbours) on the sparse job against the matrix path: 2,000 queries x 100,000 references of 1,000 codes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from public_kssd_b200 import kssd, synth

ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1), device=0, shuf_id=4242)
rc, ri = synth.synth_sketches(100_000, 1000, seed=1, cluster_size=50)
qc, qi = synth.synth_sketches(2000, 1000, seed=1, cluster_size=5)
ix = ctx.combco2mco(rc, ri)
qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
out = {}
for name, env in (("sparse", None), ("matrix", "1")):
    if env: os.environ["KSSD_TOPN_DENSE"] = env
    job = kssd.DistJob(ctx, qsz, rsz, sparse=True)
    job.accumulate(ix, qc, qi)
    for rep in range(3):
        t = time.perf_counter(); rows = job.stats(n_neighbors=10); dt = time.perf_counter() - t
    out[name] = rows.tobytes()
    print(name, "rows", len(rows), "ms", round(dt * 1e3, 2), flush=True)
    job.close()
print("identical", out["sparse"] == out["matrix"])
