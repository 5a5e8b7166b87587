"""distance.out text: GPU (kssd_dist_format_text) against the host formatter, 1,000 x 1,000 all rows and 5,000 x 100,000 filtered."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from public_kssd_b200 import kssd, synth, hostfmt

ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1), device=0, shuf_id=4242)
for (nr, nq, sparse, opts) in ((1000, 1000, False, dict()), (100_000, 5000, True, dict(skip_zero=1))):
    rc, ri = synth.synth_sketches(nr, 1000, seed=1, cluster_size=50)
    qc, qi = synth.synth_sketches(nq, 1000, seed=1, cluster_size=5)
    ix = ctx.combco2mco(rc, ri)
    qsz, rsz = np.diff(qi).astype(np.uint32), np.diff(ri).astype(np.uint32)
    job = kssd.DistJob(ctx, qsz, rsz, sparse=sparse)
    job.accumulate(ix, qc, qi)
    qn, rn = hostfmt._names_block([f"q{i}.fna" for i in range(nq)]), hostfmt._names_block([f"refs/g{i}.fna" for i in range(nr)])
    for rep in range(3):
        t = time.perf_counter(); n = job.stats(fetch=False, **opts); text = job.distance_out(qn, rn, 0, 2); t_gpu = time.perf_counter() - t
    for rep in range(3):
        t = time.perf_counter(); job.stats(fetch=False, **opts); t_stats = time.perf_counter() - t
        t = time.perf_counter(); view = job.distance_out_view(qn, rn, 0, 2); t_view = time.perf_counter() - t
    print(f"   stats {t_stats*1e3:.2f} ms, text into the context's pinned buffer {t_view*1e3:.2f} ms (kernels {ctx.last_ms(5):.3f} ms), same bytes {bytes(view) == text}", flush=True)
    t = time.perf_counter(); rows = job.stats(**opts); t_rows = time.perf_counter() - t
    t = time.perf_counter(); host = hostfmt.format_distance_out(rows, [f"q{i}.fna" for i in range(nq)], [f"refs/g{i}.fna" for i in range(nr)], 0, 2); t_host = time.perf_counter() - t
    print(f"{nq} x {nr}: rows {n} text {len(text)/1e6:.1f} MB  gpu stats+text {t_gpu*1e3:.1f} ms  host: rows {t_rows*1e3:.1f} ms + snprintf {t_host*1e3:.1f} ms  equal {text == host}", flush=True)
    job.close(); ix.close()
