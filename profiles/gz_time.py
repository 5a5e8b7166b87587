"""gzip'ed genomes: inflate on the GPU (one file per thread) against zlib on the host cores, by the number of files in the call."""
import os, sys, time, zlib, tempfile, shutil
from pathlib import Path
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from public_kssd_b200 import kssd, synth

ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1), device=0, shuf_id=4242)
n_max, glen = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, 5_000_000
work = Path(tempfile.mkdtemp(prefix="kssd_gz_", dir="/dev/shm"))
try:
    def make(i):
        txt = synth.to_fasta(synth.random_bases(glen, 1000 + i), f"g{i}", width=80).tobytes()
        co = zlib.compressobj(1, zlib.DEFLATED, 31)
        p = work / f"g{i:05d}.fa.gz"
        p.write_bytes(co.compress(txt) + co.flush())
        return p, len(txt)
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        made = list(ex.map(make, range(n_max)))
    paths = [m[0] for m in made]
    on_disk = sum(p.stat().st_size for p in paths)
    print(f"{n_max} files, {sum(m[1] for m in made)/1e9:.2f} GB of text, {on_disk/1e9:.2f} GB on disk", flush=True)
    sizes = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else sorted(set([64, 200, 400, 1000, 2000, n_max]))
    for n in sizes:
        if n > n_max: continue
        res = {}
        for mode in ("0", "1"):
            os.environ["KSSD_GZ_GPU"] = mode
            best = None
            for rep in range(2 if n > 1000 else 3):
                sk, t = ctx.sketch_files(paths[:n])
                if best is None or t["total_s"] < best["total_s"]: best = t
            res[mode] = (best, sk.ids[0].copy())
        h, g = res["0"][0], res["1"][0]
        print(f"files {n}: host zlib {h['bytes']/h['total_s']/1e9:.2f} GB/s of text ({h['total_s']*1e3:.0f} ms) | GPU inflate {g['bytes']/g['total_s']/1e9:.2f} GB/s ({g['total_s']*1e3:.0f} ms; "
              f"H2D+inflate {g['gz_gpu_s']*1e3:.0f} ms, read {g['read_s']*1e3:.0f} ms, on_gpu {g['gz_on_gpu']}) same ids {np.array_equal(res['0'][1], res['1'][1])}", flush=True)
finally:
    shutil.rmtree(work, ignore_errors=True)
