#!/usr/bin/env python
"""Stage I from files on tmpfs: kssd_stage1_files (plain and .gz) beside the unmodified reference binary on the same
files.  usage: python profiles/stage1_files.py [genomes] [genome_len]   (defaults 200 x 5,000,000 bp)"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from public_kssd_b200 import hostfmt, kssd, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
root = Path(tempfile.mkdtemp(prefix="kssd_files_", dir="/dev/shm"))
try:
    plain, gz = root / "plain", root / "gz"
    plain.mkdir(); gz.mkdir()
    t0 = time.time()
    total = 0
    for i, (name, bases) in enumerate(synth.cluster_genomes(n, glen, seed=7, cluster_size=20)):
        txt = synth.to_fasta(bases, name, 80).tobytes()
        (plain / f"g{i:05d}.fna").write_bytes(txt)
        if i < n // 4:
            with gzip.open(gz / f"g{i:05d}.fna.gz", "wb", compresslevel=1) as f:
                f.write(txt)
        total += len(txt)
    print(f"wrote {n} genomes, {total / 1e9:.2f} GB of FASTA in {time.time() - t0:.1f}s", flush=True)
    tab = synth.make_shuf_table(6, 1)
    ctx = kssd.Context(10, 6, 3, tab, shuf_id=4242)
    pp = sorted(plain.glob("*.fna"))
    gp = sorted(gz.glob("*.gz"))
    for label, paths in (("plain", pp), ("gz (level 1)", gp)):
        best = None
        for it in range(3):
            sk, t = ctx.sketch_files(paths, batch_bytes=1 << 30)
            best = t if best is None or t["total_s"] < best["total_s"] else best
        bp = best["bytes"]
        print(f"{label}: {len(paths)} files, {bp / 1e9:.2f} GB decoded in {best['total_s']:.3f} s = {bp / best['total_s'] / 1e9:.2f} GB/s "
              f"(reader threads busy {best['read_s']:.3f} s each, GPU calls {best['gpu_s']:.3f} s, {best['batches']} batches, "
              f"{os.cpu_count()} host threads); codes {len(sk.ids[0])}", flush=True)
        if label == "plain":
            mine = sk
    ref_bin = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "kssd"
    if ref_bin.exists():
        shuf = root / "L3K10.shuf"
        hostfmt_ok = True
        import struct
        with open(shuf, "wb") as f:
            f.write(struct.pack("<iiii", 4242, 10, 6, 3)); f.write(np.ascontiguousarray(tab, dtype="<i4").tobytes())
        t0 = time.time()
        r = subprocess.run([str(ref_bin), "dist", "-p", str(os.cpu_count()), "-L", str(shuf), "-o", str(root / "ref_out"), str(plain)],
                           capture_output=True, text=True)
        dt = time.time() - t0
        st = hostfmt.read_cofiles_stat(root / "ref_out")
        codes, index, _ = hostfmt.read_combco(root / "ref_out", 0)
        by_name = {Path(nm).name: np.sort(codes[int(index[i]):int(index[i + 1])]) for i, nm in enumerate(st["names"])}
        same = all(np.array_equal(by_name[p.name], mine.ids[0][int(mine.index[0][i]):int(mine.index[0][i + 1])]) for i, p in enumerate(pp))
        print(f"reference `kssd dist -p {os.cpu_count()}` on the plain files: {dt:.2f} s = {total / dt / 1e9:.3f} GB/s; same sketches: {same}", flush=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
