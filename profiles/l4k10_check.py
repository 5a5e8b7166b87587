#!/usr/bin/env python
"""BASELINE.json configs[3] shape, scaled: a many-contig genome (log-normal contig lengths, N runs, soft-masked
stretches, 60-col lines) sketched at L4K10 with the auto subk = 7 (.shuf payload 1 GiB), checked against the oracle.
usage: python profiles/l4k10_check.py [Mbp]   (default 300)"""
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from public_kssd_b200 import kssd, synth
from oracle import oracle as O

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 300
t0 = time.time()
tab = synth.make_shuf_table(7, 3)
print(f"subk=7 table {tab.nbytes / 2**30:.2f} GiB in {time.time() - t0:.0f}s", flush=True)
rng = np.random.default_rng(11)
parts = []
left = mbp * 1_000_000
c = 0
while left > 0:
    n = int(min(left, max(2000, rng.lognormal(11.0, 1.2))))
    b = synth.random_bases(n, 1000 + c)
    txt = np.frombuffer(b"ACGT", dtype=np.uint8)[b].copy()
    # soft-masked stretch and N runs (~1 %)
    s0 = int(rng.integers(0, max(n - 500, 1))); txt[s0:s0 + int(rng.integers(50, 5000))] |= 0x20
    for _ in range(max(1, n // 200_000)):
        p = int(rng.integers(0, n)); txt[p:p + int(rng.integers(1, 4000))] = ord("N")
    body = txt.tobytes()
    lines = b"\n".join(body[i:i + 60] for i in range(0, len(body), 60))
    parts.append(b">contig%d len=%d\n" % (c, n) + lines + b"\n")
    left -= n; c += 1
genome = np.frombuffer(b"".join(parts), dtype=np.uint8)
print(f"{c} contigs, {genome.size / 1e6:.0f} MB of FASTA", flush=True)
ctx = kssd.Context(10, 7, 4, tab)
print("ctx: sampled", ctx.info.n_sampled, "dim_end", ctx.info.dim_end, "hashsize", ctx.info.hashsize, flush=True)
for it in range(3):
    sk = ctx.sketch([genome], strict=False)
print(f"GPU: {len(sk.ids[0])} codes, scan {sk.scan_ms:.3f} ms = {genome.size / sk.scan_ms / 1e6:.0f} GB/s ({genome.size / sk.scan_ms / 1e6 / 6545.3:.3f} of HBM peak), status {sk.status}", flush=True)
t0 = time.time()
orc = O.Ctx(10, 7, 4, tab)
ids, comp = orc.fasta(genome)
print(f"oracle: {len(ids)} codes in {time.time() - t0:.1f}s; parity {np.array_equal(np.sort(ids), sk.ids[0])}")
