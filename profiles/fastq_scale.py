#!/usr/bin/env python
"""FASTQ Stage I at a scaled-down BASELINE.json configs[4] shape: N reads x 150 bp, Phred+33, L3K11 (16 components), -n 2.
usage: python profiles/fastq_scale.py [reads]   (default 4,000,000 = 1.3 GB of text)"""
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from public_kssd_b200 import capi, kssd, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rl = 150
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
src = torch.randint(0, 4, (2_000_000,), generator=g, device=dev, dtype=torch.uint8)
starts = torch.randint(0, src.numel() - rl, (n_reads,), generator=g, device=dev)
idx = starts[:, None] + torch.arange(rl, device=dev)[None, :]
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
bases = lut[src[idx].long()]
err = torch.rand((n_reads, rl), generator=g, device=dev) < 0.005
bases = torch.where(err, lut[torch.randint(0, 4, (n_reads, rl), generator=g, device=dev)], bases)
qual = torch.randint(35, 74, (n_reads, rl), generator=g, device=dev, dtype=torch.uint8)
hdr = torch.full((n_reads, 12), ord("x"), dtype=torch.uint8, device=dev); hdr[:, 0] = ord("@"); hdr[:, 11] = 10
num = torch.arange(n_reads, device=dev)
for d in range(10):
    hdr[:, 10 - d] = (48 + (num // (10 ** d)) % 10).to(torch.uint8)
plus = torch.tensor([43, 10], dtype=torch.uint8, device=dev).expand(n_reads, 2)
nl = torch.full((n_reads, 1), 10, dtype=torch.uint8, device=dev)
rec = torch.cat([hdr, bases, nl, plus, qual, nl], dim=1).contiguous().view(-1)
buf = torch.cat([rec, torch.full((1024,), 10, dtype=torch.uint8, device=dev)])
nbytes = int(rec.numel())
torch.cuda.synchronize()
tab6 = synth.make_shuf_table(6, 1)
ctx = kssd.Context(11, 6, 3, tab6)
goff = np.zeros(1, dtype=np.uint64); glen = np.array([nbytes], dtype=np.uint64)
res = {}
for mode, name in [(capi.MODE_FASTQ, "fastq2co -n 2"), (capi.MODE_FASTQ_ABUND, "-A")]:
    ms = []
    for it in range(4):
        t0 = time.perf_counter()
        h = ctx.sketch_raw(None, nbytes, goff, glen, mode=mode, Q=0, M=2, device_ptr=buf.data_ptr())
        ms.append((ctx.last_ms(0), (time.perf_counter() - t0) * 1e3))
        sk = ctx.fetch_sketch(h, 1, want_abund=(mode == capi.MODE_FASTQ_ABUND))
    k, w = min(m[0] for m in ms[1:]), min(m[1] for m in ms[1:])
    print(f"{name}: text {nbytes / 1e9:.2f} GB, bases {n_reads * rl / 1e9:.2f} Gbp, line index + scan {k:.2f} ms = {nbytes / k / 1e6:.0f} GB/s "
          f"({nbytes / k / 1e6 / 6545.3:.3f} of measured HBM peak, {n_reads * rl / k / 1e6:.0f} Gbp/s), call {w:.2f} ms, "
          f"codes {sum(len(x) for x in sk.ids)} in {ctx.component_num} components")
# parity of a slice against the oracle (checker)
from oracle import oracle as O
orc = O.Ctx(11, 6, 3, tab6)
small = rec[: 20000 * (12 + rl + 1 + 2 + rl + 1)].cpu().numpy()
ids, comp = orc.fastq(small, 0, 2)
sk = ctx.sketch_fastq([small], Q=0, M=2)
ok = all(np.array_equal(sk.genome_sets()[0][c], np.sort(ids[comp == c])) for c in range(16))
print("oracle parity on the first 20000 reads:", ok)
