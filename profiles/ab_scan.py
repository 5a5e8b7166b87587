#!/usr/bin/env python
"""Kernel-only A/B harness: times sketch_fasta32_kernel (CUDA events inside the library) on a resident batch.
usage: KSSD_B200_LIB=path/to/variant.so python profiles/ab_scan.py [genomes]"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from public_kssd_b200 import capi, kssd, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
table = synth.make_shuf_table(6, 1)
ctx = kssd.Context(10, 6, 3, table)
buf, goff, glen = bench.make_batch_device(n, 5_000_000, 7, torch.device("cuda", 0))
torch.cuda.synchronize()
ms = []
for i in range(8):
    h = ctx.sketch_raw(None, int(buf.numel()), goff, glen, device_ptr=buf.data_ptr())
    ms.append(ctx.last_ms(0))
    capi.lib().kssd_sketch_free(h)
b = int(glen.sum())
m = float(np.median(ms[2:]))
print(f"{capi.LIB_PATH.name}: scan {m:.3f} ms  {b / m / 1e6:.1f} GB/s  frac {b / m / 1e6 / 6545.3:.4f}  (min {min(ms):.3f})")
