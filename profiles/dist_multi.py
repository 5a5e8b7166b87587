#!/usr/bin/env python
"""configs[2]-size search only (bench_dist.run), under torchrun; KSSD_SPARSE_SHAPE=wide|narrow to force a CTA shape.
usage: torchrun --nproc-per-node N profiles/dist_multi.py [batches] [baselines 0/1]"""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.distributed as dist
import bench_dist
from public_kssd_b200 import kssd, synth
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1), device=lr)
out = bench_dist.run(ctx, world, rank, dev, 6545.3, batches=int(sys.argv[1]) if len(sys.argv) > 1 else 8,
                     baselines=(len(sys.argv) > 2 and sys.argv[2] == "1"))
if rank == 0:
    keep = {k: v for k, v in out.items() if k not in ("sharding", "timing", "content_check", "oracle_check", "roofline", "e2e", "baseline_note")}
    print(json.dumps({"world": world, "shape": os.environ.get("KSSD_SPARSE_SHAPE", "auto"), **keep}))
if world > 1:
    dist.destroy_process_group()
