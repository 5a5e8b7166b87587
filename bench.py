#!/usr/bin/env python
"""bench.py -- headline benchmark of the Kssd hot path on B200 (BASELINE.json configs[1]).

Workload ("step"): Stage I over one batch of synthetic genomes -- 1,000 x 5 Mbp, 80-column FASTA,
50 clusters x 20 members with 0.1-10 % substitutions, L3K10 (k=10, subk=6, drlevel=3) -- i.e. the
sequence -> sketch path (fasta2co + writers) for the whole batch.  metric = sketch Gbp/s.
The all-vs-all Stage II + III pass over the resulting sketches (index, 10^6 shared counts, fused
statistics) is timed in the same run and reported under "dist" (pairs/s), with its own roofline.

  value        Gbp/s with the batch already resident in HBM (kssd_sketch_batch_dev), CUDA events on the
               library's stream, barrier + synchronize on both sides, max over ranks.
  e2e          the same batch through the reference-facing C-ABI with HOST buffers
               (kssd_sketch_batch_host from pinned memory; ids/index fetched back every step).
  roofline     the scan kernel: algorithmic bytes = 1 B per input text byte over the
               kernel time measured with CUDA events inside the library, vs MEASURED_PEAKS.json.
  cpu_baseline the unmodified reference (oracle/_ref/kssd, all host cores) on a bounded sample of the
               same genomes; `--impl reference` runs only that arm.

N > 1 (torchrun): every rank sketches its own 1,000-genome batch (weak scaling, no collective on the
sketch path -- genomes are independent); value = all ranks' bp / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K, SUBK, DRLEVEL = 10, 6, 3
SHUF_SEED = 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dist-scale", action="store_true", help="skip the 10,000 x 100,000 Stage III measurement (N = 1 only)")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ Stage I leg (N = 1 only)")
    ap.add_argument("--no-files", action="store_true", help="skip the end-to-end-from-files leg (N = 1 only)")
    ap.add_argument("--dist-batches", type=int, default=8, help="query batches of the configs[2]-size search")
    ap.add_argument("--ref-sample", type=int, default=0, help="genomes in the CPU sample (0 = auto)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# synthetic batch, generated on the device (5 GB of FASTA text) -- same shape as synth.cluster_genomes
# ------------------------------------------------------------------------------------------------
def make_batch_device(n_genomes, genome_len, seed, device, cluster_size=20, width=80):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    nl = width + 1
    full = genome_len // width
    rem = genome_len - full * width
    body_len = full * nl + (rem + 1 if rem else 0)
    hdr_len = 16
    rec_len = hdr_len + body_len
    stride = (rec_len + 127) // 128 * 128
    buf = torch.full((n_genomes * stride + 1024,), 10, dtype=torch.uint8, device=device)
    goff = np.arange(n_genomes, dtype=np.uint64) * np.uint64(stride)
    glen = np.full(n_genomes, rec_len, dtype=np.uint64)
    anc = None
    for i in range(n_genomes):
        m = i % cluster_size
        if m == 0:
            anc = torch.randint(0, 4, (genome_len,), generator=g, device=device, dtype=torch.uint8)
            b = anc
        else:
            rate = 0.001 * (100.0 ** ((m - 1) / max(cluster_size - 2, 1)))
            hit = torch.rand(genome_len, generator=g, device=device) < rate
            shift = torch.randint(1, 4, (genome_len,), generator=g, device=device, dtype=torch.uint8)
            b = (anc + shift * hit) & 3
        txt = lut[b.long()]
        o = i * stride
        hdr = (">g%06d c%04d" % (i, i // cluster_size)).encode().ljust(hdr_len - 1, b" ")[: hdr_len - 1] + b"\n"
        buf[o:o + hdr_len] = torch.tensor(list(hdr), dtype=torch.uint8, device=device)
        o += hdr_len
        if full:
            blk = buf[o:o + full * nl].view(full, nl)
            blk[:, :width] = txt[: full * width].view(full, width)
            blk[:, width] = 10
            o += full * nl
        if rem:
            buf[o:o + rem] = txt[full * width:]
            buf[o + rem] = 10
    return buf, goff, glen


def fastq_leg(device, peak, n_reads=4_000_000, rl=150):
    """Stage I on FASTQ text at a scaled-down BASELINE.json configs[4] shape: n_reads x 150 bp, Phred+33 qualities, L3K11 (16
    components), -n 2.  Device resident, CUDA events inside the library; a slice of the reads is checked against the oracle."""
    import torch
    from public_kssd_b200 import capi, kssd, synth
    g = torch.Generator(device=device)
    g.manual_seed(5)
    src = torch.randint(0, 4, (2_000_000,), generator=g, device=device, dtype=torch.uint8)
    starts = torch.randint(0, src.numel() - rl, (n_reads,), generator=g, device=device)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    bases = lut[src[starts[:, None] + torch.arange(rl, device=device)[None, :]].long()]
    err = torch.rand((n_reads, rl), generator=g, device=device) < 0.005
    bases = torch.where(err, lut[torch.randint(0, 4, (n_reads, rl), generator=g, device=device)], bases)
    qual = torch.randint(35, 74, (n_reads, rl), generator=g, device=device, dtype=torch.uint8)
    hdr = torch.full((n_reads, 12), ord("x"), dtype=torch.uint8, device=device)
    hdr[:, 0] = ord("@")
    hdr[:, 11] = 10
    num = torch.arange(n_reads, device=device)
    for d in range(10):
        hdr[:, 10 - d] = (48 + (num // (10 ** d)) % 10).to(torch.uint8)
    plus = torch.tensor([43, 10], dtype=torch.uint8, device=device).expand(n_reads, 2)
    nl = torch.full((n_reads, 1), 10, dtype=torch.uint8, device=device)
    rec = torch.cat([hdr, bases, nl, plus, qual, nl], dim=1).contiguous().view(-1)
    buf = torch.cat([rec, torch.full((1024,), 10, dtype=torch.uint8, device=device)])
    nbytes = int(rec.numel())
    del bases, qual, err, hdr
    torch.cuda.synchronize()
    tab6 = synth.make_shuf_table(6, 1)
    ctx = kssd.Context(11, 6, 3, tab6, device=device.index or 0)
    goff = np.zeros(1, dtype=np.uint64)
    glen = np.array([nbytes], dtype=np.uint64)
    ms = []
    for it in range(5):
        h = ctx.sketch_raw(None, nbytes, goff, glen, mode=capi.MODE_FASTQ, Q=0, M=2, device_ptr=buf.data_ptr())
        ms.append((ctx.last_ms(0), ctx.last_ms(1)))
        sk = ctx.fetch_sketch(h, 1)
    k, w = min(m[0] for m in ms[1:]), min(m[1] for m in ms[1:])
    alg = nbytes - n_reads * (rl + 1)          # -Q 0: the quality lines are not part of the algorithm's input
    ok = None
    try:
        from oracle import oracle as O
        orc = O.Ctx(11, 6, 3, tab6)
        small = rec[: 20000 * (12 + rl + 1 + 2 + rl + 1)].cpu().numpy()
        ids, comp = orc.fastq(small, 0, 2)
        sks = ctx.sketch_fastq([small], Q=0, M=2)
        ok = bool(all(np.array_equal(sks.genome_sets()[0][c], np.sort(ids[comp == c])) for c in range(16)))
    except Exception:
        pass
    ctx.close()
    return {"workload": f"{n_reads} reads x {rl} bp, 4-line FASTQ, Phred+33, L3K11 (16 components), -n 2", "text_bytes": nbytes, "bases": n_reads * rl,
            "scan_ms": k, "call_ms": w, "gbp_per_s": n_reads * rl / (w * 1e-3) / 1e9, "text_gb_per_s": nbytes / (k * 1e-3) / 1e9,
            "codes": int(sum(len(x) for x in sk.ids)), "oracle_parity_first_20000_reads": ok,
            "roofline": {"bound": "hbm", "kernel": "nl_index_kernel + sketch_fastq3_kernel (one-pass line index, warp walk)", "achieved": alg / (k * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg, "frac_whole_text": nbytes / (k * 1e-3) / 1e9 / peak,
                         "note": "header, sequence and '+' lines (the quality lines only count with -Q > 0); whole text incl. quality: text_gb_per_s"}}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# the reference arm: the unmodified CPU kssd on the box's host cores
# ------------------------------------------------------------------------------------------------
def files_leg(ctx, host_np, goff, glen, genome_len, ids_e2e, ix_e2e, n_plain=200, n_gz=1000):
    """End to end the way a user runs it: Stage I from FILES (kssd_stage1_files: host threads read into pinned staging, H2D, scan,
    results back).  Plain FASTA and gzip (level 1, the format of the reference's own fixtures) on tmpfs; gzip twice: inflated by
    zlib on the host cores (KSSD_GZ_GPU=0) and inflated on the GPU, one file per thread (the library's choice from 640 files on).
    The ids must equal the resident path's."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    work = Path(tempfile.mkdtemp(prefix="kssd_filesleg_", dir=base))
    out = {}
    try:
        n_plain = min(n_plain, len(glen))
        n_gz = min(n_gz, len(glen))
        plain, gz = [], []
        for i in range(n_plain):
            g = host_np[int(goff[i]):int(goff[i]) + int(glen[i])]
            f = work / f"g{i:05d}.fasta"
            f.write_bytes(g.tobytes())
            plain.append(f)

        def write_gz(i):
            co = zlib.compressobj(1, zlib.DEFLATED, 31)
            raw = host_np[int(goff[i]):int(goff[i]) + int(glen[i])].tobytes()
            fz = work / f"g{i:05d}.fasta.gz"
            fz.write_bytes(co.compress(raw) + co.flush())
            return fz
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 8) as ex:      # (zlib releases the GIL)
            gz = list(ex.map(write_gz, range(n_gz)))
        saved = os.environ.get("KSSD_GZ_GPU")
        for name, paths, gz_gpu in (("plain", plain, None), ("gz", gz, "0"), ("gz_gpu", gz, "1")):
            if gz_gpu is not None:
                os.environ["KSSD_GZ_GPU"] = gz_gpu
            best, sk, best_bb = None, None, 0
            for bb in ((0, 256 << 20, 128 << 20) if name == "plain" else (0,)):      # staging batch: the library's default (1 GiB) or smaller, so
                for _ in range(3):                                                    # that reading batch b + 1 overlaps the H2D + scan of batch b
                    sk, t = ctx.sketch_files(paths, batch_bytes=bb)
                    if best is None or t["total_s"] < best["total_s"]:
                        best, best_bb = t, bb
            n = len(paths)
            bp = n * genome_len
            same = bool(np.array_equal(sk.ids[0], ids_e2e[:int(ix_e2e[n])]))
            on_disk = int(sum(f.stat().st_size for f in paths))
            out[name] = {"value": bp / best["total_s"] / 1e9, "unit": "Gbp/s", "files": n, "text_bytes": best["bytes"], "bytes_on_disk": on_disk,
                         "total_s": best["total_s"], "read_s": best["read_s"], "gpu_s": best["gpu_s"], "text_gb_per_s": best["bytes"] / best["total_s"] / 1e9,
                         "batches": best["batches"], "batch_bytes": best_bb or ("default (1 GiB)" if not best.get("gz_on_gpu") else "16 GiB decoded"),
                         "matches_device_path": same}
            if name != "plain":
                out[name].update({"inflate": "GPU (csrc/inflate.cuh, one file per thread)" if best.get("gz_on_gpu") else "zlib on the host cores",
                                  "h2d_and_inflate_s": best.get("gz_gpu_s", 0.0)})
        if saved is None:
            os.environ.pop("KSSD_GZ_GPU", None)
        else:
            os.environ["KSSD_GZ_GPU"] = saved
        out["note"] = ("kssd_stage1_files on tmpfs files, best of 3 calls, every host core reading (and, for 'gz', inflating with zlib); gz = gzip -1; "
                       "ids compared with the resident path's")
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return out


def reference_run(host_genomes, shuf_table, steps, warmup, cores, with_dist=True, files_fn=None):
    """host_genomes: list of uint8 arrays (FASTA text).  Times `kssd dist` stage I per step on tmpfs;
    stage II / III once.  Returns dict."""
    from oracle import oracle as O
    if not O.REF_BIN.exists():
        O.build()
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    work = Path(tempfile.mkdtemp(prefix="kssd_refbench_", dir=base))
    try:
        shuf = work / "L3K10.shuf"
        O.write_shuf_file(shuf, 4242, K, SUBK, DRLEVEL, shuf_table)
        ind = work / "in"
        ind.mkdir()
        total_bp = 0
        total_bytes = 0
        for i, g in enumerate(host_genomes):
            (ind / f"g{i:05d}.fasta").write_bytes(g.tobytes())
            total_bytes += g.size
            total_bp += int(np.count_nonzero((g == 65) | (g == 67) | (g == 71) | (g == 84)))
        times = []
        for it in range(warmup + steps):
            out = work / f"sk{it}"
            t0 = time.perf_counter()
            r = subprocess.run([str(O.REF_BIN), "dist", "-p", str(cores), "-L", str(shuf), "-o", str(out), str(ind)],
                               capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if r.returncode != 0 or not (out / "cofiles.stat").exists():
                raise RuntimeError("reference stage I failed: " + r.stderr[-300:])
            if it >= warmup:
                times.append(dt)
            if it < warmup + steps - 1:
                shutil.rmtree(out)
        res = {"sketch_s": float(np.mean(times)), "bp": total_bp, "bytes": total_bytes, "genomes": len(host_genomes)}
        if files_fn is not None:      # the GPU library's own file path on the very same files
            res["files"] = files_fn(sorted(ind.glob("*.fasta")))
        if with_dist:
            t0 = time.perf_counter()
            r = subprocess.run([str(O.REF_BIN), "dist", "-p", str(cores), "-o", str(out), str(out)], capture_output=True, text=True)
            res["index_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            r = subprocess.run([str(O.REF_BIN), "dist", "-p", str(cores), "-r", str(out), "-o", str(work / "dist"), str(out)],
                               capture_output=True, text=True)
            res["dist_s"] = time.perf_counter() - t0
            res["dist_pairs"] = len(host_genomes) ** 2
            res["dist_ok"] = (work / "dist" / "distance.out").exists()
        return res
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    from public_kssd_b200 import synth
    table = synth.make_shuf_table(SUBK, SHUF_SEED)
    workload = f"{args.genomes} x {args.genome_len} bp synthetic bacterial genomes, 80-col FASTA, L{DRLEVEL}K{K} (subk {SUBK})"
    config = {"workload": workload, "configs_index": 1, "genomes_per_gpu": args.genomes, "genome_len": args.genome_len,
              "k": K, "subk": SUBK, "drlevel": DRLEVEL, "l2": "inputs (5 GB) larger than L2, no flush needed",
              "sharding": "genomes across ranks, no collective"}

    if args.impl == "reference":
        if rank != 0:
            return
        n_s = args.ref_sample or max(16, min(4 * cores, 256))
        n_s = min(n_s, args.genomes)
        # the same bytes as the b200 arm's first n_s genomes: its device generator is sequential per genome, so a batch of n_s
        # genomes with the same seed is a prefix of the 1000-genome batch (host generator only where there is no GPU at all)
        same_bytes = False
        try:
            import torch
            if torch.cuda.is_available():
                dev0 = torch.device("cuda", 0)
                buf0, goff0, glen0 = make_batch_device(n_s, args.genome_len, 20260101, dev0)
                hb = buf0.cpu().numpy()
                gens = [hb[int(goff0[i]):int(goff0[i]) + int(glen0[i])].copy() for i in range(n_s)]
                del buf0, hb
                same_bytes = True
        except Exception:
            same_bytes = False
        if not same_bytes:
            gens = [synth.to_fasta(b, n, 80) for n, b in synth.cluster_genomes(n_s, args.genome_len, seed=20260101, cluster_size=20)]
        config["reference_sample_is_prefix_of_b200_batch"] = same_bytes
        r = reference_run(gens, table, args.steps, args.warmup, cores)
        val = r["bp"] / r["sketch_s"] / 1e9
        sample = f"{n_s} of the workload's genomes ({r['bytes'] / 1e6:.0f} MB FASTA on tmpfs), kssd dist -p {cores}, per step"
        line = {"impl": "reference", "metric": "sketch_gbp_per_s", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["sketch_s"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": sample},
                "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "dist": {"pairs_per_s": r.get("dist_pairs", 0) / max(r.get("dist_s", 1e-9), 1e-9), "pairs": r.get("dist_pairs"),
                         "dist_s": r.get("dist_s"), "index_s": r.get("index_s"), "note": "stage III incl. distance.out text; stage II separately"}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from public_kssd_b200 import capi, hostfmt, kssd
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = kssd.Context(K, SUBK, DRLEVEL, table, device=local_rank, shuf_id=4242)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    buf, goff, glen = make_batch_device(args.genomes, args.genome_len, 20260101 + rank, dev)
    torch.cuda.synchronize()
    nbytes = int(buf.numel())
    text_bytes = int(glen.sum())
    bp = args.genomes * args.genome_len
    L = capi.lib()

    def sketch_dev():
        return ctx.sketch_raw(None, nbytes, goff, glen, device_ptr=buf.data_ptr())

    # ---- correctness guard inside the bench: a sample of genomes against the oracle (not timed) ----
    clocks = ClockSampler(local_rank)
    clocks.start()                      # sampled through both timed regions (resident steps and end-to-end steps)
    for _ in range(args.warmup):
        h = sketch_dev()
        L.kssd_sketch_free(h)
    scan_ms = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.kssd_kernel_launch_count()
    e0.record(stream)
    last = None
    for _ in range(args.steps):
        if last is not None:
            L.kssd_sketch_free(last)
        last = sketch_dev()
        scan_ms.append(ctx.last_ms(0))
    e1.record(stream)
    barrier()
    gpu_launches = int(L.kssd_kernel_launch_count() - l0)
    step_ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([step_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms_max = float(t.item())
    value = world * bp / (step_ms_max * 1e-3) / 1e9

    # ---- Stage II + III on the sketches just made (all-vs-all), device resident ----
    sk_h = last
    ids_p, idx_p = capi.C.c_void_p(), capi.C.c_void_p()
    capi.check(L.kssd_sketch_dev_ptrs(sk_h, 0, capi.C.byref(ids_p), capi.C.byref(idx_p)))
    n_codes = int(capi.check(L.kssd_sketch_count(sk_h, 0)))
    index_host = np.empty(args.genomes + 1, dtype=np.uint64)
    capi.check(L.kssd_sketch_fetch(sk_h, 0, None, capi.ptr(index_host, capi.C.c_uint64), None, None))
    sizes = np.diff(index_host).astype(np.uint32)
    dist_info = {}
    pairs = args.genomes * args.genomes
    peak, peak_src = peaks()
    if world == 1:
        ix_ms, ct_ms, st_ms = [], [], []
        for it in range(3):
            ixh = capi.C.c_void_p()
            capi.check(L.kssd_index_build_dev(ctx._h, ids_p, idx_p, args.genomes, n_codes, capi.C.byref(ixh)))
            ix_ms.append(ctx.last_ms(2))
            ix = kssd.Index(ctx, ixh)
            job = kssd.DistJob(ctx, sizes, sizes)
            job.accumulate_dev(ix, ids_p.value, idx_p.value, n_codes)
            ct_ms.append(ctx.last_ms(3))
            nrows = job.stats(fetch=False)
            st_ms.append(ctx.last_ms(4))
            if it == 2:
                ct = job.counts()
                dist_info["shared_total"] = int(ct.sum(dtype=np.uint64))
                dist_info["diag_ok"] = bool(np.array_equal(np.diag(ct), sizes))
                dist_info["rows"] = int(nrows)
                # end to end like the reference's Stage III: statistics + the distance.out text on the host.  Main line: the GPU writes
                # the text from the rows on the device (kssd_dist_format_text); beside it the rows fetched and formatted by the host threads
                names = [f"g{i}.fna" for i in range(args.genomes)]
                name_block = hostfmt._names_block(names)
                for rep in range(2):
                    t0 = time.perf_counter()
                    job.stats(fetch=False)
                    text = job.distance_out_view(name_block, name_block, 0, 2)
                    dist_info["text_s"] = time.perf_counter() - t0
                dist_info["text_kernels_ms"] = ctx.last_ms(5)
                t0 = time.perf_counter()
                rows_host = job.stats()
                text_host = hostfmt.format_distance_out(rows_host, names, names, 0, 2)
                dist_info["text_host_formatter_s"] = time.perf_counter() - t0
                dist_info["text_bytes"] = len(text)
                dist_info["text_gpu_equals_host"] = bool(bytes(text) == text_host)
                del text, text_host, rows_host
            job.close(); ix.close()
        d_ct, d_st, d_ix = float(np.min(ct_ms)), float(np.min(st_ms)), float(np.min(ix_ms))
        dist_bytes = 4 * n_codes + 8 * n_codes + 4 * dist_info.get("shared_total", 0) + 4 * pairs
        dist_info.update({"metric": "dist_pairs_per_s", "pairs": pairs, "pairs_per_s": pairs / ((d_ct + d_st) * 1e-3), "count_ms": d_ct,
                          "stats_ms": d_st, "index_ms": d_ix,
                          "pairs_per_s_incl_text": pairs / ((d_ct) * 1e-3 + dist_info.get("text_s", 0.0)),
                          "text_note": "text_s = statistics kernel + distance.out written by the GPU (fmt_exact.cuh: glibc's %.6lf / %E in integer arithmetic) + D2H of the text into the context's pinned buffer; "
                                       "text_host_formatter_s = statistics + D2H of the rows + snprintf on every host thread",
                          "stats_rows": "all Q*R rows (Jaccard, MashD, P-value, FDR, CIs), fp64",
                          "roofline": {"bound": "hbm", "achieved": dist_bytes / (d_ct * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": dist_bytes / (d_ct * 1e-3) / 1e9 / peak, "traffic": None,
                                       "note": "count kernel; bytes = 4*Nq + 8*Nq + 4*P + 4*Q*R (SURVEY.md s8d); 10^6 cells is launch-latency scale -- "
                                               "profiles/r1_dist_ncu_summary.md has the 10^9-pair run (0.52 of peak)"}})
    else:
        dist_info.update({"metric": "dist_pairs_per_s", "note": "at N > 1 the search is measured at configs[2] size only (configs2_scale below)"})

    # ---- end to end through the C-ABI with host buffers ----
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(buf)
    torch.cuda.synchronize()
    host_np = host.numpy()
    e2e_steps = max(2, min(args.steps, 5))
    d2h = 0

    def sketch_host():
        nonlocal d2h
        h = ctx.sketch_raw(host_np, nbytes, goff, glen)
        n = int(capi.check(L.kssd_sketch_count(h, 0)))
        ids = np.empty(n, dtype=np.uint32)
        ixx = np.empty(args.genomes + 1, dtype=np.uint64)
        capi.check(L.kssd_sketch_fetch(h, 0, capi.ptr(ids, capi.C.c_uint32), capi.ptr(ixx, capi.C.c_uint64), None, None))
        L.kssd_sketch_free(h)
        d2h = ids.nbytes + ixx.nbytes
        return ids, ixx

    sketch_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ids_e2e, ix_e2e = sketch_host()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * bp / float(te.item()) / 1e9
    clk = clocks.stop()
    # what the box's host -> device path alone does with the same bytes (every rank copies its pinned batch at the same time):
    # the ceiling of any end-to-end number from host memory
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    buf.copy_(host, non_blocking=True)
    h1.record()
    barrier()
    th = torch.tensor([h0.elapsed_time(h1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    h2d_ms = float(th.item())

    # the device-resident and host paths must agree with each other
    ids_dev = np.empty(n_codes, dtype=np.uint32)
    capi.check(L.kssd_sketch_fetch(sk_h, 0, capi.ptr(ids_dev, capi.C.c_uint32), None, None, None))
    same = bool(np.array_equal(ids_dev, ids_e2e) and np.array_equal(index_host, ix_e2e))
    L.kssd_sketch_free(sk_h)

    scan = float(np.median(scan_ms))
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "r2_sketch_traffic.json"          # dram bytes of this kernel from one `ncu --set full` capture
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            if tj.get("kernel", "").startswith("sketch_fasta3"):
                traffic, traffic_src = tj["traffic_bytes_per_launch"] / tj["algorithmic_bytes_per_launch"] * text_bytes, "profiles/r2_sketch_traffic.json (ncu dram__bytes_read+write per algorithmic byte, 200-genome launch, scaled to this launch)"
        except Exception:
            pass
    roof = {"bound": "hbm", "achieved": text_bytes / (scan * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": text_bytes / (scan * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": "sketch_fasta3_kernel<3, true>", "kernel_ms": scan, "algorithmic_bytes": text_bytes}

    # (after the end-to-end leg: the host-side generation below must not disturb where its pinned buffer lives)
    if not args.no_dist_scale and args.genomes >= 1000:
        # the second headline metric at BASELINE.json configs[2] size, at every N: 100,000 reference sketches sharded over the
        # ranks, batches of 10,000 queries (10^9 pairs each), content checked against one GPU and the oracle (bench_dist.py)
        try:
            import bench_dist
            dist_info["configs2_scale"] = bench_dist.run(ctx, world, rank, dev, peak, batches=args.dist_batches)
            if world > 1:
                dist_info["pairs_per_s"] = dist_info["configs2_scale"]["pairs_per_s"]
                dist_info["pairs"] = dist_info["configs2_scale"]["pairs_per_batch"] * dist_info["configs2_scale"]["batches"]
        except Exception as ex:  # must not take the headline measurement down
            import traceback
            dist_info["configs2_scale"] = {"failed": str(ex), "trace": traceback.format_exc()[-800:]}

    fastq_info = None
    if world == 1 and not args.no_fastq:
        try:
            fastq_info = fastq_leg(dev, peak)
        except Exception as ex:  # must not take the headline measurement down
            fastq_info = {"failed": str(ex)}

    files_info = None
    if world == 1 and not args.no_files:
        try:
            files_info = files_leg(ctx, host_np, goff, glen, args.genome_len, ids_e2e, ix_e2e)
        except Exception as ex:  # must not take the headline measurement down
            files_info = {"failed": str(ex)}

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            n_s = args.ref_sample or max(16, min(2 * cores, 128))
            n_s = min(n_s, args.genomes)
            gens = []
            for i in range(n_s):
                o = int(goff[i])
                gens.append(host_np[o:o + int(glen[i])].copy())
            def files_fn(paths):
                best = None
                for _ in range(3):                                   # the first call pins the staging buffers
                    sk_f, t = ctx.sketch_files(paths)
                    best = t if best is None or t["total_s"] < best["total_s"] else best
                same_f = bool(np.array_equal(sk_f.ids[0], ids_e2e[:int(ix_e2e[len(paths)])]))
                return {"total_s": best["total_s"], "bytes": best["bytes"], "matches_device_path": same_f}
            r = reference_run(gens, table, 1, 0, cores, files_fn=files_fn)
            # parity of the sampled genomes against the oracle restatement (checker only)
            from oracle import oracle as O
            orc = O.Ctx(K, SUBK, DRLEVEL, table)
            ok = True
            n_par = min(32, n_s)
            for i in range(n_par):
                ids_o, _ = orc.fasta(gens[i])
                ok &= bool(np.array_equal(np.sort(ids_o), ids_e2e[int(ix_e2e[i]):int(ix_e2e[i + 1])]))
            cpu_baseline = {"value": r["bp"] / r["sketch_s"] / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                            "sample": f"first {n_s} genomes of rank 0's batch ({r['bytes'] / 1e6:.0f} MB FASTA on tmpfs), one `kssd dist -p {cores}` run",
                            "sketch_s": r["sketch_s"], "index_s": r.get("index_s"), "dist_s": r.get("dist_s"),
                            "dist_pairs_per_s": r.get("dist_pairs", 0) / max(r.get("dist_s", 1e-9), 1e-9), "oracle_parity_on_sample": ok, "oracle_parity_genomes": n_par,
                            "gpu_same_files": {"value": r["bp"] / r["files"]["total_s"] / 1e9, "unit": "Gbp/s", "total_s": r["files"]["total_s"],
                                               "matches_device_path": r["files"]["matches_device_path"],
                                               "note": "kssd_stage1_files on the same tmpfs files the reference just read: host threads read into "
                                                       "pinned staging, H2D, scan, results back on the host (best of 3 calls)"}}
        except Exception as ex:  # the baseline must not take the measurement down
            cpu_baseline = {"value": None, "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {"metric": "sketch_gbp_per_s", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic", "config": config, "clocks": clk, "library": capi.library_provenance(),
                "e2e": {"value": e2e_val, "unit": "Gbp/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": float(te.item()) * 1e3, "matches_device_path": same,
                        "h2d_copy_alone_ms": h2d_ms, "h2d_aggregate_gb_per_s": world * nbytes / (h2d_ms * 1e-3) / 1e9,
                        "h2d_share_of_step": h2d_ms / (float(te.item()) * 1e3),
                        "from_files": files_info},
                "gpu_launches": gpu_launches, "roofline": roof, "cpu_baseline": cpu_baseline, "dist": dist_info,
                "fastq": fastq_info,
                "sketch": {"codes": n_codes, "text_bytes": text_bytes, "scan_kernel_ms": scan, "step_ms": step_ms_max}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
