#!/usr/bin/env python
"""BASELINE.json configs[4] at reduced scale, end to end on one GPU: synthetic metagenome reads (150 bp FASTQ, Phred+33,
0.5 % substitutions, 0.1 % N, log-normal abundances over a few source genomes) sketched at L3K11 (16 components) with the
abundance filter `-n 2`, then a containment search (-M 1, -D 0.05) against a reference index of the real genomes padded with
synthetic sketches.  Checks that the hits are the source genomes.
usage: python profiles/cfg5_metagenome.py [reads] [padding_refs]   (defaults 4,000,000 reads, 100,000 refs)"""
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from public_kssd_b200 import capi, kssd, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n_pad = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
rl, n_real, glen = 150, 40, 2_000_000
dev = torch.device("cuda", 0)
tab = synth.make_shuf_table(6, 1)
ctx = kssd.Context(11, 6, 3, tab)
NC = ctx.component_num

# ---- reference genomes (clusters of 4, so every source has relatives at 1-10 % divergence) and their sketches
t0 = time.time()
genomes = [b for _, b in synth.cluster_genomes(n_real, glen, seed=11, cluster_size=4)]
fasta = [synth.to_fasta(b, f"g{i}", 80) for i, b in enumerate(genomes)]
refs = ctx.sketch(fasta)
print(f"{n_real} reference genomes x {glen} bp sketched at L3K11: {sum(len(x) for x in refs.ids)} codes in {NC} components "
      f"({time.time() - t0:.1f}s incl. host generation)", flush=True)
# padding sketches: random 28-bit ids, ~glen/4096/16 per component
per = max(glen // 4096 // NC, 1)
R = n_real + n_pad
ref_sizes = np.zeros(R, dtype=np.uint32)
index_list = []
t_ix = 0.0
for c in range(NC):
    a = refs.ids[c]
    ix = refs.index[c]
    pc, pi = synth.synth_sketches(n_pad, per, seed=100 + c, cluster_size=20)
    codes = np.concatenate([a, pc])
    index = np.concatenate([ix, pi[1:] + ix[-1]])
    ref_sizes += np.diff(index).astype(np.uint32)
    index_list.append(ctx.combco2mco(codes, index))
    t_ix += ctx.last_ms(2)
print(f"reference index: {R} refs, 16 components, {sum(i.n_postings for i in index_list)} postings, build {t_ix:.1f} ms (device time)", flush=True)

# ---- reads on the device: 5 sources, log-normal abundances
src_ids = [1, 9, 18, 26, 35]
w = np.exp(np.random.default_rng(5).normal(0, 1, len(src_ids)))
w /= w.sum()
counts = np.maximum((w * n_reads).astype(np.int64), 1)
g = torch.Generator(device=dev); g.manual_seed(7)
lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
parts = []
for s, cnt in zip(src_ids, counts):
    gb = torch.from_numpy(genomes[s]).to(dev)
    st = torch.randint(0, glen - rl, (int(cnt),), generator=g, device=dev)
    parts.append(lut[gb[st[:, None] + torch.arange(rl, device=dev)[None, :]].long()])
bases = torch.cat(parts)
n_reads = bases.shape[0]
bases = bases[torch.randperm(n_reads, generator=g, device=dev)]
err = torch.rand((n_reads, rl), generator=g, device=dev) < 0.005
bases = torch.where(err, lut[torch.randint(0, 4, (n_reads, rl), generator=g, device=dev)], bases)
bases = torch.where(torch.rand((n_reads, rl), generator=g, device=dev) < 0.001, torch.full_like(bases, ord("N")), bases)
qual = torch.randint(35, 74, (n_reads, rl), generator=g, device=dev, dtype=torch.uint8)
hdr = torch.full((n_reads, 12), ord("x"), dtype=torch.uint8, device=dev); hdr[:, 0] = ord("@"); hdr[:, 11] = 10
plus = torch.tensor([43, 10], dtype=torch.uint8, device=dev).expand(n_reads, 2)
nl = torch.full((n_reads, 1), 10, dtype=torch.uint8, device=dev)
rec = torch.cat([hdr, bases, nl, plus, qual, nl], dim=1).contiguous().view(-1)
buf = torch.cat([rec, torch.full((1024,), 10, dtype=torch.uint8, device=dev)])
nbytes = int(rec.numel())
del bases, qual, err, parts
torch.cuda.synchronize()
print(f"metagenome: {n_reads} reads x {rl} bp = {nbytes / 1e9:.2f} GB of FASTQ text, sources {src_ids} at {np.round(w, 3).tolist()}", flush=True)

# ---- Stage I on the reads (-n 2), then the containment search
goff = np.zeros(1, dtype=np.uint64); glen_a = np.array([nbytes], dtype=np.uint64)
best = None
for it in range(3):
    h = ctx.sketch_raw(None, nbytes, goff, glen_a, mode=capi.MODE_FASTQ, Q=0, M=2, device_ptr=buf.data_ptr())
    ms = (ctx.last_ms(0), ctx.last_ms(1))
    best = ms if best is None or ms[1] < best[1] else best
    q = ctx.fetch_sketch(h, 1)
print(f"fastq2co -n 2: line index + scan {best[0]:.2f} ms, whole call {best[1]:.2f} ms = {nbytes / best[1] / 1e6:.0f} GB/s of text, "
      f"{n_reads * rl / best[1] / 1e6:.0f} Gbp/s; query sketch {sum(len(x) for x in q.ids)} codes", flush=True)
qsz = np.array([sum(len(x) for x in q.ids)], dtype=np.uint32)
res = {}
for sparse in (False, True):
    t_best = None
    for it in range(3):
        job = kssd.DistJob(ctx, qsz, ref_sizes, sparse=sparse)
        t_c = 0.0
        for c in range(NC):
            job.accumulate(index_list[c], q.ids[c], q.index[c])
            t_c += 0.0 if sparse else ctx.last_ms(3)
        rows = job.stats(metric=1, dthreshold=0.05)
        t = t_c + ctx.last_ms(4) + (ctx.last_ms(3) if sparse else 0.0)
        t_best = t if t_best is None or t < t_best else t_best
        job.close()
    res[sparse] = rows
    print(f"containment search 1 x {R} ({'sparse' if sparse else 'dense'} job, 16 components): {t_best:.3f} ms device time, "
          f"{len(rows)} rows with AafD <= 0.05", flush=True)
assert res[True].tobytes() == res[False].tobytes()
hits = sorted(int(r) for r in res[True]["ref"])
print("hits:", [(int(r["ref"]), int(r["shared"]), round(float(r["metric"]), 3)) for r in res[True]])
ok = set(src_ids) <= set(hits) and all(h < n_real for h in hits)
print("every source genome found, no padding reference reported:", ok, flush=True)
# parity of the read sketch on a slice against the oracle (checker)
from oracle import oracle as O
orc = O.Ctx(11, 6, 3, tab)
small = rec[: 20000 * (12 + rl + 1 + 2 + rl + 1)].cpu().numpy()
ids, comp = orc.fastq(small, 0, 2)
sk = ctx.sketch_fastq([small], Q=0, M=2)
print("oracle parity on the first 20000 reads:", all(np.array_equal(sk.genome_sets()[0][c], np.sort(ids[comp == c])) for c in range(NC)))
