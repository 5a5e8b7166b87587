"""bench.py's second metric at BASELINE.json configs[2] size: 100,000 reference sketches x ~1,220 codes searched with batches of
10,000 queries (10^9 pairs per batch), at every N.

Sharding: the ranks form a Gr x Gq grid -- Gr reference shards by GENOME range (each rank's rows are final as
they are, no reduction) times Gq query groups.  The per-query part of the search (code lookups, clearing and reading out the
query's table) does not shrink when only the references are split, and a 100 000-genome index is 0.5 GB, so the main line is
1 x N: every rank holds the whole index and searches its share of each batch's queries, which is resident in its HBM when the
clock starts (like the genomes of the sketch metric); every rank runs the sparse Stage III job (count + filter + list +
statistics, no Q x R matrix) and there is no exchange at all.  The same loop with rank 0 holding each batch and sending every
rank the codes of its group point to point (NCCL), one batch ahead of the compute, is timed next to it.  The grids with reference shards (2 x N/2, N x 1 -- what an
index beyond one GPU's memory needs), the north-star variant (index sharded by code range + NCCL reduce-scatter of dense partial matrices) and the
peer-memory variant are timed next to it as the baselines they are.  Times are CUDA events on the library stream, max over ranks.
"""
from __future__ import annotations

import time

import numpy as np

N_REF, N_QRY, CODES, BATCHES = 100_000, 10_000, 1220, 8


def _checksum(rows) -> tuple[int, int]:
    """order-independent digest of (qry, ref, shared) rows"""
    if len(rows) == 0:
        return 0, 0
    q = rows["qry"].astype(np.uint64)
    r = rows["ref"].astype(np.uint64)
    s = rows["shared"].astype(np.uint64)
    with np.errstate(over="ignore"):
        h = (q * np.uint64(0x9E3779B97F4A7C15)) ^ (r * np.uint64(0xC2B2AE3D27D4EB4F)) ^ (s * np.uint64(0x165667B19E3779F9))
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
    return int(h.sum(dtype=np.uint64)), int(len(rows))


def run(ctx, world: int, rank: int, dev, peak_gbs: float, n_ref: int = N_REF, n_qry: int = N_QRY, batches: int = BATCHES,
        baselines: bool = True) -> dict:
    import torch
    import torch.distributed as dist
    from public_kssd_b200 import kssd, parallel, synth

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- data: rank 0 generates (cluster / mutation model of SURVEY.md s8d), everybody gets the references; the query
    # batches stay on rank 0 until they are broadcast inside the timed loop
    t0 = time.perf_counter()
    meta = torch.zeros(2 + 2 * batches, dtype=torch.int64, device=dev)
    q_gen = []
    if rank == 0:      # (synth_sketches_torch: the numpy generator's sketches element for element, made on the GPU)
        t_rc, t_ri = synth.synth_sketches_torch(n_ref, CODES, seed=5, device=dev, cluster_size=20)
        for b in range(batches):
            q_gen.append(synth.synth_sketches_torch(n_qry, CODES, seed=5, device=dev, cluster_size=2, member_seed=None if b == 0 else 1000 + b))
        meta[0], meta[1] = int(t_rc.numel()), n_ref
        for b, (qc, qi) in enumerate(q_gen):
            meta[2 + 2 * b], meta[3 + 2 * b] = int(qc.numel()), int(qi.numel())
    if world > 1:
        dist.broadcast(meta, 0)
    m = meta.cpu().numpy()
    n_codes = int(m[0])
    if rank != 0:
        t_rc = torch.empty(n_codes, dtype=torch.int32, device=dev)
        t_ri = torch.empty(n_ref + 1, dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(t_rc, 0)
        dist.broadcast(t_ri, 0)
    ref_index_host = t_ri.cpu().numpy().view(np.uint64)
    ref_sizes = np.diff(ref_index_host).astype(np.uint32)
    q_dev = []
    for b in range(batches):
        if rank == 0:
            q_dev.append(q_gen[b])
        else:
            q_dev.append((torch.empty(int(m[2 + 2 * b]), dtype=torch.int32, device=dev), torch.empty(int(m[3 + 2 * b]), dtype=torch.int64, device=dev)))
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0

    cm = (n_ref * n_qry) & 0xFFFFFFFF
    # sketch sizes and per-genome extents are metadata (cofiles.stat / combco.index): everybody has them before the search;
    # only the query CODES travel inside the timed loop
    q_index_host = []
    for b in range(batches):
        if world > 1:
            dist.broadcast(q_dev[b][1], 0)
        q_index_host.append(q_dev[b][1].cpu().numpy().view(np.uint64).copy())

    def measure(Gr: int, nb: int) -> dict:
        """ranks as a Gr x Gq grid: rank r holds reference shard r % Gr (genome range) and searches query group r // Gr"""
        Gq = world // Gr
        gr, gq = rank % Gr, rank // Gr
        g = parallel.genome_shard(n_ref, Gr, gr)
        lo, hi = g.start, g.stop
        a, bnd = int(ref_index_host[lo]), int(ref_index_host[hi])
        shard_codes = t_rc[a:bnd]
        shard_index = (t_ri[lo:hi + 1] - t_ri[lo]).contiguous()
        ix_ms, index = [], None
        for _ in range(2):
            if index is not None:
                index.close()
            torch.cuda.synchronize()
            index = ctx.combco2mco_dev(shard_codes.data_ptr(), shard_index.data_ptr(), hi - lo, bnd - a)
            ix_ms.append(ctx.last_ms(2))
        shard_sizes = ref_sizes[lo:hi]
        qg = parallel.genome_shard(n_qry, Gq, gq)
        qlo, qhi = qg.start, qg.stop
        # per batch: this rank's query rows, their local index (device), sizes (host), and where their codes sit in the batch
        plan = []
        for b in range(nb):
            qi = q_index_host[b]
            c0, c1 = int(qi[qlo]), int(qi[qhi])
            li = torch.from_numpy((qi[qlo:qhi + 1] - qi[qlo]).astype(np.int64)).to(dev)
            buf = q_dev[b][0][c0:c1] if rank == 0 else torch.empty(c1 - c0, dtype=torch.int32, device=dev)
            plan.append({"li": li, "qsz": np.diff(qi[qlo:qhi + 1]).astype(np.uint32), "buf": buf, "c0": c0, "c1": c1})

        def send(b):
            """rank 0 holds batch b: every rank gets the codes of ITS query group (point to point over NVLink)"""
            if world == 1:
                return None
            ops = []
            if rank == 0:
                qi = q_index_host[b]
                for r in range(1, world):
                    rg = parallel.genome_shard(n_qry, Gq, r // Gr)
                    ops.append(dist.P2POp(dist.isend, q_dev[b][0][int(qi[rg.start]):int(qi[rg.stop])], r))
            else:
                ops.append(dist.P2POp(dist.irecv, plan[b]["buf"], 0))
            return dist.batch_isend_irecv(ops) if ops else None

        def wait(w):
            if w is not None:
                for x in w:
                    x.wait()
                torch.cuda.current_stream().synchronize()

        def search(b, fetch=False, phases=False):
            pl = plan[b]
            job = kssd.DistJob(ctx, pl["qsz"], shard_sizes, sparse=True)
            job.accumulate_dev(index, pl["buf"].data_ptr(), pl["li"].data_ptr(), int(pl["buf"].numel()))
            rows = job.stats(skip_zero=1, fetch=fetch, cmprsn_num=cm)
            t = (ctx.last_ms(3), ctx.last_ms(4)) if phases else None        # (reading the second one waits for the rows kernel)
            job.close()
            return rows, t

        ph = []
        for b in range(nb):                                          # warm-up, and the kernel times of an unpipelined pass
            wait(send(b))
            ph.append(search(b, phases=True)[1])
        def timed(scatter: bool):
            """all batches, queued back to back; scatter: rank 0 sends batch b + 1 while batch b is searched"""
            best = None
            for rep in range(2):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                nxt = send(0) if scatter else None
                jobs = []
                for b in range(nb):
                    if scatter:
                        wait(nxt)
                        nxt = send(b + 1) if b + 1 < nb else None        # the next batch travels while this one is searched
                    pl = plan[b]
                    job = kssd.DistJob(ctx, pl["qsz"], shard_sizes, sparse=True)
                    job.accumulate_dev(index, pl["buf"].data_ptr(), pl["li"].data_ptr(), int(pl["buf"].numel()))
                    job.stats_async(skip_zero=1, cmprsn_num=cm)           # queued: the host runs ahead of the GPU
                    jobs.append(job)
                nrows = 0
                for job in jobs:                                         # every batch's rows are on the device when this returns
                    nrows += int(job.stats_wait())
                e1.record(stream)
                barrier()
                for job in jobs:
                    job.close()
                ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if best is None or float(ms.item()) < best[0]:
                    best = (float(ms.item()), nrows)
            return best

        # value: every rank's share of every batch is resident in its HBM when the clock starts (the warm-up pass above put it
        # there), as in the sketch metric; next to it, the same loop with rank 0 scattering the query codes inside the timed region
        best = timed(False)
        scatter_ms = timed(True)[0] if world > 1 else None
        total_ms, nrows = best
        tr = torch.tensor([nrows], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(tr)
        # content of batch 0 over all ranks: additive digest (mod 2^64) + row count
        rows0, _ = search(0, fetch=True)
        rows0["ref"] += lo
        rows0["qry"] += qlo
        h, n = _checksum(rows0)
        agg = torch.tensor([int(np.uint64(h).astype(np.int64)), n], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(agg)
        index.close()
        return {"grid": {"ref_shards": Gr, "query_groups": Gq}, "batches": nb, "ms_total": total_ms, "ms_per_batch": total_ms / nb,
                "pairs_per_s": nb * n_qry * n_ref / (total_ms * 1e-3), "printed_rows": int(tr.item()), "index_ms_per_rank": float(min(ix_ms)),
                "rank0_kernel_ms_per_batch_unpipelined": {"count_list": float(np.mean([p[0] for p in ph])), "rows": float(np.mean([p[1] for p in ph]))},
                "ms_per_batch_with_rank0_scatter_in_loop": None if scatter_ms is None else scatter_ms / nb,
                "digest": int(agg[0].item()) & 0xFFFFFFFFFFFFFFFF, "rows_batch0": int(agg[1].item())}

    # Which grid?  The per-query part of the search (code lookups, clearing and reading out the query's table) does not shrink
    # when only the references are sharded, so the ranks split the QUERIES first: a 100 000-genome index is 0.5 GB of the
    # 180 GB a rank has, every rank holds all of it (1 x N) and no rank repeats another's lookups.  Reference shards (2 x N/2,
    # N x 1) are what an index beyond one GPU's memory needs; they are timed next to it on fewer batches.
    grids = [1] if world == 1 else sorted({1, 2, world} & {g for g in (1, 2, world) if world % g == 0})
    res = {1: measure(1, batches)}
    for g in grids[1:]:
        res[g] = measure(g, min(batches, 4))
    Gr = 1
    main = res[Gr]
    out = {"pairs_per_batch": n_qry * n_ref, "refs": n_ref, "queries_per_batch": n_qry, "ref_postings": n_codes,
           "sharding": (f"ranks as a {Gr} x {world // Gr} grid: every rank holds the whole reference index, each batch's queries are split over "
                        f"{world // Gr} group(s) and resident on their rank when the clock starts; sparse job per rank, no exchange, no reduction "
                        f"(ms_per_batch_with_rank0_scatter_in_loop: the same loop with rank 0 sending every rank its query codes point to point, "
                        f"one batch ahead, inside the timed region)"),
           "timing": "CUDA events on the library stream around all batches (host gaps included), max over ranks", "generation_s": gen_s}
    out.update({k: v for k, v in main.items() if k not in ("digest", "rows_batch0")})
    if world > 1:
        out["other_grids"] = []
        for g in grids[1:]:
            alt = {k: v for k, v in res[g].items() if k not in ("digest", "rows_batch0", "printed_rows")}
            alt["content_same_as_main"] = bool(res[g]["digest"] == main["digest"] and res[g]["rows_batch0"] == main["rows_batch0"])
            out["other_grids"].append(alt)

    # ---- content: batch 0's rows over all ranks against one GPU holding the whole index (rank 0), and against the oracle
    if rank == 0:
        tq, ti = q_dev[0]
        qsz0 = np.diff(q_index_host[0]).astype(np.uint32)
        full = ctx.combco2mco_dev(t_rc.data_ptr(), t_ri.data_ptr(), n_ref, n_codes)
        dj = kssd.DistJob(ctx, qsz0, ref_sizes)                    # dense job on one GPU: another kernel, the whole matrix
        dj.accumulate_dev(full, tq.data_ptr(), ti.data_ptr(), int(tq.numel()))
        dense_count_ms = ctx.last_ms(3)
        want = dj.stats(skip_zero=1, cmprsn_num=cm)
        dense_stats_ms = ctx.last_ms(4)
        hw, nw = _checksum(want)
        out["content_ok"] = bool(main["digest"] == (hw & 0xFFFFFFFFFFFFFFFF) and main["rows_batch0"] == nw)
        out["content_check"] = "batch 0: digest + count of (qry, ref, shared) rows over all ranks == the dense job of one GPU holding the whole index"
        # oracle (checker): brute force through the CPU restatement on the first 16 queries x first 2000 references
        try:
            from oracle import oracle as O
            O.build()
            nq_o, nr_o = 16, 2000
            rch, rih = t_rc[: int(ref_index_host[nr_o])].cpu().numpy().view(np.uint32), ref_index_host[: nr_o + 1]
            qih = ti[: nq_o + 1].cpu().numpy().view(np.uint64)
            qch = tq[: int(qih[nq_o])].cpu().numpy().view(np.uint32)
            uc, uo, gids = O.csr_from_combco(rch, rih)
            exp = O.dist_counts(qch, qih, uc, uo, gids, nr_o)
            sel = want[(want["qry"] < nq_o) & (want["ref"] < nr_o)]
            got_m = np.zeros((nq_o, nr_o), dtype=np.uint32)
            got_m[sel["qry"], sel["ref"]] = sel["shared"]
            ok = bool(np.array_equal(got_m, exp))
            for row in sel[:: max(1, len(sel) // 64)]:
                keep, v = O.output_ctrl(ref_sizes[row["ref"]], qsz0[row["qry"]], row["shared"], 0, 0, 20, 6, 1.0, cm)
                ok &= bool(keep and abs(row["dist"] - v[1]) <= 1e-6 * abs(v[1]) and abs(row["metric"] - v[0]) <= 1e-6 * abs(v[0]))
            out["oracle_ok"] = ok
            out["oracle_check"] = f"shared counts of the first {nq_o} queries x {nr_o} references and 64 of their statistics rows against the CPU oracle"
        except Exception as ex:      # the checker must not take the measurement down
            out["oracle_ok"] = None
            out["oracle_check"] = f"failed: {ex}"
        # ---- one GPU, whole index: the dense job's count kernel against the HBM roofline, atomics/s, end to end from host buffers
        if world == 1:
            P = int(want["shared"].sum(dtype=np.uint64))
            nq_codes = int(tq.numel())
            alg = 4 * nq_codes + 16 * nq_codes + 4 * P + 4 * n_qry * n_ref
            cts, sts = [dense_count_ms], [dense_stats_ms]
            for _ in range(2):
                dj2 = kssd.DistJob(ctx, qsz0, ref_sizes)
                dj2.accumulate_dev(full, tq.data_ptr(), ti.data_ptr(), nq_codes)
                cts.append(ctx.last_ms(3))
                dj2.stats(skip_zero=1, fetch=False, cmprsn_num=cm)
                sts.append(ctx.last_ms(4))
                dj2.close()
            c_ms, s_ms = float(min(cts)), float(min(sts))
            out["dense_job"] = {"count_ms": c_ms, "stats_ms": s_ms, "pairs_per_s": n_qry * n_ref / ((c_ms + s_ms) * 1e-3), "increments": P,
                                "atomics_per_s": P / (c_ms * 1e-3)}
            traffic = None
            try:
                import json
                from pathlib import Path
                tj = json.loads((Path(__file__).resolve().parent / "profiles" / "r2_dist_traffic.json").read_text())
                traffic = tj.get("traffic_bytes_per_launch")
            except Exception:
                pass
            out["roofline"] = {"bound": "hbm", "kernel": "dist_count_rows_kernel", "achieved": alg / (c_ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                               "frac": alg / (c_ms * 1e-3) / 1e9 / peak_gbs, "traffic": traffic, "algorithmic_bytes": alg,
                               "note": "bytes = 4 Nq (codes) + 16 Nq (two offsets per lookup) + 4 P (postings) + 4 Q R (matrix written once), SURVEY.md s8d"}
            # end to end: host query buffers in, statistics rows on the host out (sparse job)
            qc_h, qi_h = tq.cpu().numpy().view(np.uint32), ti.cpu().numpy().view(np.uint64)
            from public_kssd_b200 import capi as _capi
            row_b = _capi.STAT_ROW_DTYPE.itemsize
            pin_q = torch.from_numpy(qc_h.view(np.int32).copy()).pin_memory()              # the caller's buffers: pinned, allocated once
            pin_rows = torch.empty(len(want) * row_b + 4096, dtype=torch.uint8, pin_memory=True)
            qc_p = pin_q.numpy().view(np.uint32)
            e2e, n_rows = [], 0
            for _ in range(3):
                t1 = time.perf_counter()
                sj = kssd.DistJob(ctx, qsz0, ref_sizes, sparse=True)
                sj.accumulate(full, qc_p, qi_h)
                n_rows = int(sj.stats(skip_zero=1, fetch=False, cmprsn_num=cm))
                assert n_rows * row_b <= pin_rows.numel()
                sj.fetch_stats_into(pin_rows.data_ptr())
                sj.close()
                e2e.append(time.perf_counter() - t1)
            rows_h = pin_rows.numpy()[: n_rows * row_b].view(_capi.STAT_ROW_DTYPE)
            same_rows = bool(n_rows == len(want) and rows_h.tobytes() == want.tobytes())
            out["e2e"] = {"value": n_qry * n_ref / min(e2e), "unit": "pairs/s", "ms": min(e2e) * 1e3, "h2d_bytes": int(qc_h.nbytes + qi_h.nbytes + qsz0.nbytes),
                          "d2h_bytes": int(n_rows * row_b), "rows_equal_dense_job": same_rows,
                          "note": "kssd_dist_create_sparse + add_host + stats + fetch_stats: query sketches from pinned host memory, all statistics rows back into pinned host memory"}
        dj.close()
        full.close()
    if world > 1:
        ok = torch.tensor([1 if out.get("content_ok", True) else 0], device=dev, dtype=torch.int64)
        dist.broadcast(ok, 0)

    # ---- Stage I -> Stage II exchange at this size (SURVEY.md s8e "Index: one exchange step"): every rank starts with the sketches of
    # ITS genome block, as Stage I leaves them; one all-to-all of (code, gid) pairs routes every code to the rank that owns its
    # code range, and the rank indexes what it received.  Not on the main line above (which replicates the index), but what a
    # reference set beyond one GPU's memory would do.
    if world > 1:
        try:
            g = parallel.genome_shard(n_ref, world, rank)
            a, bnd = int(ref_index_host[g.start]), int(ref_index_host[g.stop])
            loc_c = t_rc[a:bnd].contiguous()
            loc_i = (t_ri[g.start:g.stop + 1] - t_ri[g.start]).contiguous()
            ex_ms, ix_ms2, got_n = [], [], 0
            for it in range(3):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                xc, xi = parallel.exchange_codes_by_range(loc_c, loc_i, g.start, n_ref, world, rank, code_bits=28, device=dev)
                e1.record()
                torch.cuda.synchronize()
                xix = ctx.combco2mco_dev(xc.data_ptr(), xi.data_ptr(), n_ref, int(xc.numel()))
                ix_ms2.append(ctx.last_ms(2))
                xix.close()
                barrier()
                ex_ms.append(e0.elapsed_time(e1))
                got_n = int(xc.numel())
            lo_c, hi_c = parallel.code_range(rank, world, 28)
            want_n = int(((t_rc >= lo_c) & (t_rc < hi_c)).sum().item())                      # codes of ALL genomes in this rank's range
            # (sum of code * (gid + 1)) mod 2^63 of what arrived == the same sum over the whole reference filtered to the range
            gid_all = torch.repeat_interleave(torch.arange(n_ref, device=dev, dtype=torch.int64), (t_ri[1:] - t_ri[:-1]))
            keep = (t_rc >= lo_c) & (t_rc < hi_c)
            want_sum = int((t_rc[keep].to(torch.int64) * (gid_all[keep] + 1)).sum().item())
            gid_got = torch.repeat_interleave(torch.arange(n_ref, device=dev, dtype=torch.int64), (xi[1:] - xi[:-1]))
            got_sum = int((xc.to(torch.int64) * (gid_got + 1)).sum().item())
            tt = torch.tensor([min(ex_ms[1:]), min(ix_ms2[1:]), 1.0 if (got_n == want_n and got_sum == want_sum) else 0.0], device=dev, dtype=torch.float64)
            dist.all_reduce(tt[:2], op=dist.ReduceOp.MAX)
            dist.all_reduce(tt[2:], op=dist.ReduceOp.MIN)
            out["exchange"] = {"ms": float(tt[0].item()), "index_ms_per_rank_after": float(tt[1].item()), "postings": n_codes,
                               "bytes_sent_per_rank": int(8 * (bnd - a) * (world - 1) // world), "content_ok": bool(tt[2].item() == 1.0),
                               "note": "exchange_codes_by_range: sketches sharded by genome -> (code, gid) pairs routed by code range with one NCCL all-to-all (8 B per "
                                       "posting), max over ranks; content = count and a (code, gid) checksum of what each rank received against the whole reference "
                                       "filtered to its range"}
        except Exception as ex:
            out["exchange"] = {"failed": str(ex)}

    # ---- the baselines at N > 1, one batch: index by CODE range + NCCL reduce-scatter of dense partial matrices (north
    # star), and the same placement with the count kernel adding into the owner's rows over peer memory
    if world > 1 and baselines:
        rc_h = t_rc.cpu().numpy().view(np.uint32)
        tq, ti = q_dev[0]
        for mode in ("code", "code_p2p"):
            try:
                sd = parallel.ShardedDist(ctx, world, rank, code_bits=28, mode=mode).build_reference(rc_h, ref_index_host)
                ts = []
                for it in range(3):
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    torch.cuda.current_stream().synchronize()
                    sd.search(tq if rank == 0 else None, ti if rank == 0 else None, src=0, stats_opts={"skip_zero": 1}, fetch_counts=False, fetch_stats=False)
                    e1.record(stream)
                    barrier()
                    if it:
                        ts.append(e0.elapsed_time(e1))
                tt = torch.tensor([min(ts)], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                out[f"baseline_{mode}_ms_per_batch"] = float(tt.item())
                sd.close()
            except Exception as ex:
                out[f"baseline_{mode}_ms_per_batch"] = f"failed: {ex}"
        out["baseline_note"] = ("code = reference index sharded by code range, queries broadcast, dense partial Q x R matrices combined by ncclReduceScatter "
                                "(north star); code_p2p = same placement, RED.ADD into the owner's rows over NVLink peer mappings; both one batch, "
                                "library-stream events incl. broadcast and barriers")
    return out
