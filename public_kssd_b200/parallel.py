"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL over NVLink on the box, gloo in CPU tests).

SURVEY.md s8(e) / BASELINE.json north_star:
  * Stage I shards by input GENOME: every rank sketches its own genomes, there is no data-path collective;
    the (tiny) per-genome code lists are gathered to whoever writes combco.
  * Stage III shards the REFERENCE INDEX BY CODE RANGE: rank r indexes only the reference codes that fall in
    its slice of the code space, the query sketches are broadcast, every rank counts into a full Q x R matrix of
    partial counts, and one reduce-scatter (uint32 sum) leaves rank r with the final counts of its block of
    query rows; statistics run on the owner.

The planning functions are pure Python (tested with gloo on CPU); the compute they drive is the CUDA library.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------
# plans (pure functions)
# ------------------------------------------------------------------------------------------------
def genome_shard(n_genomes: int, world: int, rank: int) -> range:
    """Contiguous block of genomes for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_genomes, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def balanced_genome_shards(sizes, world: int):
    """Size-balanced assignment (longest-processing-time first): list of genome-id lists, one per rank."""
    order = np.argsort(-np.asarray(sizes, dtype=np.int64), kind="stable")
    load = [0] * world
    out = [[] for _ in range(world)]
    for g in order:
        r = int(np.argmin(load))
        out[r].append(int(g))
        load[r] += int(sizes[g])
    return [sorted(x) for x in out]


def code_range(rank: int, world: int, code_bits: int = 28) -> tuple[int, int]:
    """[lo, hi) slice of the code space owned by `rank` (equal-width ranges; codes are hash-like, so equal width
    is equal load)."""
    space = 1 << code_bits
    return (space * rank) // world, (space * (rank + 1)) // world


def row_block(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Query rows whose final counts land on `rank` after the reduce-scatter (equal blocks, padded at the end)."""
    per = (n_rows + world - 1) // world
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def filter_codes_to_range(codes: np.ndarray, index: np.ndarray, lo: int, hi: int):
    """Restrict a combco (codes, index) to codes in [lo, hi): the shard of the reference a rank indexes."""
    codes = np.asarray(codes, dtype=np.uint32)
    index = np.asarray(index, dtype=np.uint64)
    keep = (codes >= lo) & (codes < hi)
    csum = np.concatenate([[0], np.cumsum(keep, dtype=np.uint64)])
    return codes[keep], csum[index.astype(np.int64)]


# ------------------------------------------------------------------------------------------------
# collectives
# ------------------------------------------------------------------------------------------------
def reduce_scatter_rows(partial, world: int, rank: int):
    """partial: torch int32 tensor [rows_padded, R] with rows_padded % world == 0.  Returns this rank's
    [rows_padded / world, R] block of the element-wise sum over ranks."""
    import torch
    import torch.distributed as dist
    rows = partial.shape[0] // world
    if world == 1:
        return partial
    out = torch.empty((rows, partial.shape[1]), dtype=partial.dtype, device=partial.device)
    if dist.get_backend() == "nccl":
        dist.reduce_scatter_tensor(out, partial, op=dist.ReduceOp.SUM)
    else:   # gloo (CPU tests): same result through all_reduce + slice
        tmp = partial.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        out.copy_(tmp[rank * rows:(rank + 1) * rows])
    return out


def exchange_codes_by_range(codes, index, gid_lo: int, n_genomes: int, world: int, rank: int, code_bits: int = 28, device=None):
    """The exchange step between Stage I (sharded by genome) and Stage II (sharded by code range), SURVEY.md s8(e):
    this rank holds the sketches (codes, index) of the contiguous genome block starting at `gid_lo`; after one
    all-to-all every rank holds, for ALL `n_genomes` genomes, the codes that fall in ITS slice of the code space --
    a combco (codes, index[n_genomes + 1]) ready for combco2mco.  Genome blocks ascend with the rank, the routing
    sort is stable, so the received codes arrive grouped by genome id without any re-sort.
    codes / index: numpy arrays or torch tensors (CUDA tensors stay on the device under NCCL).  Returns torch tensors
    (codes int32, index int64) on `device`."""
    import torch
    import torch.distributed as dist
    dev = torch.device(device) if device is not None else (codes.device if torch.is_tensor(codes) else torch.device("cpu"))
    c = (codes if torch.is_tensor(codes) else torch.from_numpy(np.ascontiguousarray(codes).view(np.int32))).to(dev).to(torch.int32)
    ix = (index if torch.is_tensor(index) else torch.from_numpy(np.ascontiguousarray(index).astype(np.int64))).to(dev).to(torch.int64)
    n_local = ix.numel() - 1
    counts = ix[1:] - ix[:-1]
    gids = torch.repeat_interleave(torch.arange(gid_lo, gid_lo + n_local, device=dev, dtype=torch.int32), counts)
    bounds = torch.tensor([code_range(r, world, code_bits)[0] for r in range(1, world)], device=dev, dtype=torch.int32)
    dest = torch.bucketize(c, bounds, right=True)                       # owner rank of every code
    # stable order by owner: a one-byte key is ONE radix pass (the int64 ranks bucketize returns would be eight)
    order = torch.sort(dest.to(torch.uint8) if world <= 256 else dest, stable=True).indices
    send_counts = torch.bincount(dest, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    if world > 1:
        dist.all_to_all_single(recv_counts, send_counts)
    else:
        recv_counts.copy_(send_counts)
    ss, rs = send_counts.tolist(), recv_counts.tolist()
    payload = torch.stack([c[order], gids[order]], dim=1).contiguous()   # (code, gid) pairs, 8 B per posting
    got = torch.empty((sum(rs), 2), dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_to_all_single(got, payload, output_split_sizes=rs, input_split_sizes=ss)
    else:
        got.copy_(payload)
    new_index = torch.zeros(n_genomes + 1, dtype=torch.int64, device=dev)
    new_index[1:] = torch.cumsum(torch.bincount(got[:, 1].to(torch.int64), minlength=n_genomes), 0)
    return got[:, 0].contiguous(), new_index


def genome_block(codes: np.ndarray, index: np.ndarray, lo: int, hi: int):
    """combco (codes, index) of the genomes [lo, hi) only, index rebased to start at 0."""
    index = np.asarray(index, dtype=np.uint64)
    a, b = int(index[lo]), int(index[hi])
    return np.asarray(codes, dtype=np.uint32)[a:b], index[lo:hi + 1] - index[lo]


class ShardedDist:
    """Stage III over `world` GPUs.  Every rank calls the same methods (SPMD).

    mode "code"   (north star): reference index sharded by code range, queries broadcast, every rank counts a full
                  Q x R matrix of partial counts, one reduce-scatter leaves each rank the final counts of its block
                  of query rows.
    mode "code_p2p" same placement as "code", but the count kernel adds straight into the owner's rows through peer
                  mappings (CUDA IPC over NVLink): compute and reduction are one kernel, only the non-zero increments
                  cross the links, no partial matrix and no NCCL collective on the data path.
    mode "genome" (zero-communication alternative): rank r indexes the reference genomes of its
                  block; its Q x R/world count columns are final as they are -- no reduction at all.  When the counts are
                  not fetched the rank runs a SPARSE job (no Q x R/world matrix; kssd_dist_create_sparse) -- it falls
                  back to the matrix by itself if the options print zero-shared cells."""

    def __init__(self, ctx, world: int, rank: int, code_bits: int = 28, mode: str = "code"):
        assert mode in ("code", "code_p2p", "genome")
        self.ctx, self.world, self.rank, self.code_bits, self.mode = ctx, world, rank, code_bits, mode
        self.index = None
        self.ref_sizes = None
        self.col_lo = self.col_hi = 0
        self._partial = None
        self._blocks = None

    def build_reference(self, ref_codes: np.ndarray, ref_index: np.ndarray):
        """Every rank sees the reference combco (or at least its own shard of it) and indexes its slice."""
        self.ref_sizes = np.diff(np.asarray(ref_index, dtype=np.uint64)).astype(np.uint32)
        if self.mode in ("code", "code_p2p"):
            lo, hi = code_range(self.rank, self.world, self.code_bits)
            c, ix = filter_codes_to_range(ref_codes, ref_index, lo, hi)
        else:
            g = genome_shard(len(ref_index) - 1, self.world, self.rank)
            self.col_lo, self.col_hi = g.start, g.stop
            c, ix = genome_block(ref_codes, ref_index, g.start, g.stop) if g.stop > g.start else (np.zeros(0, np.uint32), np.zeros(2, np.uint64))
        self.index = self.ctx.combco2mco(c, ix) if len(ix) > 1 and (self.mode != "genome" or self.col_hi > self.col_lo) else None
        return self

    def build_reference_exchanged(self, local_codes, local_index, gid_lo: int, n_genomes: int, ref_sizes):
        """Stage I -> II across ranks without ever gathering the reference: this rank brings the sketches of ITS genome
        block (as Stage I left them), `exchange_codes_by_range` routes every code to the rank that owns its code range,
        and the rank indexes what it received.  ref_sizes: sketch sizes of all genomes (a few bytes each, all-gathered
        by the caller)."""
        assert self.mode in ("code", "code_p2p")
        import torch
        self.ref_sizes = np.asarray(ref_sizes, dtype=np.uint32)
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        c, ix = exchange_codes_by_range(local_codes, local_index, gid_lo, n_genomes, self.world, self.rank, self.code_bits, device=dev)
        if c.is_cuda:
            self.index = self.ctx.combco2mco_dev(c.data_ptr(), ix.data_ptr(), n_genomes, int(c.numel()))
            torch.cuda.synchronize()
        else:
            self.index = self.ctx.combco2mco(c.numpy().view(np.uint32), ix.numpy().astype(np.uint64))
        return self

    # ---- peer-memory row blocks (mode "code_p2p") ----
    def _peer_blocks(self, per: int, R: int):
        """Allocate this rank's uint32[per][R] block (plain cudaMalloc, IPC-exportable), exchange the IPC handles and open
        the peers' blocks.  Cached per shape."""
        import ctypes as C
        import torch.distributed as dist
        from .capi import check, lib
        if self._blocks is not None and self._blocks[0] == (per, R):
            return self._blocks[1], self._blocks[2]
        self._release_blocks()
        own = C.c_void_p()
        check(lib().kssd_dev_alloc(self.ctx._h, per * R * 4, C.byref(own)))
        handle = (C.c_uint8 * 64)()
        check(lib().kssd_ipc_export(self.ctx._h, own, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle))
        ptrs = (C.c_void_p * self.world)()
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs[r] = own.value
            else:
                p = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(h)
                check(lib().kssd_ipc_open(self.ctx._h, hb, C.byref(p)))
                ptrs[r] = p.value
        self._blocks = ((per, R), own, ptrs)
        return own, ptrs

    def _release_blocks(self):
        import torch.distributed as dist
        from .capi import lib
        if self._blocks is None:
            return
        _, own, ptrs = self._blocks
        for r in range(self.world):
            if r != self.rank and ptrs[r]:
                lib().kssd_ipc_close(self.ctx._h, ptrs[r])
        if dist.is_initialized():
            dist.barrier()                     # nobody frees a block a peer still maps
        lib().kssd_dev_free(self.ctx._h, own)
        self._blocks = None

    def close(self):
        self._release_blocks()

    def search(self, qry_codes=None, qry_index=None, src: int = 0, stats_opts: dict | None = None, fetch_counts: bool = True,
               fetch_stats: bool = True):
        """Broadcast the query sketches from `src`, count, (code mode) reduce-scatter, statistics.  qry_codes/qry_index:
        numpy arrays or CUDA tensors (int32 view of the uint32 codes, int64 index) on `src`, None elsewhere.
        Returns (lo, hi, counts block as numpy uint32 or None, stats rows / row count / None): in code mode the block is
        the query ROWS [lo, hi) x all refs, in genome mode all queries x the reference COLUMNS [lo, hi)."""
        import torch
        import torch.distributed as dist
        from . import kssd
        dev = torch.device("cuda", self.ctx.device)

        def to_dev(a, np_dtype, view, tdtype):
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                return a.to(dev)
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype).view(view)).to(dev)

        tq, ti = to_dev(qry_codes, np.uint32, np.int32, torch.int32), to_dev(qry_index, np.uint64, np.int64, torch.int64)
        if self.world > 1:
            meta = torch.zeros(2, dtype=torch.int64, device=dev)
            if self.rank == src:
                meta[0], meta[1] = ti.numel() - 1, tq.numel()
            dist.broadcast(meta, src)
            nq, nc = int(meta[0]), int(meta[1])
            if self.rank != src:
                tq = torch.empty(max(nc, 1), dtype=torch.int32, device=dev)
                ti = torch.empty(nq + 1, dtype=torch.int64, device=dev)
            dist.broadcast(tq, src)
            dist.broadcast(ti, src)
        else:
            nq, nc = ti.numel() - 1, tq.numel()
        qsizes = (ti[1:] - ti[:-1]).cpu().numpy().astype(np.uint32)
        R = int(self.ref_sizes.size)
        if self.mode == "genome":
            # this rank's reference columns are final: count, then statistics, no collective
            lo, hi = self.col_lo, self.col_hi
            if hi <= lo:
                return lo, hi, np.zeros((nq, 0), dtype=np.uint32), None
            job = kssd.DistJob(self.ctx, qsizes, self.ref_sizes[lo:hi], sparse=(stats_opts is not None and not fetch_counts))
            torch.cuda.synchronize()
            job.accumulate_dev(self.index, tq.data_ptr(), ti.data_ptr(), nc)
            rows = None
            if stats_opts is not None:
                rows = job.stats(cmprsn_num=(R * nq) & 0xFFFFFFFF, fetch=fetch_stats, **stats_opts)
                if fetch_stats:
                    rows["ref"] += lo
            block = job.counts() if fetch_counts else None
            job.close()
            return lo, hi, block, rows
        per = (nq + self.world - 1) // self.world
        if self.mode == "code_p2p":
            import ctypes as C
            from .capi import check, lib
            own, ptrs = self._peer_blocks(per, R)
            check(lib().kssd_dev_zero(self.ctx._h, own, per * R * 4))
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()                 # every owner has zeroed its rows
            check(lib().kssd_dist_accumulate_peer(self.ctx._h, self.index._h, C.c_void_p(tq.data_ptr()), C.c_void_p(ti.data_ptr()), nq, R,
                                                  ptrs, per, self.world))
            if self.world > 1:
                dist.barrier()                 # every rank's increments have landed
            lo, hi = row_block(nq, self.world, self.rank)
            rows = None
            if stats_opts is not None and hi > lo:
                sj = kssd.DistJob(self.ctx, qsizes[lo:hi], self.ref_sizes, ct_dev_ptr=own.value, already_filled=True)
                rows = sj.stats(cmprsn_num=(R * nq) & 0xFFFFFFFF, fetch=fetch_stats, **stats_opts)
                if fetch_stats:
                    rows["qry"] += lo
                block = sj.counts() if fetch_counts else None
                sj.close()
            else:
                block = None
                if fetch_counts and hi > lo:
                    sj = kssd.DistJob(self.ctx, qsizes[lo:hi], self.ref_sizes, ct_dev_ptr=own.value, already_filled=True)
                    block = sj.counts()
                    sj.close()
            return lo, hi, block, rows
        rows_padded = per * self.world
        if self._partial is None or tuple(self._partial.shape) != (rows_padded, R):
            self._partial = torch.zeros((rows_padded, R), dtype=torch.int32, device=dev)     # padding rows stay zero
        partial = self._partial
        job = kssd.DistJob(self.ctx, qsizes, self.ref_sizes, ct_dev_ptr=partial.data_ptr())
        torch.cuda.synchronize()
        job.accumulate_dev(self.index, tq.data_ptr(), ti.data_ptr(), nc)      # the kernel zeroes and fills rows [0, nq)
        job.close()
        mine = reduce_scatter_rows(partial, self.world, self.rank)
        torch.cuda.synchronize()
        lo, hi = row_block(nq, self.world, self.rank)
        rows = None
        if stats_opts is not None and hi > lo:
            sj = kssd.DistJob(self.ctx, qsizes[lo:hi], self.ref_sizes, ct_dev_ptr=mine.data_ptr(), already_filled=True)
            rows = sj.stats(cmprsn_num=(R * nq) & 0xFFFFFFFF, fetch=fetch_stats, **stats_opts)
            if fetch_stats:
                rows["qry"] += lo
            sj.close()
        return lo, hi, (mine[: hi - lo].cpu().numpy().view(np.uint32) if fetch_counts else None), rows
