"""Host-side mirror of the reference's stage functions on top of the C-ABI (capi.py).

Names follow the reference seams they stand for:

    Context            read_dim_shuffle_file + get_hashsz + seq2co_global_var_initial
    Context.sketch     fasta2co / uniq_fasta2co (+ the wrt_* writers) for a batch of genomes
    Context.combco2mco co2mco.c:25 combco2mco for one component
    DistJob            mco_cbdco_nobin_dist hot loop + output_ctrl numbers

All compute happens in libkssd_b200.so on the GPU; this module only moves buffers and keeps the
reference's data shapes (combco / combco.index / mco / sharedk_ct matrices).
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass
from pathlib import Path
from typing import Sequence

import numpy as np

from . import capi
from .capi import KssdError, check, lib, ptr


def read_shuf_file(path) -> tuple[dict, np.ndarray]:
    """.shuf: 16-byte dim_shuffle_stat_t {id,k,subk,drlevel} + int32[16^subk] (command_shuffle.c:192-207)."""
    path = str(path)
    if not path.endswith(".shuf"):
        raise ValueError(f"read_dim_shuffle_file(): input file {path} is not .shuf file")
    with open(path, "rb") as f:
        sid, k, subk, drlevel = struct.unpack("<iiii", f.read(16))
        table = np.fromfile(f, dtype="<i4", count=1 << (4 * subk))
    return dict(id=sid, k=k, subk=subk, drlevel=drlevel), table


def write_shuf_file(path, shuf_id: int, k: int, subk: int, drlevel: int, table: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", shuf_id, k, subk, drlevel))
        f.write(np.ascontiguousarray(table, dtype="<i4").tobytes())


def pack_genomes(genomes: Sequence[bytes | np.ndarray], align: int = 128, pinned: bool = False):
    """Lay genomes out in one byte buffer, each start aligned (>= 16 required by the C-ABI).
    Returns (buffer uint8, goff uint64[n], glen uint64[n])."""
    n = len(genomes)
    glen = np.array([len(g) for g in genomes], dtype=np.uint64)
    goff = np.zeros(n, dtype=np.uint64)
    o = 0
    for i in range(n):
        goff[i] = o
        o += (int(glen[i]) + align - 1) // align * align
    total = max(o, align)
    if pinned:
        import torch
        buf = torch.empty(total, dtype=torch.uint8, pin_memory=True).numpy()
    else:
        buf = np.empty(total, dtype=np.uint8)
    buf[:] = 0x0A
    for i, g in enumerate(genomes):
        a = np.frombuffer(g, dtype=np.uint8) if not isinstance(g, np.ndarray) else g
        buf[int(goff[i]): int(goff[i]) + a.size] = a
    return buf, goff, glen


@dataclass
class Sketch:
    """One batch of sketches: per component the content of combco.<c> / combco.index.<c> (ids ascending per genome)."""
    ids: list          # per component uint32[]
    index: list        # per component uint64[n+1]
    abund: list        # per component uint16[] (abundance mode) or None
    ord: list          # per component uint64[] first-occurrence byte offsets or None
    status: np.ndarray  # per genome 0 / KSSD_E_*
    n_occurrences: int
    scan_ms: float
    total_ms: float

    def genome_sets(self):
        n = len(self.index[0]) - 1
        return [[self.ids[c][int(self.index[c][g]): int(self.index[c][g + 1])] for c in range(len(self.ids))] for g in range(n)]

    def ctx_ct(self) -> np.ndarray:
        """per-genome sketch size over all components (cofiles.stat ctx_ct list)."""
        return sum(np.diff(ix).astype(np.uint32) for ix in self.index)


class Context:
    """One GPU + one .shuf: the globals of the reference's Stage I as an object."""

    def __init__(self, k: int, subk: int, drlevel: int, table: np.ndarray, device: int = 0, component_sz: int = 7,
                 shuf_id: int = 0):
        self._h = C.c_void_p()
        table = np.ascontiguousarray(table, dtype=np.int32)
        if table.size != 1 << (4 * subk):
            raise ValueError("shuf table must hold 16^subk entries")
        check(lib().kssd_ctx_create(C.byref(self._h), device, ptr(table, C.c_int32), k, subk, drlevel, component_sz))
        self.info = capi.CtxInfo()
        check(lib().kssd_ctx_info(self._h, C.byref(self.info)))
        self.k, self.subk, self.drlevel, self.shuf_id = k, subk, drlevel, shuf_id
        self.component_num = self.info.component_num
        self.device = device

    @classmethod
    def from_shuf_file(cls, path, device: int = 0, component_sz: int = 7) -> "Context":
        st, table = read_shuf_file(path)
        return cls(st["k"], st["subk"], st["drlevel"], table, device=device, component_sz=component_sz, shuf_id=st["id"])

    def close(self):
        if self._h:
            lib().kssd_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(lib().kssd_ctx_stream(self._h) or 0)

    def last_ms(self, which: int) -> float:
        return float(lib().kssd_ctx_last_ms(self._h, which))

    # ---------------- Stage I ----------------
    def sketch_raw(self, buf, nbytes: int, goff: np.ndarray, glen: np.ndarray, mode: int = capi.MODE_FASTA, Q: int = 0, M: int = 1,
                   span_bytes: int = 0, device_ptr: int | None = None):
        """Returns an opaque sketch handle (int).  `buf` numpy uint8 (host) or device_ptr (int) for resident input."""
        goff = np.ascontiguousarray(goff, dtype=np.uint64)
        glen = np.ascontiguousarray(glen, dtype=np.uint64)
        opts = capi.SketchOpts(mode, Q, M, 1, span_bytes, 0)
        h = C.c_void_p()
        if device_ptr is not None:
            check(lib().kssd_sketch_batch_dev(self._h, C.c_void_p(device_ptr), nbytes, ptr(goff, C.c_uint64), ptr(glen, C.c_uint64),
                                              len(goff), C.byref(opts), C.byref(h)))
        else:
            check(lib().kssd_sketch_batch_host(self._h, buf.ctypes.data_as(C.c_void_p), nbytes, ptr(goff, C.c_uint64),
                                               ptr(glen, C.c_uint64), len(goff), C.byref(opts), C.byref(h)))
        return h

    def fetch_sketch(self, h, n_genomes: int, free: bool = True, want_ord: bool = True, want_abund: bool = False) -> Sketch:
        ids, index, abund, ordl = [], [], [], []
        for c in range(self.component_num):
            n = check(lib().kssd_sketch_count(h, c))
            a = np.empty(n, dtype=np.uint32)
            ix = np.empty(n_genomes + 1, dtype=np.uint64)
            ab = np.empty(n, dtype=np.uint16) if want_abund else None
            od = np.empty(n, dtype=np.uint64) if want_ord else None
            check(lib().kssd_sketch_fetch(h, c, ptr(a, C.c_uint32), ptr(ix, C.c_uint64), ptr(ab, C.c_uint16) if want_abund else None,
                                          ptr(od, C.c_uint64) if want_ord else None))
            ids.append(a); index.append(ix); abund.append(ab); ordl.append(od)
        status = np.zeros(n_genomes, dtype=np.int32)
        check(lib().kssd_sketch_status(h, ptr(status, C.c_int32)))
        nocc = C.c_uint64(0)
        ms = C.c_float(0)
        check(lib().kssd_sketch_stats(h, C.byref(nocc), C.byref(ms)))
        total_ms = self.last_ms(1)
        if free:
            lib().kssd_sketch_free(h)
        return Sketch(ids, index, abund, ordl, status, int(nocc.value), float(ms.value), total_ms)

    def _raise_status(self, sk: Sketch):
        for g, s in enumerate(sk.status):
            if s == capi.E_CROWD:
                raise KssdError(s, f"the context space is too crowd, try rerun the program using -k{self.k + 1} (genome {g})")
            if s == capi.E_HEADER_EOF:
                raise KssdError(s, f"fasta2co(): can not find seqences head start from '>' (genome {g})")
            if s == capi.E_LONGLINE:
                raise KssdError(s, f"FASTQ line longer than the reference's fgets buffer (genome {g})")

    def sketch(self, genomes: Sequence[bytes | np.ndarray], uniq: bool = False, span_bytes: int = 0, strict: bool = True) -> Sketch:
        """fasta2co / uniq_fasta2co + writer for each genome of the batch (host buffers in, host arrays out).
        strict: raise where the reference would have exited (crowded context space, header at EOF)."""
        buf, goff, glen = pack_genomes(genomes)
        h = self.sketch_raw(buf, buf.size, goff, glen, capi.MODE_FASTA_UNIQ if uniq else capi.MODE_FASTA, span_bytes=span_bytes)
        sk = self.fetch_sketch(h, len(genomes))
        if strict:
            self._raise_status(sk)
        return sk

    def sketch_fastq(self, files: Sequence[bytes | np.ndarray], Q: int = 0, M: int = 1, abundance: bool = False, strict: bool = True) -> Sketch:
        """fastq2co(Q, M) + write_fqco2file, or with abundance=True mt_shortreads2koc + write_fqkoc2files (-A)."""
        buf, goff, glen = pack_genomes(files)
        h = self.sketch_raw(buf, buf.size, goff, glen, capi.MODE_FASTQ_ABUND if abundance else capi.MODE_FASTQ, Q=Q, M=M)
        sk = self.fetch_sketch(h, len(files), want_abund=abundance)
        if strict:
            self._raise_status(sk)
        return sk

    def sketch_files(self, paths: Sequence[str], mode: int = capi.MODE_FASTA, Q: int = 0, M: int = 1, threads: int = 0,
                     batch_bytes: int = 0, strict: bool = True, pipecmd: str | None = None):
        """Stage I from files (run_stageI's file loop + the zcat decode): plain or .gz FASTA / FASTQ files, read and
        inflated by host threads into pinned staging buffers and sketched batch by batch (kssd_stage1_files).
        pipecmd: the reference's -P <cmd> -- every file is read from the stdout of "<cmd> <file>".
        Returns (Sketch over all files in input order, timing dict)."""
        arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
        opts = capi.SketchOpts(mode, Q, M, 0, 0, 0)
        h = C.c_void_p()
        check(lib().kssd_stage1_files_ex(self._h, arr, len(paths), C.byref(opts), threads, batch_bytes,
                                         pipecmd.encode() if pipecmd else None, C.byref(h)))
        try:
            ids, index, abund = [], [], []
            for c in range(self.component_num):
                n = check(lib().kssd_stage1_count(h, c))
                a = np.empty(n, dtype=np.uint32)
                ix = np.empty(len(paths) + 1, dtype=np.uint64)
                ab = np.empty(n, dtype=np.uint16) if mode == capi.MODE_FASTQ_ABUND else None
                check(lib().kssd_stage1_fetch(h, c, ptr(a, C.c_uint32), ptr(ix, C.c_uint64), ptr(ab, C.c_uint16) if ab is not None else None))
                ids.append(a); index.append(ix); abund.append(ab)
            status = np.zeros(len(paths), dtype=np.int32)
            check(lib().kssd_stage1_status(h, ptr(status, C.c_int32)))
            rs, gs, ts, nb, nbat = C.c_double(), C.c_double(), C.c_double(), C.c_uint64(), C.c_int()
            check(lib().kssd_stage1_timing(h, C.byref(rs), C.byref(gs), C.byref(ts), C.byref(nb), C.byref(nbat)))
            gz_gpu, gz_s = C.c_int(), C.c_double()
            check(lib().kssd_stage1_gz_info(h, C.byref(gz_gpu), C.byref(gz_s)))
        finally:
            lib().kssd_stage1_free(h)
        sk = Sketch(ids, index, abund, [None] * self.component_num, status, 0, 0.0, 0.0)
        if strict:
            self._raise_status(sk)
        return sk, dict(read_s=rs.value, gpu_s=gs.value, total_s=ts.value, bytes=int(nb.value), batches=int(nbat.value),
                        gz_on_gpu=bool(gz_gpu.value), gz_gpu_s=gz_s.value)

    def reads2mco(self, files: Sequence[bytes | np.ndarray], strict: bool = True):
        """`kssd dist --byread` (reference reads2mco, iseq2comem.c:78-186) for every file of FASTA-formatted reads in the
        batch.  Returns one dict per file: {"n_reads": readn, "ids": [per component: ids in stream order, duplicates
        kept], "index": [per component: the reference's combco.index.<c>, n_reads + 1 inclusive cumulative counts],
        "read_of": [per component: record number of every id]}."""
        buf, goff, glen = pack_genomes(files)
        h = self.sketch_raw(buf, buf.size, goff, glen, capi.MODE_BYREAD)
        try:
            n = len(files)
            nr = np.zeros(n, dtype=np.uint64)
            check(lib().kssd_sketch_read_counts(h, ptr(nr, C.c_uint64)))
            sk = self.fetch_sketch(h, n, free=False)
            if strict:
                self._raise_status(sk)
            out = []
            for g in range(n):
                rec = {"n_reads": int(nr[g]), "ids": [], "index": [], "read_of": []}
                for c in range(self.component_num):
                    a, b = int(sk.index[c][g]), int(sk.index[c][g + 1])
                    ix = np.empty(int(nr[g]) + 1, dtype=np.uint64)
                    check(lib().kssd_sketch_fetch_read_index(h, c, g, ptr(ix, C.c_uint64)))
                    rec["ids"].append(sk.ids[c][a:b])
                    rec["read_of"].append(sk.ord[c][a:b])
                    rec["index"].append(ix)
                out.append(rec)
            return out
        finally:
            lib().kssd_sketch_free(h)

    # ---------------- kssd set ----------------
    def set_union(self, combco: np.ndarray, uniq: bool = False) -> np.ndarray:
        """`kssd set -u` / `-q` for one component (sketch_union / uniq_sketch_union): pan.<c> / uniq_pan.<c>."""
        a = np.ascontiguousarray(combco, dtype=np.uint32)
        out = np.empty(max(a.size, 1), dtype=np.uint32)
        n = C.c_uint64(0)
        check(lib().kssd_set_union_host(self._h, ptr(a, C.c_uint32), a.size, int(uniq), ptr(out, C.c_uint32), C.byref(n)))
        return out[:n.value].copy()

    def set_operate(self, combco: np.ndarray, index: np.ndarray, pan: np.ndarray, intersect: bool):
        """`kssd set -i <pan>` (intersect=True) / `-s <pan>` for one component (sketch_operate): the filtered
        (combco.<c>, combco.index.<c>)."""
        a = np.ascontiguousarray(combco, dtype=np.uint32)
        ix = np.ascontiguousarray(index, dtype=np.uint64)
        pn = np.ascontiguousarray(pan, dtype=np.uint32)
        out = np.empty(max(a.size, 1), dtype=np.uint32)
        oix = np.empty(ix.size, dtype=np.uint64)
        check(lib().kssd_set_operate_host(self._h, ptr(a, C.c_uint32), ptr(ix, C.c_uint64), ix.size - 1, ptr(pn, C.c_uint32), pn.size,
                                          int(intersect), ptr(out, C.c_uint32), ptr(oix, C.c_uint64)))
        return out[:int(oix[-1])].copy(), oix

    def set_group(self, combco: np.ndarray, index: np.ndarray, groups: Sequence[Sequence[int]]):
        """`kssd set -g <grouping file>` for one component (grouping_genomes): per group of genome ids the DISTINCT codes of its
        members in the order of their first occurrence (members walked in the given order).  Returns (codes, index[n_groups + 1]);
        hostfmt.group_slot_order turns a group's codes into the order the reference writes them."""
        a = np.ascontiguousarray(combco, dtype=np.uint32)
        ix = np.ascontiguousarray(index, dtype=np.uint64)
        members = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.uint32) for g in groups]) if len(groups) else np.zeros(0, np.uint32))
        gix = np.zeros(len(groups) + 1, dtype=np.uint64)
        gix[1:] = np.cumsum([len(g) for g in groups])
        total = int(sum(int(ix[int(g) + 1] - ix[int(g)]) for g in members)) if members.size and members.max(initial=0) < ix.size - 1 else a.size
        out = np.empty(max(total, 1), dtype=np.uint32)
        oix = np.empty(len(groups) + 1, dtype=np.uint64)
        check(lib().kssd_set_group_host(self._h, ptr(a, C.c_uint32), ptr(ix, C.c_uint64), ix.size - 1, ptr(members, C.c_uint32), ptr(gix, C.c_uint64),
                                        len(groups), ptr(out, C.c_uint32), ptr(oix, C.c_uint64)))
        return out[:int(oix[-1])].copy(), oix

    # ---------------- kssd composite ----------------
    def composite(self, ref_indexes, qry_codes, qry_index, qry_abund, min_kmers: int = 0) -> np.ndarray:
        """`kssd composite -r <refs> -q <-A queries>` (get_species_abundance): per-component lists of the reference
        inverted indexes (combco2mco of the reference sketches) and of the query combco.<c> / combco.index.<c> /
        combco.<c>.a.  Returns the rows in the reference's print order (capi.COMP_ROW_DTYPE)."""
        nc = len(ref_indexes)
        qc = [np.ascontiguousarray(a, dtype=np.uint32) for a in qry_codes]
        qi = [np.ascontiguousarray(a, dtype=np.uint64) for a in qry_index]
        qa = [np.ascontiguousarray(a, dtype=np.uint16) for a in qry_abund]
        VP = C.c_void_p * nc
        rows, n = C.c_void_p(), C.c_uint64(0)
        check(lib().kssd_composite_host(self._h, nc, VP(*[ix._h for ix in ref_indexes]), VP(*[a.ctypes.data for a in qc]),
                                        VP(*[a.ctypes.data for a in qi]), VP(*[a.ctypes.data for a in qa]), len(qi[0]) - 1, min_kmers,
                                        C.byref(rows), C.byref(n)))
        try:
            if not n.value:
                return np.zeros(0, dtype=capi.COMP_ROW_DTYPE)
            return np.frombuffer(C.string_at(rows, n.value * capi.COMP_ROW_DTYPE.itemsize), dtype=capi.COMP_ROW_DTYPE).copy()
        finally:
            if rows:
                lib().kssd_host_free(rows)

    # ---------------- Stage II ----------------
    def combco2mco(self, combco: np.ndarray, cbdcoindex: np.ndarray) -> "Index":
        combco = np.ascontiguousarray(combco, dtype=np.uint32)
        cbdcoindex = np.ascontiguousarray(cbdcoindex, dtype=np.uint64)
        h = C.c_void_p()
        check(lib().kssd_index_build_host(self._h, ptr(combco, C.c_uint32), ptr(cbdcoindex, C.c_uint64), len(cbdcoindex) - 1, C.byref(h)))
        return Index(self, h)

    def combco2mco_dev(self, codes_dev_ptr: int, index_dev_ptr: int, n_genomes: int, n_codes: int) -> "Index":
        """combco2mco on device-resident input: uint32 codes[n_codes], uint64 index[n_genomes + 1] (kssd_index_build_dev)."""
        h = C.c_void_p()
        check(lib().kssd_index_build_dev(self._h, C.c_void_p(codes_dev_ptr), C.c_void_p(index_dev_ptr), n_genomes, n_codes, C.byref(h)))
        return Index(self, h)

    def index_from_dense(self, dense_incl: np.ndarray, gids: np.ndarray, n_genomes: int) -> "Index":
        dense_incl = np.ascontiguousarray(dense_incl, dtype=np.uint64)
        gids = np.ascontiguousarray(gids, dtype=np.uint32)
        h = C.c_void_p()
        check(lib().kssd_index_from_dense_host(self._h, ptr(dense_incl, C.c_uint64), ptr(gids, C.c_uint32), gids.size, n_genomes, C.byref(h)))
        return Index(self, h)


class Index:
    """Inverted index of one component (mco.<c> + mco.index.<c>), resident on the GPU."""

    def __init__(self, ctx: Context, h):
        self.ctx, self._h = ctx, h
        nu, npst, ng = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        check(lib().kssd_index_sizes(h, C.byref(nu), C.byref(npst), C.byref(ng)))
        self.n_unique, self.n_postings, self.n_genomes = int(nu.value), int(npst.value), int(ng.value)

    def csr(self):
        """(unique codes, exclusive offsets, gids) -- gids is the content of mco.<c>."""
        uc = np.empty(self.n_unique, dtype=np.uint32)
        uo = np.empty(self.n_unique + 1, dtype=np.uint64)
        g = np.empty(self.n_postings, dtype=np.uint32)
        check(lib().kssd_index_fetch(self._h, ptr(uc, C.c_uint32), ptr(uo, C.c_uint64), ptr(g, C.c_uint32)))
        return uc, uo, g

    def dense(self) -> np.ndarray:
        """mco.index.<c>: uint64[16^COMPONENT_SZ] inclusive prefix (co2mco.c:57-61)."""
        d = np.empty(1 << (4 * self.ctx.info.component_sz), dtype=np.uint64)
        check(lib().kssd_index_fetch_dense(self._h, ptr(d, C.c_uint64)))
        return d

    def close(self):
        if self._h:
            lib().kssd_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistJob:
    """Q x R shared-k-mer count matrix (sharedk_ct.dat) and the statistics of distance.out."""

    def __init__(self, ctx: Context, qry_ctx_ct: np.ndarray, ref_ctx_ct: np.ndarray, ct_dev_ptr: int | None = None,
                 already_filled: bool = False, sparse: bool = False):
        """ct_dev_ptr: optional caller-owned device buffer (Q*R uint32) for the count matrix, e.g. a torch tensor that
        takes part in a reduce-scatter (parallel.py).
        sparse: no count matrix -- accumulate() only registers the component and stats() counts, filters and lists in
        one kernel (kssd_dist_create_sparse); same rows, work proportional to the postings touched."""
        self.ctx = ctx
        self.sparse = sparse
        q = np.ascontiguousarray(qry_ctx_ct, dtype=np.uint32)
        r = np.ascontiguousarray(ref_ctx_ct, dtype=np.uint32)
        self.n_qry, self.n_ref = q.size, r.size
        self._h = C.c_void_p()
        if sparse:
            assert ct_dev_ptr is None
            check(lib().kssd_dist_create_sparse(ctx._h, q.size, r.size, ptr(q, C.c_uint32), ptr(r, C.c_uint32), C.byref(self._h)))
        elif ct_dev_ptr is None:
            check(lib().kssd_dist_create(ctx._h, q.size, r.size, ptr(q, C.c_uint32), ptr(r, C.c_uint32), C.byref(self._h)))
        else:
            check(lib().kssd_dist_create_ext(ctx._h, q.size, r.size, ptr(q, C.c_uint32), ptr(r, C.c_uint32), C.c_void_p(ct_dev_ptr),
                                             int(already_filled), C.byref(self._h)))

    def accumulate(self, ref_index: Index, qcodes: np.ndarray, qindex: np.ndarray):
        qcodes = np.ascontiguousarray(qcodes, dtype=np.uint32)
        qindex = np.ascontiguousarray(qindex, dtype=np.uint64)
        f = lib().kssd_dist_sparse_add_host if self.sparse else lib().kssd_dist_accumulate_host
        check(f(self._h, ref_index._h, ptr(qcodes, C.c_uint32), ptr(qindex, C.c_uint64)))

    def accumulate_dev(self, ref_index: Index, qcodes_ptr: int, qindex_ptr: int, n_qcodes: int):
        if self.sparse:      # the pointers must stay valid until stats() returns
            check(lib().kssd_dist_sparse_add_dev(self._h, ref_index._h, C.c_void_p(qcodes_ptr), C.c_void_p(qindex_ptr), n_qcodes))
            return
        check(lib().kssd_dist_accumulate_dev(self._h, ref_index._h, C.c_void_p(qcodes_ptr), C.c_void_p(qindex_ptr), n_qcodes))

    def counts(self) -> np.ndarray:
        ct = np.empty((self.n_qry, self.n_ref), dtype=np.uint32)
        check(lib().kssd_dist_fetch_counts(self._h, ptr(ct, C.c_uint32)))
        return ct

    def stats(self, metric: int = 0, correction: int = 0, kmerlen: int | None = None, dim_rd_len: int | None = None,
              dthreshold: float = 1.0, n_neighbors: int = 0, skip_zero: int = 0, fetch: bool = True, cmprsn_num: int = 0):
        o = capi.StatOpts(metric, correction, kmerlen if kmerlen is not None else 2 * self.ctx.k,
                          dim_rd_len if dim_rd_len is not None else 2 * self.ctx.drlevel, dthreshold, n_neighbors, skip_zero,
                          cmprsn_num)
        n = check(lib().kssd_dist_stats(self._h, C.byref(o)))
        if not fetch:
            return n
        rows = np.empty(n, dtype=capi.STAT_ROW_DTYPE)
        check(lib().kssd_dist_fetch_stats(self._h, rows.ctypes.data_as(C.c_void_p)))
        return rows

    def fetch_stats_into(self, host_ptr: int) -> None:
        """copy the rows of the last stats() / stats_wait() into caller-owned host memory (n_rows * 88 bytes; pinned memory makes it
        a single DMA at link speed)"""
        check(lib().kssd_dist_fetch_stats(self._h, C.c_void_p(host_ptr)))

    def distance_out(self, qry_names, ref_names, metric: int = 0, outfields: int = 2, header: bool = True) -> bytes:
        """distance.out of the last stats(fetch=False) / stats_wait(), formatted on the GPU (kssd_dist_format_text): header + one line
        per row, byte-identical to dist_print_nobin / output_ctrl (command_dist.c:1188-1195, 1267-1285).  Names: lists of str, or the
        256-byte name blocks of cofiles.stat / mcofiles.stat as bytes."""
        from . import hostfmt
        qn = qry_names if isinstance(qry_names, (bytes, bytearray)) else hostfmt._names_block(qry_names)
        rn = ref_names if isinstance(ref_names, (bytes, bytearray)) else hostfmt._names_block(ref_names)
        text, n = C.c_void_p(), C.c_size_t()
        check(lib().kssd_dist_format_text(self._h, bytes(qn), bytes(rn), 256, metric, outfields, int(header), C.byref(text), C.byref(n)))
        try:
            return C.string_at(text, n.value)
        finally:
            lib().kssd_host_free(text)

    def distance_out_view(self, qry_names, ref_names, metric: int = 0, outfields: int = 2, header: bool = True) -> memoryview:
        """distance_out() without a copy: a view of the context's pinned text buffer (kssd_dist_text), valid until the next call on
        this context -- write it to the file, or bytes() it."""
        from . import hostfmt
        qn = qry_names if isinstance(qry_names, (bytes, bytearray)) else hostfmt._names_block(qry_names)
        rn = ref_names if isinstance(ref_names, (bytes, bytearray)) else hostfmt._names_block(ref_names)
        text, n = C.c_void_p(), C.c_size_t()
        check(lib().kssd_dist_text(self._h, bytes(qn), bytes(rn), 256, metric, outfields, int(header), C.byref(text), C.byref(n)))
        if n.value == 0:
            return memoryview(b"")
        return memoryview((C.c_char * n.value).from_address(text.value)).cast("B")

    def stats_async(self, metric: int = 0, correction: int = 0, kmerlen: int | None = None, dim_rd_len: int | None = None,
                    dthreshold: float = 1.0, n_neighbors: int = 0, skip_zero: int = 0, cmprsn_num: int = 0):
        """Queue the whole sparse search (count + list + statistics) without waiting for the GPU; stats_wait() returns the row
        count.  Several jobs of one context may be in flight (kssd_dist_stats_async)."""
        o = capi.StatOpts(metric, correction, kmerlen if kmerlen is not None else 2 * self.ctx.k,
                          dim_rd_len if dim_rd_len is not None else 2 * self.ctx.drlevel, dthreshold, n_neighbors, skip_zero, cmprsn_num)
        check(lib().kssd_dist_stats_async(self._h, C.byref(o)))

    def stats_wait(self, fetch: bool = False):
        n = check(lib().kssd_dist_stats_wait(self._h))
        if not fetch:
            return n
        rows = np.empty(n, dtype=capi.STAT_ROW_DTYPE)
        check(lib().kssd_dist_fetch_stats(self._h, rows.ctypes.data_as(C.c_void_p)))
        return rows

    def close(self):
        if self._h:
            lib().kssd_dist_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def num_cof_batch(mem_limit_bytes: int, n_ref: int, page_sz: int = 4096) -> int:
    """Query rows per batch of the reference's Stage III (command_dist.c:731-734): as many whole pages of uint32[n_ref] rows as
    `-m` allows.  Raises where the reference exits ("at least %fG memory needed")."""
    unit = mem_limit_bytes // (n_ref * 4 * page_sz)
    if unit < 1:
        raise KssdError(-1, f"at least {n_ref * 4 * page_sz / 1073741824:f}G memory needed to map ./onedist, specify more memory use -m")
    return unit * page_sz


def batched_search(ctx: Context, ref_indexes: Sequence[Index], qcodes: Sequence[np.ndarray], qindex: Sequence[np.ndarray], qry_ctx_ct: np.ndarray,
                   ref_ctx_ct: np.ndarray, rows_per_batch: int, sparse: bool = False, fetch_counts: bool = False, **stats_opts):
    """mco_cbdco_nobin_dist's batch loop (command_dist.c:763-790) on the GPU: the queries are searched `rows_per_batch` at a time,
    so neither the Q x R count matrix nor the row list of one job ever has to hold the whole search (a 100k x 100k all-vs-all
    is 40 GB of counts and 10^10 rows at -D 1).  Yields (first_query, counts block or None, statistics rows) per batch, in query
    order -- concatenated, the rows are those of one unbatched job; `cmprsn_num` is that of the whole search (:1186).
    qcodes / qindex: per component, as for DistJob.accumulate."""
    q_sz = np.ascontiguousarray(qry_ctx_ct, dtype=np.uint32)
    r_sz = np.ascontiguousarray(ref_ctx_ct, dtype=np.uint32)
    n_qry, n_ref = q_sz.size, r_sz.size
    stats_opts.setdefault("cmprsn_num", (n_ref * n_qry) & 0xFFFFFFFF)
    for lo in range(0, n_qry, rows_per_batch):
        hi = min(lo + rows_per_batch, n_qry)
        job = DistJob(ctx, q_sz[lo:hi], r_sz, sparse=sparse)
        try:
            for ix, qc, qi in zip(ref_indexes, qcodes, qindex):
                qi = np.ascontiguousarray(qi, dtype=np.uint64)
                a, b = int(qi[lo]), int(qi[hi])
                job.accumulate(ix, np.ascontiguousarray(qc, dtype=np.uint32)[a:b], qi[lo:hi + 1] - qi[lo])
            rows = job.stats(**stats_opts)
            rows["qry"] += lo
            yield lo, (job.counts() if fetch_counts else None), rows
        finally:
            job.close()
