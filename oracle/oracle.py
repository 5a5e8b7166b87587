"""oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of oracle/_build/liboracle.so (the plain-C restatement, kssd_oracle.c), plus
helpers that run the UNMODIFIED reference binary oracle/_ref/kssd and parse the files it writes.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; public_kssd_b200/ never does.

Parity status: pinned -- see tests/test_oracle_vs_ref.py and tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import struct
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"
REF_BIN = HERE / "_ref" / "kssd"
REF_BIN_MC = HERE / "_ref" / "kssd_mc"   # co2mco.c double-free fix, only for comp_num > 1

_lib = None


def build(force: bool = False) -> None:
    """Compile the restatement (and the reference binary when /root/reference is present)."""
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (HERE / "kssd_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    if Path("/root/reference").is_dir() and not (REF_BIN.exists() and REF_BIN_MC.exists()):
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        u8p, u16p, u32p, u64p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int32))
        L.orc_ctx_sizeof.restype = C.c_size_t
        L.orc_ctx_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
        L.orc_fasta2co.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_int, u64p]
        L.orc_fastq2co.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_int, C.c_int, u64p, C.POINTER(C.c_int)]
        L.orc_shortreads2koc.argtypes = [C.c_void_p, u8p, C.c_size_t, u64p]
        L.orc_reads2mco.argtypes = [C.c_void_p, u8p, C.c_size_t, u32p, i32p, u64p, C.c_size_t, u64p]
        L.orc_reads2mco.restype = C.c_long
        L.orc_set_union.argtypes = [u32p, C.c_size_t, C.c_int, C.c_int, u32p]
        L.orc_set_union.restype = C.c_size_t
        L.orc_set_operate.argtypes = [u32p, u64p, C.c_int, u32p, C.c_size_t, C.c_int, C.c_int, u32p, u64p]
        L.orc_set_operate.restype = None
        L.orc_write_co.argtypes = [C.c_void_p, u64p, C.c_int, u32p, i32p, u16p]
        L.orc_write_co.restype = C.c_size_t
        L.orc_combco2mco.argtypes = [u32p, u64p, C.c_int, C.c_int, u64p, u32p]
        L.orc_combco2mco.restype = None
        L.orc_dist_counts_dense.argtypes = [u32p, u64p, C.c_int, u64p, u32p, C.c_int, u32p, C.c_int]
        L.orc_dist_counts_dense.restype = None
        L.orc_dist_counts_csr.argtypes = [u32p, u64p, C.c_int, u32p, u64p, C.c_size_t, u32p, C.c_int, u32p, C.c_int]
        L.orc_dist_counts_csr.restype = None
        L.orc_output_ctrl.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_double, C.c_uint64, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


# ----------------------------------------------------------------------------------------------
# deterministic .shuf tables (any permutation is a valid .shuf payload)
# ----------------------------------------------------------------------------------------------
def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def make_shuf_table(subk: int, seed: int) -> np.ndarray:
    """Deterministic permutation of 0..16^subk-1: rank of a splitmix64 hash (ties impossible in practice,
    broken by index through the stable sort)."""
    n = 1 << (4 * subk)
    with np.errstate(over="ignore"):
        h = splitmix64(np.arange(n, dtype=np.uint64) ^ np.uint64(seed * 0x2545F4914F6CDD1D & 0xFFFFFFFFFFFFFFFF))
    order = np.argsort(h, kind="stable")
    perm = np.empty(n, dtype=np.int32)
    perm[order] = np.arange(n, dtype=np.int32)
    return perm


def write_shuf_file(path, shuf_id: int, k: int, subk: int, drlevel: int, table: np.ndarray) -> None:
    """command_shuffle.c:184-185: 16-byte header {id,k,subk,drlevel} + int32[16^subk]."""
    with open(path, "wb") as f:
        f.write(struct.pack("<iiii", shuf_id, k, subk, drlevel))
        f.write(np.ascontiguousarray(table, dtype="<i4").tobytes())


# ----------------------------------------------------------------------------------------------
# the restatement
# ----------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, k: int, s: int, L: int, shuf: np.ndarray, component_sz: int = 7):
        self.shuf = np.ascontiguousarray(shuf, dtype=np.int32)
        assert self.shuf.size == 1 << (4 * s)
        self._buf = C.create_string_buffer(lib().orc_ctx_sizeof())
        rc = lib().orc_ctx_init(self._buf, k, s, L, component_sz, _p(self.shuf, C.c_int32))
        if rc != 0:
            raise ValueError("get_hashsz(): primer_ind out of range")
        self.k, self.s, self.L, self.component_sz = k, s, L, component_sz
        self.component_num = (1 << (4 * (k - L - component_sz))) if k - L > component_sz else 1
        self.comp_code_bits = 4 * (k - L - component_sz) if k - L > component_sz else 0
        pi = 4 * (k - L) - 8 - 7
        self.hashsize = [251, 509, 1021, 2039, 4093, 8191, 16381, 32749, 65521, 131071, 262139, 524287, 1048573,
                         2097143, 4194301, 8388593, 16777213, 33554393, 67108859, 134217689, 268435399, 536870909,
                         1073741789, 2147483647, 4294967291][pi]
        self._co = np.zeros(self.hashsize, dtype=np.uint64)

    @property
    def ptr(self):
        return C.cast(self._buf, C.c_void_p)

    def _emit(self, mode: int):
        n = int(np.count_nonzero(self._co))
        ids = np.empty(max(n, 1), dtype=np.uint32)
        comp = np.empty(max(n, 1), dtype=np.int32)
        ab = np.empty(max(n, 1), dtype=np.uint16)
        w = lib().orc_write_co(self.ptr, _p(self._co, C.c_uint64), mode, _p(ids, C.c_uint32), _p(comp, C.c_int32),
                               _p(ab, C.c_uint16))
        return ids[:w].copy(), comp[:w].copy(), ab[:w].copy()

    def fasta(self, data: bytes | np.ndarray, uniq: bool = False):
        """-> (ids, comp) in the reference's hash-slot order."""
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        rc = lib().orc_fasta2co(self.ptr, _p(a, C.c_uint8), a.size, int(uniq), _p(self._co, C.c_uint64))
        if rc != 0:
            raise RuntimeError(f"orc_fasta2co rc={rc}")
        ids, comp, _ = self._emit(0)
        return ids, comp

    def byread(self, data):
        """reads2mco (--byread): -> (n_reads, {component: (ids in stream order, index)}) with index = the reference's
        combco.index.<c>: inclusive cumulative counts for record 0 (before the first '>') .. n_reads."""
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        cap = max(int(a.size), 1)
        ids, comp, rd = np.empty(cap, np.uint32), np.empty(cap, np.int32), np.empty(cap, np.uint64)
        nr = C.c_uint64(0)
        n = lib().orc_reads2mco(self.ptr, _p(a, C.c_uint8), a.size, _p(ids, C.c_uint32), _p(comp, C.c_int32), _p(rd, C.c_uint64), cap,
                                C.byref(nr))
        if n < 0:
            raise RuntimeError(f"orc_reads2mco rc={n}")
        ids, comp, rd = ids[:n], comp[:n], rd[:n]
        out = {}
        for c in range(self.component_num):
            m = comp == c
            cnt = np.bincount(rd[m].astype(np.int64), minlength=int(nr.value) + 1)
            out[c] = (ids[m].copy(), np.cumsum(cnt, dtype=np.uint64))
        return int(nr.value), out

    def fastq(self, data, Q: int = 0, M: int = 1):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        rd = C.c_int(0)
        rc = lib().orc_fastq2co(self.ptr, _p(a, C.c_uint8), a.size, Q, M, _p(self._co, C.c_uint64), C.byref(rd))
        if rc != 0:
            raise RuntimeError(f"orc_fastq2co rc={rc}")
        ids, comp, _ = self._emit(1)
        return ids, comp

    def fastq_abund(self, data):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        rc = lib().orc_shortreads2koc(self.ptr, _p(a, C.c_uint8), a.size, _p(self._co, C.c_uint64))
        if rc != 0:
            raise RuntimeError(f"orc_shortreads2koc rc={rc}")
        return self._emit(2)


def sketch_sets(ctx: Ctx, genomes, mode: str = "fasta", **kw):
    """Per genome, per component: sorted unique ids (the canonical form parity is stated on)."""
    out = []
    for g in genomes:
        if mode == "fasta":
            ids, comp = ctx.fasta(g, uniq=kw.get("uniq", False))
        elif mode == "fastq":
            ids, comp = ctx.fastq(g, kw.get("Q", 0), kw.get("M", 1))
        else:
            ids, comp, _ = ctx.fastq_abund(g)
        out.append([np.sort(ids[comp == c]) for c in range(ctx.component_num)])
    return out


def combco2mco(combco: np.ndarray, index: np.ndarray, component_sz: int = 7, dense: bool = False):
    """-> (mco gids, dense inclusive table or None)."""
    combco = np.ascontiguousarray(combco, dtype=np.uint32)
    index = np.ascontiguousarray(index, dtype=np.uint64)
    n = len(index) - 1
    mco = np.empty(max(int(index[-1]), 1), dtype=np.uint32)
    d = np.empty(1 << (4 * component_sz), dtype=np.uint64) if dense else None
    lib().orc_combco2mco(_p(combco, C.c_uint32), _p(index, C.c_uint64), n, component_sz,
                         _p(d, C.c_uint64) if dense else None, _p(mco, C.c_uint32))
    return mco[: int(index[-1])], d


def csr_from_combco(combco: np.ndarray, index: np.ndarray):
    """numpy CSR of the inverted index: (unique codes, exclusive offsets, gids) -- co2mco.c:42-71 semantics."""
    combco = np.asarray(combco, dtype=np.uint32)
    index = np.asarray(index, dtype=np.uint64)
    gid = np.repeat(np.arange(len(index) - 1, dtype=np.uint32), np.diff(index).astype(np.int64))
    order = np.lexsort((gid, combco))
    codes = combco[order]
    ucodes, counts = np.unique(codes, return_counts=True)
    off = np.zeros(len(ucodes) + 1, dtype=np.uint64)
    np.cumsum(counts, out=off[1:])
    return ucodes, off, gid[order]


def dist_counts(qcodes, qindex, ucodes, uoff, mco, refnum, nthreads: int = 1) -> np.ndarray:
    qcodes = np.ascontiguousarray(qcodes, dtype=np.uint32)
    qindex = np.ascontiguousarray(qindex, dtype=np.uint64)
    ucodes = np.ascontiguousarray(ucodes, dtype=np.uint32)
    uoff = np.ascontiguousarray(uoff, dtype=np.uint64)
    mco = np.ascontiguousarray(mco, dtype=np.uint32)
    qn = len(qindex) - 1
    ct = np.zeros((qn, refnum), dtype=np.uint32)
    lib().orc_dist_counts_csr(_p(qcodes, C.c_uint32), _p(qindex, C.c_uint64), qn, _p(ucodes, C.c_uint32),
                              _p(uoff, C.c_uint64), len(ucodes), _p(mco, C.c_uint32), refnum, _p(ct, C.c_uint32),
                              nthreads)
    return ct


def dist_counts_dense(qcodes, qindex, dense_incl, mco, refnum, nthreads: int = 1) -> np.ndarray:
    qcodes = np.ascontiguousarray(qcodes, dtype=np.uint32)
    qindex = np.ascontiguousarray(qindex, dtype=np.uint64)
    mco = np.ascontiguousarray(mco, dtype=np.uint32)
    qn = len(qindex) - 1
    ct = np.zeros((qn, refnum), dtype=np.uint32)
    lib().orc_dist_counts_dense(_p(qcodes, C.c_uint32), _p(qindex, C.c_uint64), qn, _p(dense_incl, C.c_uint64),
                                _p(mco, C.c_uint32), refnum, _p(ct, C.c_uint32), nthreads)
    return ct


def output_ctrl(X, Y, I, metric=0, correction=0, kmerlen=20, dim_rd_len=6, dthreshold=1.0, cmprsn_num=1):
    out = (C.c_double * 9)()
    keep = lib().orc_output_ctrl(int(X), int(Y), int(I), metric, correction, kmerlen, dim_rd_len, dthreshold,
                                 int(cmprsn_num), out)
    return keep, np.array(out[:], dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# running the unmodified reference binary and parsing what it writes (SURVEY.md s8b formats)
# ----------------------------------------------------------------------------------------------
def ref_available() -> bool:
    return REF_BIN.exists()


def set_union(combco: np.ndarray, uniq: bool = False, code_bits: int = 28) -> np.ndarray:
    """kssd set -u / -q for one component: pan.<c> / uniq_pan.<c>."""
    a = np.ascontiguousarray(combco, dtype=np.uint32)
    out = np.empty(max(a.size, 1), dtype=np.uint32)
    n = lib().orc_set_union(_p(a, C.c_uint32), a.size, int(uniq), code_bits, _p(out, C.c_uint32))
    return out[:n].copy()


def set_operate(combco: np.ndarray, index: np.ndarray, pan: np.ndarray, intersect: bool, code_bits: int = 28):
    """kssd set -i / -s <pan> for one component: (combco.<c>, combco.index.<c>) after the filter."""
    a = np.ascontiguousarray(combco, dtype=np.uint32)
    ix = np.ascontiguousarray(index, dtype=np.uint64)
    pn = np.ascontiguousarray(pan, dtype=np.uint32)
    out = np.empty(max(a.size, 1), dtype=np.uint32)
    oix = np.empty(ix.size, dtype=np.uint64)
    lib().orc_set_operate(_p(a, C.c_uint32), _p(ix, C.c_uint64), ix.size - 1, _p(pn, C.c_uint32), pn.size, int(intersect), code_bits,
                          _p(out, C.c_uint32), _p(oix, C.c_uint64))
    return out[:int(oix[-1])].copy(), oix


_PRIMER = (251, 509, 1021, 2039, 4093, 8191, 16381, 32749, 65521, 131071, 262139, 524287, 1048573, 2097143, 4194301, 8388593, 16777213,
           33554393, 67108859, 134217689, 268435399, 536870909, 1073741789, 2147483647, 4294967291)


def set_group(combco: np.ndarray, index: np.ndarray, groups):
    """kssd set -g for one component (grouping_genomes, command_set.c:726-775): for every group (list of genome ids, in the
    reference's order) every code of every member is inserted into an open-addressing table of primer[LOG2(1.5 * codes) - 7]
    slots (0 = empty, so code 0 is lost); the table's non-empty slots are written in slot order.  Returns (combco, index)."""
    codes = np.asarray(combco, dtype=np.uint32)
    ix = np.asarray(index, dtype=np.uint64)
    out, oix = [], [0]
    for gids in groups:
        hashsize = int(sum(int(ix[g + 1] - ix[g]) for g in gids))
        x15 = int(hashsize * 1.5)
        ind = x15.bit_length() - 1 if x15 > 0 else 0
        H = _PRIMER[ind - 7] if ind > 7 else _PRIMER[0]
        table = {}
        for g in gids:
            for key in codes[int(ix[g]):int(ix[g + 1])].tolist():
                h1, h2 = key % H, 1 + key % (H - 1)
                for x in range(H):
                    y = ((h1 + ((x * h2) & 0xFFFFFFFF)) & 0xFFFFFFFF) % H
                    if table.get(y, 0) == 0:
                        table[y] = key
                        break
                    if table[y] == key:
                        break
        row = [table[k] for k in sorted(table) if table[k] != 0]
        out.extend(row)
        oix.append(len(out))
    return np.array(out, dtype=np.uint32), np.array(oix, dtype=np.uint64)


def composite(ref_codes, ref_index, qry_codes, qry_index, qry_abund, min_kmers: int = 6):
    """get_species_abundance (command_composite.c:389-547), numbers only.  Arguments are per-component lists (combco.<c>,
    combco.index.<c>, combco.<c>.a).  For every query, the references that share >= MIN_KM_S (6) k-mers with it, most
    shared first (glibc qsort is a stable merge sort here: ties keep reference order): rows
    (qry, ref, kmer_num, mean, pct, median, max) with mean = (float)sum/kmer_num, pct = mean of the sorted abundances
    a[floor(0.98 k)] .. a[n <= 0.99 k] (1-based), median = a[k/2], max = a[k] (:512-531)."""
    ncomp = len(ref_codes)
    nr, nq = len(ref_index[0]) - 1, len(qry_index[0]) - 1
    rows = []
    for qn in range(nq):
        lists = [[] for _ in range(nr)]
        for c in range(ncomp):
            a, b = int(qry_index[c][qn]), int(qry_index[c][qn + 1])
            km = dict(zip(qry_codes[c][a:b].tolist(), qry_abund[c][a:b].tolist()))
            rc, ri = ref_codes[c], ref_index[c]
            for rn in range(nr):
                for code in rc[int(ri[rn]):int(ri[rn + 1])].tolist():
                    v = km.get(code)
                    if v is not None:
                        lists[rn].append(v)
        for rn in sorted(range(nr), key=lambda i: -len(lists[i])):
            k = len(lists[rn])
            if k < min_kmers:
                break
            a = [0] + sorted(lists[rn])                     # 1-based like ref_abund[rn][1..k]
            total = sum(a)
            lastsum = lastn = 0
            n = int(k * 0.98)
            while n <= k * 0.99:
                lastsum += a[n]
                lastn += 1
                n += 1
            rows.append((qn, rn, k, np.float32(total) / np.float32(k), np.float32(lastsum) / np.float32(lastn), a[k // 2], a[k]))
    return rows


def composite_text(rows, qry_names, ref_names) -> str:
    """the printf at command_composite.c:531"""
    return "".join("%s\t%s\t%d\t%f\t%f\t%d\t%d\n" % (qry_names[q], ref_names[r], k, float(m), float(p), med, mx)
                   for q, r, k, m, p, med, mx in rows)


def run_ref(args, cwd=None, binary=None, timeout=3600) -> subprocess.CompletedProcess:
    return subprocess.run([str(binary or REF_BIN)] + [str(a) for a in args], cwd=cwd, capture_output=True,
                          text=True, timeout=timeout)


def read_cofiles_stat(d):
    """co_dstat_t (global_basic.h:94-103) + ctx_ct[n] + names[n][256] (command_dist.c:361-377)."""
    raw = Path(d, "cofiles.stat").read_bytes()
    shuf_id, koc, kmerlen, dim_rd_len, comp_num, infile_num, all_ctx_ct = struct.unpack_from("<I?xxxiiiiQ", raw, 0)
    ct = np.frombuffer(raw, dtype="<u4", count=infile_num, offset=32)
    names = [raw[32 + 4 * infile_num + 256 * i: 32 + 4 * infile_num + 256 * (i + 1)].split(b"\0")[0].decode()
             for i in range(infile_num)]
    return dict(shuf_id=shuf_id, koc=koc, kmerlen=kmerlen, dim_rd_len=dim_rd_len, comp_num=comp_num,
                infile_num=infile_num, all_ctx_ct=all_ctx_ct, ctx_ct=ct.copy(), names=names)


def read_mcofiles_stat(d):
    """mco_dstat_t (command_dist.h:57-64) + ctx_ct + names (command_dist.c:397-409)."""
    raw = Path(d, "mcofiles.stat").read_bytes()
    shuf_id, kmerlen, dim_rd_len, comp_num, infile_num = struct.unpack_from("<Iiiii", raw, 0)
    ct = np.frombuffer(raw, dtype="<u4", count=infile_num, offset=20)
    names = [raw[20 + 4 * infile_num + 256 * i: 20 + 4 * infile_num + 256 * (i + 1)].split(b"\0")[0].decode()
             for i in range(infile_num)]
    return dict(shuf_id=shuf_id, kmerlen=kmerlen, dim_rd_len=dim_rd_len, comp_num=comp_num,
                infile_num=infile_num, ctx_ct=ct.copy(), names=names)


def read_combco(d, comp: int):
    codes = np.fromfile(Path(d, f"combco.{comp}"), dtype="<u4")
    index = np.fromfile(Path(d, f"combco.index.{comp}"), dtype="<u8")
    ap = Path(d, f"combco.{comp}.a")
    ab = np.fromfile(ap, dtype="<u2") if ap.exists() else None
    return codes, index, ab


def read_distance_out(path):
    """-> header list, rows as list of raw column lists."""
    lines = Path(path).read_text().splitlines()
    return lines[0].split("\t"), [ln.split("\t") for ln in lines[1:]]


class RefRun:
    """Scratch directory driving `kssd dist` stage I / II / III of the unmodified reference."""

    def __init__(self, k, s, L, table, shuf_id=12345, workdir=None):
        self.dir = Path(workdir or tempfile.mkdtemp(prefix="kssdref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
        self.shuf = self.dir / f"L{L}K{k}.shuf"
        write_shuf_file(self.shuf, shuf_id, k, s, L, table)
        self.k, self.s, self.L = k, s, L

    def sketch(self, files_dir, outname, extra=(), p=1):
        """Stage I.  With comp_num > 1 (K11) the reference corrupts its heap intermittently (glibc aborts in
        sysmalloc after the last genome, thread-count dependent; observed here with -p 1 and -p 4, not -p 2);
        the sets it writes when it survives are deterministic, so a crashed run is retried with another -p."""
        out = self.dir / outname
        last = None
        for pp in [p] + [x for x in (2, 3, 5, 1) if x != p]:
            shutil.rmtree(out, ignore_errors=True)
            r = run_ref(["dist", "-p", pp, "-L", self.shuf, "-o", out, *extra, files_dir], cwd=self.dir)
            if r.returncode == 0 and (out / "cofiles.stat").exists():
                return out
            last = r
            if r.returncode != -6:
                break
        raise RuntimeError(f"reference sketch failed rc={last.returncode}: {last.stderr[-500:]} {last.stdout[-300:]}")

    def index(self, sketch_dir, binary=None, p=1):
        """Stage II.  combco2mco frees 16^7 never-initialised pointers (co2mco.c:31,70) and sometimes dies in glibc
        even with one component; a crashed run is retried, then handed to the calloc-patched build (same output)."""
        last = None
        for b in ([binary] if binary else [None, None, REF_BIN_MC]):
            for f in Path(sketch_dir).glob("mco*"):
                f.unlink()
            r = run_ref(["dist", "-p", p, "-o", sketch_dir, sketch_dir], cwd=self.dir, binary=b)
            if r.returncode == 0 and (Path(sketch_dir) / "mcofiles.stat").exists():
                return sketch_dir
            last = r
        raise RuntimeError(f"reference index failed rc={last.returncode}: {last.stderr[-500:]}")

    def dist(self, ref_dir, qry_dir, outname, extra=(), p=1):
        out = self.dir / outname
        r = run_ref(["dist", "-p", p, "-r", ref_dir, "-o", out, *extra, qry_dir], cwd=self.dir)
        if r.returncode != 0:
            raise RuntimeError(f"reference dist failed rc={r.returncode}: {r.stderr[-500:]}")
        return out

    def cleanup(self):
        shutil.rmtree(self.dir, ignore_errors=True)
