"""ctypes binding of libkssd_b200.so (include/kssd_b200.h).

This is the only way Python reaches the CUDA kernels: there is no CPU implementation behind these
calls.  Importing works without a GPU (so the symbol table can be checked on a CPU box); creating a
context without a usable sm_100 device raises KssdError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("KSSD_B200_LIB", PKG_DIR / "libkssd_b200.so"))

# every symbol include/kssd_b200.h declares (tests/test_capi_symbols.py checks header <-> this list <-> .so)
SYMBOLS = [
    "kssd_last_error", "kssd_version", "kssd_kernel_launch_count",
    "kssd_ctx_create", "kssd_ctx_destroy", "kssd_ctx_info", "kssd_ctx_stream", "kssd_ctx_sync", "kssd_ctx_last_ms",
    "kssd_sketch_batch_host", "kssd_sketch_batch_dev", "kssd_sketch_count", "kssd_sketch_status", "kssd_sketch_fetch",
    "kssd_sketch_dev_ptrs", "kssd_sketch_stats", "kssd_sketch_free", "kssd_sketch_read_counts", "kssd_sketch_fetch_read_index",
    "kssd_stage1_files", "kssd_stage1_files_ex", "kssd_stage1_count", "kssd_stage1_fetch", "kssd_stage1_status", "kssd_stage1_timing", "kssd_stage1_gz_info", "kssd_gunzip_host", "kssd_stage1_free",
    "kssd_index_build_host", "kssd_index_build_dev", "kssd_index_sizes", "kssd_index_fetch", "kssd_index_fetch_dense",
    "kssd_index_from_dense_host", "kssd_index_free",
    "kssd_dist_create", "kssd_dist_create_ext", "kssd_dist_accumulate_host", "kssd_dist_accumulate_dev", "kssd_dist_fetch_counts",
    "kssd_dist_counts_dev", "kssd_dist_create_sparse", "kssd_dist_sparse_add_dev", "kssd_dist_sparse_add_host", "kssd_dist_accumulate_peer", "kssd_dev_alloc", "kssd_dev_zero", "kssd_dev_free",
    "kssd_ipc_export", "kssd_ipc_open", "kssd_ipc_close", "kssd_dist_stats", "kssd_dist_stats_async", "kssd_dist_stats_wait", "kssd_dist_fetch_stats", "kssd_dist_free",
    "kssd_format_distance_rows", "kssd_dist_format_text", "kssd_dist_text", "kssd_format_distance_rows_gpu", "kssd_format_selftest", "kssd_host_free",
    "kssd_set_union_host", "kssd_set_union_dev", "kssd_set_operate_host", "kssd_set_operate_dev", "kssd_set_group_host",
    "kssd_composite_host",
]

MODE_FASTA, MODE_FASTA_UNIQ, MODE_FASTQ, MODE_FASTQ_ABUND, MODE_BYREAD = 0, 1, 2, 3, 4
METRIC_JACCARD, METRIC_CONTAINMENT = 0, 1

E_CROWD, E_HEADER_EOF, E_LONGLINE = -4, -5, -9

COMP_ROW_DTYPE = np.dtype([("qry", "<u4"), ("ref", "<u4"), ("kmer_num", "<u4"), ("median", "<u4"), ("max", "<u4"), ("mean", "<f4"),
                           ("pct", "<f4"), ("reserved", "<u4")])


class KssdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kssd_b200 error {code}: {msg}")
        self.code = code


class CtxInfo(C.Structure):
    _fields_ = [("k", C.c_int32), ("subk", C.c_int32), ("drlevel", C.c_int32), ("component_sz", C.c_int32),
                ("component_num", C.c_int32), ("comp_code_bits", C.c_int32), ("dim_end", C.c_uint32),
                ("hashsize", C.c_uint32), ("hashlimit", C.c_uint32), ("n_sampled", C.c_uint32),
                ("device", C.c_int32), ("sm_count", C.c_int32)]


class SketchOpts(C.Structure):
    _fields_ = [("mode", C.c_int32), ("Q", C.c_int32), ("M", C.c_int32), ("want_ord", C.c_int32),
                ("span_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class StatOpts(C.Structure):
    _fields_ = [("metric", C.c_int32), ("correction", C.c_int32), ("kmerlen", C.c_int32), ("dim_rd_len", C.c_int32),
                ("dthreshold", C.c_double), ("n_neighbors", C.c_int32), ("skip_zero", C.c_int32), ("cmprsn_num", C.c_uint64)]


STAT_ROW_DTYPE = np.dtype([("qry", "<u4"), ("ref", "<u4"), ("shared", "<u4"), ("rs_u", "<u4"), ("ref_size", "<u4"),
                           ("qry_size", "<u4"), ("metric", "<f8"), ("dist", "<f8"), ("pvalue", "<f8"), ("fdr", "<f8"),
                           ("ci_metric_lo", "<f8"), ("ci_metric_hi", "<f8"), ("ci_dist_lo", "<f8"), ("ci_dist_hi", "<f8")])

_lib = None


def build_library(force: bool = False) -> Path:
    """nvcc-compile the library in-tree for sm_100a (cross-compiles without a GPU)."""
    src = PKG_DIR / "csrc"
    newest = max(p.stat().st_mtime for p in list(src.glob("*.cu")) + list(src.glob("*.cuh")) + list(src.glob("*.inc")) + [PKG_DIR.parent / "include" / "kssd_b200.h"])
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        r = subprocess.run(["make", "-C", str(src)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libkssd_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def sources_digest() -> str:
    """sha256 (first 16 hex digits) over the library's sources: csrc/*.cu, *.cuh, the Makefile and the C-ABI header."""
    import hashlib
    src = PKG_DIR / "csrc"
    h = hashlib.sha256()
    for f in sorted(list(src.glob("*.cu")) + list(src.glob("*.cuh")) + list(src.glob("*.inc")) + [src / "Makefile", PKG_DIR.parent / "include" / "kssd_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def library_provenance() -> dict:
    """What a bench line says about the binary it ran: the .so's own digest, when it was built, and whether it is newer than every
    source it was built from (build() rebuilds it otherwise)."""
    import hashlib
    src = PKG_DIR / "csrc"
    newest = max(p.stat().st_mtime for p in list(src.glob("*.cu")) + list(src.glob("*.cuh")) + list(src.glob("*.inc")) + [PKG_DIR.parent / "include" / "kssd_b200.h"])
    st = LIB_PATH.stat()
    return {"path": str(LIB_PATH.relative_to(PKG_DIR.parent)), "sha256_16": hashlib.sha256(LIB_PATH.read_bytes()).hexdigest()[:16], "bytes": st.st_size,
            "built_unix": int(st.st_mtime), "newer_than_sources": bool(st.st_mtime >= newest), "sources_sha256_16": sources_digest(),
            "flags": "nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a (csrc/Makefile)", "version": lib().kssd_version().decode()}


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise KssdError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            f"(there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, u8p, u16p, u32p, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_uint32), \
        C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    L.kssd_last_error.restype = C.c_char_p
    L.kssd_version.restype = C.c_char_p
    L.kssd_kernel_launch_count.restype = C.c_uint64
    L.kssd_ctx_create.argtypes = [C.POINTER(vp), C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.kssd_ctx_destroy.argtypes = [vp]
    L.kssd_ctx_destroy.restype = None
    L.kssd_ctx_info.argtypes = [vp, C.POINTER(CtxInfo)]
    L.kssd_ctx_stream.argtypes = [vp]
    L.kssd_ctx_stream.restype = vp
    L.kssd_ctx_sync.argtypes = [vp]
    L.kssd_ctx_last_ms.argtypes = [vp, C.c_int]
    L.kssd_ctx_last_ms.restype = C.c_float
    for name in ("kssd_sketch_batch_host", "kssd_sketch_batch_dev"):
        getattr(L, name).argtypes = [vp, vp, C.c_size_t, u64p, u64p, C.c_int, C.POINTER(SketchOpts), C.POINTER(vp)]
    L.kssd_sketch_count.argtypes = [vp, C.c_int]
    L.kssd_sketch_count.restype = C.c_int64
    L.kssd_sketch_status.argtypes = [vp, i32p]
    L.kssd_sketch_fetch.argtypes = [vp, C.c_int, u32p, u64p, u16p, u64p]
    L.kssd_sketch_read_counts.argtypes = [vp, u64p]
    L.kssd_sketch_fetch_read_index.argtypes = [vp, C.c_int, C.c_int, u64p]
    L.kssd_stage1_files.argtypes = [vp, C.POINTER(C.c_char_p), C.c_int, C.POINTER(SketchOpts), C.c_int, C.c_size_t, C.POINTER(vp)]
    L.kssd_stage1_files_ex.argtypes = [vp, C.POINTER(C.c_char_p), C.c_int, C.POINTER(SketchOpts), C.c_int, C.c_size_t, C.c_char_p, C.POINTER(vp)]
    L.kssd_stage1_count.argtypes = [vp, C.c_int]
    L.kssd_stage1_count.restype = C.c_int64
    L.kssd_stage1_fetch.argtypes = [vp, C.c_int, u32p, u64p, u16p]
    L.kssd_stage1_status.argtypes = [vp, i32p]
    L.kssd_stage1_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    L.kssd_stage1_gz_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.kssd_gunzip_host.argtypes = [C.c_char_p, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.kssd_stage1_free.argtypes = [vp]
    L.kssd_stage1_free.restype = None
    L.kssd_sketch_dev_ptrs.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.kssd_sketch_stats.argtypes = [vp, u64p, C.POINTER(C.c_float)]
    L.kssd_sketch_free.argtypes = [vp]
    L.kssd_sketch_free.restype = None
    L.kssd_index_build_host.argtypes = [vp, u32p, u64p, C.c_int, C.POINTER(vp)]
    L.kssd_index_build_dev.argtypes = [vp, vp, vp, C.c_int, C.c_uint64, C.POINTER(vp)]
    L.kssd_index_sizes.argtypes = [vp, u64p, u64p, C.POINTER(C.c_int)]
    L.kssd_index_fetch.argtypes = [vp, u32p, u64p, u32p]
    L.kssd_index_fetch_dense.argtypes = [vp, u64p]
    L.kssd_index_from_dense_host.argtypes = [vp, u64p, u32p, C.c_uint64, C.c_int, C.POINTER(vp)]
    L.kssd_index_free.argtypes = [vp]
    L.kssd_index_free.restype = None
    L.kssd_dist_create.argtypes = [vp, C.c_int, C.c_int, u32p, u32p, C.POINTER(vp)]
    L.kssd_dist_create_ext.argtypes = [vp, C.c_int, C.c_int, u32p, u32p, vp, C.c_int, C.POINTER(vp)]
    L.kssd_dist_accumulate_host.argtypes = [vp, vp, u32p, u64p]
    L.kssd_dist_accumulate_dev.argtypes = [vp, vp, vp, vp, C.c_uint64]
    L.kssd_dist_accumulate_peer.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.c_int, C.c_int]
    L.kssd_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.kssd_dev_zero.argtypes = [vp, vp, C.c_size_t]
    L.kssd_dev_free.argtypes = [vp, vp]
    L.kssd_dev_free.restype = None
    L.kssd_ipc_export.argtypes = [vp, vp, u8p]
    L.kssd_ipc_open.argtypes = [vp, u8p, C.POINTER(vp)]
    L.kssd_ipc_close.argtypes = [vp, vp]
    L.kssd_dist_create_sparse.argtypes = [vp, C.c_int, C.c_int, u32p, u32p, C.POINTER(vp)]
    L.kssd_dist_sparse_add_dev.argtypes = [vp, vp, vp, vp, C.c_uint64]
    L.kssd_dist_sparse_add_host.argtypes = [vp, vp, u32p, u64p]
    L.kssd_dist_fetch_counts.argtypes = [vp, u32p]
    L.kssd_dist_counts_dev.argtypes = [vp]
    L.kssd_dist_counts_dev.restype = vp
    L.kssd_dist_stats.argtypes = [vp, C.POINTER(StatOpts)]
    L.kssd_dist_stats.restype = C.c_int64
    L.kssd_dist_stats_async.argtypes = [vp, C.POINTER(StatOpts)]
    L.kssd_dist_stats_wait.argtypes = [vp]
    L.kssd_dist_stats_wait.restype = C.c_int64
    L.kssd_dist_fetch_stats.argtypes = [vp, vp]
    L.kssd_dist_free.argtypes = [vp]
    L.kssd_dist_free.restype = None
    L.kssd_format_distance_rows.argtypes = [vp, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.kssd_dist_format_text.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.kssd_dist_text.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.kssd_format_distance_rows_gpu.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.kssd_format_selftest.argtypes = [C.c_uint64, C.c_uint64, u64p]
    L.kssd_format_selftest.restype = C.c_int64
    L.kssd_set_union_host.argtypes = [vp, u32p, C.c_uint64, C.c_int, u32p, u64p]
    L.kssd_set_union_dev.argtypes = [vp, vp, C.c_uint64, C.c_int, vp, C.c_uint64, u64p]
    L.kssd_set_operate_host.argtypes = [vp, u32p, u64p, C.c_int, u32p, C.c_uint64, C.c_int, u32p, u64p]
    L.kssd_set_operate_dev.argtypes = [vp, vp, vp, C.c_int, C.c_uint64, vp, C.c_uint64, C.c_int, vp, vp]
    L.kssd_set_group_host.argtypes = [vp, u32p, u64p, C.c_int, u32p, u64p, C.c_int, u32p, u64p]
    L.kssd_composite_host.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int,
                                      C.POINTER(vp), u64p]
    L.kssd_host_free.argtypes = [vp]
    L.kssd_host_free.restype = None
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise KssdError(int(rc), lib().kssd_last_error().decode(errors="replace"))
    return rc


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))
