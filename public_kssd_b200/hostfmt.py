"""Host-side file formats and ordering helpers of the reference -- no compute kernels.

combco.<c> in the reference is in HASH-SLOT order (iseq2comem.c:538-546).  The GPU returns each
genome's ids ascending plus the byte offset of every id's first occurrence; `slot_order` replays the
open-addressing insertion of the distinct keys in first-occurrence order, which reproduces the
reference file byte for byte.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np


def slot_order(ids: np.ndarray, first_occurrence: np.ndarray, hashsize: int, keys: np.ndarray | None = None) -> np.ndarray:
    """ids of ONE genome (one component when comp_num == 1) -> the order wrt_co2cmpn_use_inn_subctx writes them.
    keys: the full drtuple values when they differ from ids (comp_num > 1); default = ids."""
    keys = ids.astype(np.uint64) if keys is None else keys.astype(np.uint64)
    order = np.argsort(first_occurrence, kind="stable")
    H = int(hashsize)
    table = {}
    for j in order:
        key = int(keys[j])
        h1, h2 = key % H, 1 + key % (H - 1)
        i = 0
        while True:
            n = (h1 + i * h2) % H
            if n not in table:
                table[n] = int(ids[j])
                break
            i += 1
    return np.array([table[s] for s in sorted(table)], dtype=np.uint32)


def format_composite_rows(rows, qry_names, ref_names) -> str:
    """The text `kssd composite` prints (command_composite.c:531): qry, ref, kmer_num, mean, 98-99 % band mean, median, max."""
    return "".join("%s\t%s\t%d\t%f\t%f\t%d\t%d\n" % (qry_names[int(r["qry"])], ref_names[int(r["ref"])], r["kmer_num"], float(r["mean"]),
                                                     float(r["pct"]), r["median"], r["max"]) for r in rows)


def read_list_file(path) -> list:
    """The reference's `-l <list>`: one input path per line (command_dist.c organize step); blank lines ignored."""
    return [ln.strip() for ln in Path(path).read_text().splitlines() if ln.strip()]


def _names_block(names) -> bytes:
    out = []
    for nm in names:
        b = nm.encode()[:255]
        out.append(b + b"\0" * (256 - len(b)))
    return b"".join(out)


def write_cofiles_stat(d, shuf_id: int, koc: bool, kmerlen: int, dim_rd_len: int, comp_num: int, ctx_ct: np.ndarray, names) -> None:
    """co_dstat_t (global_basic.h:94-103) + ctx_ct + 256-byte names (command_dist.c:361-377)."""
    with open(Path(d, "cofiles.stat"), "wb") as f:
        f.write(struct.pack("<I?xxxiiiiQ", shuf_id & 0xFFFFFFFF, bool(koc), kmerlen, dim_rd_len, comp_num, len(names),
                            int(np.sum(ctx_ct, dtype=np.uint64))))
        f.write(np.ascontiguousarray(ctx_ct, dtype="<u4").tobytes())
        f.write(_names_block(names))


def write_mcofiles_stat(d, shuf_id: int, kmerlen: int, dim_rd_len: int, comp_num: int, ctx_ct: np.ndarray, names) -> None:
    """mco_dstat_t (command_dist.h:57-64) + ctx_ct + names (command_dist.c:397-409)."""
    with open(Path(d, "mcofiles.stat"), "wb") as f:
        f.write(struct.pack("<Iiiii", shuf_id & 0xFFFFFFFF, kmerlen, dim_rd_len, comp_num, len(names)))
        f.write(np.ascontiguousarray(ctx_ct, dtype="<u4").tobytes())
        f.write(_names_block(names))


def read_cofiles_stat(d) -> dict:
    raw = Path(d, "cofiles.stat").read_bytes()
    shuf_id, koc, kmerlen, dim_rd_len, comp_num, infile_num, all_ctx_ct = struct.unpack_from("<I?xxxiiiiQ", raw, 0)
    ct = np.frombuffer(raw, dtype="<u4", count=infile_num, offset=32).copy()
    base = 32 + 4 * infile_num
    names = [raw[base + 256 * i: base + 256 * (i + 1)].split(b"\0")[0].decode() for i in range(infile_num)]
    return dict(shuf_id=shuf_id, koc=koc, kmerlen=kmerlen, dim_rd_len=dim_rd_len, comp_num=comp_num, infile_num=infile_num,
                all_ctx_ct=all_ctx_ct, ctx_ct=ct, names=names)


def read_mcofiles_stat(d) -> dict:
    raw = Path(d, "mcofiles.stat").read_bytes()
    shuf_id, kmerlen, dim_rd_len, comp_num, infile_num = struct.unpack_from("<Iiiii", raw, 0)
    ct = np.frombuffer(raw, dtype="<u4", count=infile_num, offset=20).copy()
    base = 20 + 4 * infile_num
    names = [raw[base + 256 * i: base + 256 * (i + 1)].split(b"\0")[0].decode() for i in range(infile_num)]
    return dict(shuf_id=shuf_id, kmerlen=kmerlen, dim_rd_len=dim_rd_len, comp_num=comp_num, infile_num=infile_num, ctx_ct=ct,
                names=names)


def write_combco(d, comp: int, ids: np.ndarray, index: np.ndarray, abund: np.ndarray | None = None) -> None:
    np.ascontiguousarray(ids, dtype="<u4").tofile(Path(d, f"combco.{comp}"))
    np.ascontiguousarray(index, dtype="<u8").tofile(Path(d, f"combco.index.{comp}"))
    if abund is not None:
        np.ascontiguousarray(abund, dtype="<u2").tofile(Path(d, f"combco.{comp}.a"))


def read_combco(d, comp: int):
    ids = np.fromfile(Path(d, f"combco.{comp}"), dtype="<u4")
    index = np.fromfile(Path(d, f"combco.index.{comp}"), dtype="<u8")
    ap = Path(d, f"combco.{comp}.a")
    return ids, index, (np.fromfile(ap, dtype="<u2") if ap.exists() else None)


def write_mco(d, comp: int, gids: np.ndarray, dense_incl: np.ndarray) -> None:
    np.ascontiguousarray(gids, dtype="<u4").tofile(Path(d, f"mco.{comp}"))
    np.ascontiguousarray(dense_incl, dtype="<u8").tofile(Path(d, f"mco.index.{comp}"))


def read_mco(d, comp: int):
    return np.fromfile(Path(d, f"mco.{comp}"), dtype="<u4"), np.fromfile(Path(d, f"mco.index.{comp}"), dtype="<u8")


# ------------------------------------------------------------------------------------------------
# distance.out text (dist_print_nobin / output_ctrl, command_dist.c:1188-1195, :1268-1285)
# ------------------------------------------------------------------------------------------------
_HEADER = (("Jaccard\tMashD", "P-value(J)\tFDR(J)", "Jaccard_CI\tMashD_CI"),
           ("ContainmentM\tAafD", "P-value(C)\tFDR(C)", "ContainmentM_CI\tAafD_CI"))


def _c_double(v: float, spec: str) -> str:
    """glibc printf of a double for '%.6lf' (spec 'f') and '%E' (spec 'E'), including -nan / inf spellings."""
    import math
    if math.isnan(v):
        neg = math.copysign(1.0, v) < 0
        s = "nan" if spec == "f" else "NAN"
        return ("-" if neg else "") + s
    if math.isinf(v):
        s = "inf" if spec == "f" else "INF"
        return ("-" if v < 0 else "") + s
    return ("%.6f" % v) if spec == "f" else ("%E" % v)


def distance_out_header(metric: int, outfields: int) -> str:
    return "Qry\tRef\tShared_k|Ref_s|Qry_s" + "".join("\t" + _HEADER[metric][i] for i in range(outfields + 1)) + "\n"


def format_stat_rows(rows, qry_names, ref_names, metric: int = 0, outfields: int = 2) -> str:
    """rows: numpy structured array from DistJob.stats() -> the body of distance.out (one line per row)."""
    out = []
    for r in rows:
        s = "%s\t%s\t%d-%d|%d|%d\t%s\t%s" % (qry_names[int(r["qry"])], ref_names[int(r["ref"])], r["shared"], r["rs_u"], r["ref_size"],
                                           r["qry_size"], _c_double(float(r["metric"]), "f"), _c_double(float(r["dist"]), "f"))
        if outfields >= 1:
            s += "\t%s\t%s" % (_c_double(float(r["pvalue"]), "E"), _c_double(float(r["fdr"]), "E"))
        if outfields >= 2:
            s += "\t[%s,%s]\t[%s,%s]" % tuple(_c_double(float(r[n]), "f") for n in ("ci_metric_lo", "ci_metric_hi", "ci_dist_lo", "ci_dist_hi"))
        out.append(s + "\n")
    return "".join(out)


def format_distance_out(rows, qry_names, ref_names, metric: int = 0, outfields: int = 2, header: bool = True, threads: int = 0) -> bytes:
    """distance.out through the library's native multi-threaded formatter (kssd_format_distance_rows): header +
    one line per row, byte-identical to dist_print_nobin / output_ctrl (command_dist.c:1188-1195, 1267-1285).
    `format_stat_rows` above is the plain-Python statement of the same text, kept for the tests."""
    import ctypes as C
    from .capi import check, lib
    rows = np.ascontiguousarray(rows)
    assert rows.dtype.itemsize == 88, "rows must be kssd_stat_row_t records"
    qn, rn = _names_block(qry_names), _names_block(ref_names)
    text, n = C.c_void_p(), C.c_size_t()
    check(lib().kssd_format_distance_rows(rows.ctypes.data_as(C.c_void_p), len(rows), qn, rn, 256, metric, outfields, int(header), threads,
                                          C.byref(text), C.byref(n)))
    try:
        return C.string_at(text, n.value)
    finally:
        lib().kssd_host_free(text)


# ------------------------------------------------------------------------------------------------
# sketch directories: what run_stageI leaves behind, and combine_queries (command_dist.c:1323-1475)
# ------------------------------------------------------------------------------------------------
def write_sketch_dir(d, shuf_id: int, k: int, drlevel: int, sketch, names, koc: bool = False) -> None:
    """cofiles.stat + combco.<c> + combco.index.<c> (+ combco.<c>.a) from a kssd.Sketch (run_stageI's output,
    command_dist.c:314-378).  Genome order is the order of `names` (the reference shuffles its own)."""
    d = Path(d)
    d.mkdir(parents=True, exist_ok=True)
    for c in range(len(sketch.ids)):
        write_combco(d, c, sketch.ids[c], sketch.index[c], sketch.abund[c] if koc else None)
    write_cofiles_stat(d, shuf_id, koc, 2 * k, 2 * drlevel, len(sketch.ids), sketch.ctx_ct(), names)


def combine_queries(dirs, out) -> dict:
    """Concatenate sketch directories that share a shuf_id into `out`, as `kssd dist -o out dirA dirB ...` does:
    combco.<c> appended, combco.index.<c> rebased, ctx_ct lists and names appended, header counts summed.
    Directories with another shuf_id or with abundances are skipped with a message, like the reference."""
    out = Path(out)
    out.mkdir(parents=True, exist_ok=True)
    first = read_cofiles_stat(dirs[0])
    if first["koc"]:
        raise ValueError("combine_queries(): abundance model not supported yet")
    comp = first["comp_num"]
    cts, names = [first["ctx_ct"]], list(first["names"])
    all_ctx = first["all_ctx_ct"]
    codes = [[read_combco(dirs[0], c)[0]] for c in range(comp)]
    index = [[read_combco(dirs[0], c)[1]] for c in range(comp)]
    for i, dq in enumerate(dirs[1:], start=1):
        try:
            st = read_cofiles_stat(dq)
        except FileNotFoundError:
            print(f"{i}th query {dq} is not a valid query: no cofiles.stat file")
            continue
        if st["shuf_id"] != first["shuf_id"]:
            print(f"combine_queries(): {i}th shuf_id: {st['shuf_id']} not match 0th shuf_id: {first['shuf_id']}")
            continue
        if st["koc"]:
            print(f"combine_queries(): {i}th query abundance model not supported yet ")
            continue
        all_ctx += st["all_ctx_ct"]
        cts.append(st["ctx_ct"])
        names += st["names"]
        for c in range(comp):
            ids, ix, _ = read_combco(dq, c)
            codes[c].append(ids)
            index[c].append(ix[1:] + index[c][-1][-1])
    ct = np.concatenate(cts)
    for c in range(comp):
        write_combco(out, c, np.concatenate(codes[c]), np.concatenate(index[c]))
    with open(out / "cofiles.stat", "wb") as f:
        f.write(struct.pack("<I?xxxiiiiQ", first["shuf_id"], False, first["kmerlen"], first["dim_rd_len"], comp, len(names), int(all_ctx)))
        f.write(np.ascontiguousarray(ct, dtype="<u4").tobytes())
        f.write(_names_block(names))
    return read_cofiles_stat(out)


# ---- kssd set -g / -c: the host-only parts (grouping file, hash-slot order of a group's pan sketch, pan concatenation) ----
PRIMER = (251, 509, 1021, 2039, 4093, 8191, 16381, 32749, 65521, 131071, 262139, 524287, 1048573, 2097143, 4194301, 8388593, 16777213,
          33554393, 67108859, 134217689, 268435399, 536870909, 1073741789, 2147483647, 4294967291)      # global_basic.c:74


def _next_prime(n: int) -> int:
    """global_basic.c:389 (trial division up to (int)sqrt(n))."""
    while True:
        if all(n % j for j in range(2, int(np.sqrt(n)) + 1)):
            return n
        n += 1


def organize_taxf(text: str):
    """The reference's grouping file (organize_taxf, command_set.c:536-600): line i = "<taxid>[\t<taxname>]" for genome i.  Returns the
    groups in the order the reference walks them -- the slot order of its taxon hash table (double hashing on the taxid, table size
    nextPrime(lines / 0.6)) -- as dicts(taxid, taxname, gids); taxid 0 marks genomes that belong to no group (still listed: callers skip them)."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    ln = len(lines)
    hs = _next_prime(int(ln / 0.6))
    table = {}
    for i, line in enumerate(lines):
        parts = [p for p in line.split("\t") if p != ""]          # strtok skips empty fields
        taxid = int(parts[0]) if parts else 0                       # (atoi of a number)
        taxname = parts[1] if len(parts) > 1 else None
        for n in range(hs):
            hv = (taxid % hs + n * (1 + taxid % (hs - 1))) % hs
            if hv not in table:
                table[hv] = {"taxid": taxid, "taxname": taxname, "gids": [i]}
                break
            if table[hv]["taxid"] == taxid:
                if table[hv]["taxname"] != taxname:
                    raise ValueError(f"organize_taxf: taxid {taxid} has different taxnames in lines {table[hv]['gids'][0]} and {i}")
                table[hv]["gids"].append(i)
                break
    return [table[k] for k in sorted(table)], ln


def group_hashsize(n_member_codes: int) -> int:
    """size of a group's hash table in grouping_genomes (command_set.c:735-737): primer[LOG2(codes * 1.5) - 7] (codes with repeats)."""
    x = int(n_member_codes * 1.5)
    ind = x.bit_length() - 1 if x > 0 else 0
    return PRIMER[ind - 7] if ind > 7 else PRIMER[0]


def group_slot_order(codes_first_occurrence: np.ndarray, n_member_codes: int) -> np.ndarray:
    """A group's distinct codes in first-occurrence order (Context.set_group) -> the order grouping_genomes writes them: the slots of
    its open-addressing table (HASH = (K % H + x (1 + K % (H - 1))) % H with 32-bit unsigned arithmetic, command_set.c:738-752).  A
    code equal to 0 is the table's empty marker and is never written."""
    H = group_hashsize(n_member_codes)
    table = {}
    for key in np.asarray(codes_first_occurrence, dtype=np.uint32).tolist():
        if key == 0:
            continue
        h1, h2 = key % H, 1 + key % (H - 1)
        for x in range(H):
            y = ((h1 + ((x * h2) & 0xFFFFFFFF)) & 0xFFFFFFFF) % H
            if y not in table:
                table[y] = key
                break
            if table[y] == key:
                break
    return np.array([table[k] for k in sorted(table)], dtype=np.uint32)


def group_names(groups) -> list:
    """names grouping_genomes writes to cofiles.stat: "<taxid>_<taxname>" or "<taxid>" (command_set.c:794-799)."""
    return [f"{g['taxid']}_{g['taxname']}" if g["taxname"] is not None else str(g["taxid"]) for g in groups if g["taxid"] != 0]


def combine_pans(pans_per_input) -> tuple:
    """`kssd set -c` (combin_pans, command_set.c:444-512) for one component: the pan.<c> / uniq_pan.<c> arrays of the inputs
    concatenated into one combco.<c> with its combco.index.<c>; host-only in the reference too."""
    arrs = [np.ascontiguousarray(p, dtype=np.uint32) for p in pans_per_input]
    index = np.zeros(len(arrs) + 1, dtype=np.uint64)
    index[1:] = np.cumsum([a.size for a in arrs])
    return (np.concatenate(arrs) if arrs else np.zeros(0, np.uint32)), index
