"""Deterministic synthetic inputs (genomes, FASTA/FASTQ text, sketches) for tests and bench.py.

Everything is a pure function of integer seeds through a splitmix64 counter hash, so the same bytes
come out on every box and numpy version (golden fixtures under tests/golden/ depend on that).
Shapes follow SURVEY.md s8(d): clusters of genomes derived from an ancestor by i.i.d. substitutions
(mirrors test_fna's seq_mutX), 80-column FASTA, optional N runs / soft-masking / many contigs.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _stream(seed: int, n: int, salt: int = 0) -> np.ndarray:
    """n 64-bit pseudo-random words for (seed, salt)."""
    base = np.uint64((seed * 0x9E3779B97F4A7C15 + salt * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        return splitmix64(np.arange(n, dtype=np.uint64) + base)


def make_shuf_table(subk: int, seed: int) -> np.ndarray:
    """Deterministic permutation of 0..16^subk-1 usable as a .shuf payload."""
    n = 1 << (4 * subk)
    h = _stream(seed, n, salt=77)
    order = np.argsort(h, kind="stable")
    perm = np.empty(n, dtype=np.int32)
    perm[order] = np.arange(n, dtype=np.int32)
    return perm


def random_bases(n: int, seed: int) -> np.ndarray:
    """uint8 codes 0..3."""
    return (_stream(seed, n, salt=1) >> np.uint64(33)).astype(np.uint8) & 3


def mutate(bases: np.ndarray, rate: float, seed: int) -> np.ndarray:
    """i.i.d. substitutions at `rate` (always to a different base)."""
    r = _stream(seed, bases.size, salt=2)
    hit = (r >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53)) < rate
    shift = ((r & np.uint64(0xFF)) % np.uint64(3)).astype(np.uint8) + 1
    out = bases.copy()
    out[hit] = (out[hit] + shift[hit]) & 3
    return out


_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def to_fasta(bases: np.ndarray, name: str = "seq", width: int = 80, crlf: bool = False) -> np.ndarray:
    """One-contig FASTA text as a uint8 array: '>name\\n' + lines of `width` + final newline."""
    n = bases.size
    seq = _ACGT[bases]
    eol = b"\r\n" if crlf else b"\n"
    nl = len(eol)
    if width <= 0:
        body = np.concatenate([seq, np.frombuffer(eol, dtype=np.uint8)])
    else:
        nlines = (n + width - 1) // width
        body = np.empty(n + nlines * nl, dtype=np.uint8)
        full = n // width
        if full:
            blk = body[: full * (width + nl)].reshape(full, width + nl)
            blk[:, :width] = seq[: full * width].reshape(full, width)
            blk[:, width:] = np.frombuffer(eol, dtype=np.uint8)
        rem = n - full * width
        if rem:
            tail = body[full * (width + nl):]
            tail[:rem] = seq[full * width:]
            tail[rem:] = np.frombuffer(eol, dtype=np.uint8)
    hdr = np.frombuffer(b">" + name.encode() + eol, dtype=np.uint8)
    return np.concatenate([hdr, body])


def messy_fasta(nbases: int, seed: int, ncontigs: int = 7, width: int = 60, n_rate: float = 0.002,
                lower_frac: float = 0.2, crlf: bool = False, iupac: bool = True) -> np.ndarray:
    """Many-contig FASTA with N runs, soft-masked stretches, IUPAC codes, odd bytes and blank lines --
    the shapes SURVEY.md s8a S1 / A5-A6 list as edge cases."""
    parts = []
    per = max(nbases // ncontigs, 50)
    for c in range(ncontigs):
        b = random_bases(per + 13 * c, seed * 1000 + c)
        txt = _ACGT[b].copy()
        r = _stream(seed * 1000 + c, txt.size, salt=5)
        u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        # soft-masked blocks
        blk = (np.arange(txt.size) // 97) % 5 == (c % 5)
        if lower_frac > 0:
            txt[blk] |= 0x20
        # N runs: start with prob n_rate/8, length 1..16
        starts = np.nonzero(u < n_rate / 8)[0]
        for s0 in starts:
            ln = int(r[s0] & np.uint64(15)) + 1
            txt[s0:s0 + ln] = ord("N") if (int(r[s0]) >> 8) & 1 else ord("n")
        if iupac:
            odd = np.nonzero((u > 0.5) & (u < 0.5 + n_rate / 4))[0]
            alphabet = np.frombuffer(b"RYKMSWBDHVryk*-.0 9", dtype=np.uint8)
            txt[odd] = alphabet[(r[odd] >> np.uint64(20)).astype(np.int64) % alphabet.size]
        eol = b"\r\n" if crlf else b"\n"
        w = width + (c % 3) * 10
        lines = [b">contig_%d some description > with gt ACGTACGTACGTACGTACGTACGTACGT len=%d" % (c, txt.size) + eol]
        tb = txt.tobytes()
        for i in range(0, len(tb), w):
            lines.append(tb[i:i + w] + eol)
        if c % 2 == 1:
            lines.append(eol)           # blank line between contigs
        parts.append(b"".join(lines))
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def to_read_fasta(bases: np.ndarray, n_reads: int, seed: int, min_len: int = 30, max_len: int = 3000, width: int = 70,
                  messy: bool = False, leading_sequence: bool = False) -> np.ndarray:
    """FASTA-formatted reads (the input of `--byread`, reference reads2mco): n_reads records cut from `bases` at
    pseudo-random places, lengths in [min_len, max_len], every third record on a single line.  messy: N / lower-case
    stretches, a '>' in the middle of a sequence line, CRLF records, blank lines, an empty record.
    leading_sequence: sequence text before the first header (it lands in record 0 of the reference's index)."""
    r = _stream(seed, 3 * n_reads + 3, salt=17)
    parts = []
    if leading_sequence:
        parts.append(_ACGT[bases[:200]].tobytes() + b"\n")
    for i in range(n_reads):
        ln = min_len + int(r[3 * i] % np.uint64(max_len - min_len + 1))
        st = int(r[3 * i + 1] % np.uint64(max(bases.size - ln, 1)))
        txt = _ACGT[bases[st:st + ln]].copy()
        flags = int(r[3 * i + 2] & np.uint64(0xffff))
        eol = b"\n"
        if messy:
            if flags & 1:
                txt[ln // 3: ln // 3 + 1 + (flags >> 8) % 40] = ord("N")
            if flags & 2:
                txt[ln // 2:] |= 0x20
            if flags & 4 and ln > 80:
                txt[ln // 4] = ord(">")                       # header start in the middle of a sequence line
            if flags & 8:
                eol = b"\r\n"
            if (flags & 0xf0) == 0xf0:
                txt = txt[:0]                                 # empty record
        parts.append(b">read_%d/1 len=%d" % (i, txt.size) + eol)
        tb = txt.tobytes()
        if i % 3 == 0 or width <= 0:
            parts.append(tb + eol)
        else:
            for j in range(0, len(tb), width):
                parts.append(tb[j:j + width] + eol)
        if messy and (flags & 0x300) == 0x300:
            parts.append(eol)
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def cluster_genomes(n_genomes: int, genome_len: int, seed: int, cluster_size: int = 20,
                    min_rate: float = 0.001, max_rate: float = 0.1):
    """Yield (name, bases) for n_genomes genomes arranged as clusters of mutated copies of an ancestor."""
    n_clusters = (n_genomes + cluster_size - 1) // cluster_size
    g = 0
    for c in range(n_clusters):
        anc = random_bases(genome_len, seed * 7919 + c)
        for m in range(cluster_size):
            if g >= n_genomes:
                return
            if m == 0:
                b = anc
            else:
                rate = min_rate * (max_rate / min_rate) ** ((m - 1) / max(cluster_size - 2, 1))
                b = mutate(anc, rate, seed * 104729 + g)
            yield f"c{c}_m{m}", b
            g += 1


def to_fastq(bases: np.ndarray, n_reads: int, read_len: int, seed: int, err: float = 0.005,
             n_rate: float = 0.001, qual_lo: int = 35, qual_hi: int = 73, trailing_newline: bool = True) -> np.ndarray:
    """4-line FASTQ of reads sampled uniformly from `bases` (forward strand only), Phred+33 qualities."""
    r = _stream(seed, n_reads, salt=9)
    starts = (r % np.uint64(max(bases.size - read_len, 1))).astype(np.int64)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    rb = bases[idx]
    rr = _stream(seed, n_reads * read_len, salt=10).reshape(n_reads, read_len)
    u = (rr >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    sub = u < err
    rb = np.where(sub, (rb + 1 + ((rr & np.uint64(0xFF)) % np.uint64(3)).astype(np.uint8)) & 3, rb)
    txt = _ACGT[rb]
    txt = np.where((u > 0.5) & (u < 0.5 + n_rate), np.uint8(ord("N")), txt)
    qual = (qual_lo + ((rr >> np.uint64(40)) % np.uint64(qual_hi - qual_lo + 1)).astype(np.int64)).astype(np.uint8)
    recs = []
    tb, qb = txt.tobytes(), qual.tobytes()
    for i in range(n_reads):
        recs.append(b"@r%d\n" % i + tb[i * read_len:(i + 1) * read_len] + b"\n+\n" + qb[i * read_len:(i + 1) * read_len] + b"\n")
    out = b"".join(recs)
    if not trailing_newline:
        out = out[:-1]
    return np.frombuffer(out, dtype=np.uint8).copy()


def synth_sketches(n_genomes: int, codes_per_genome: int, seed: int, code_bits: int = 28, cluster_size: int = 20,
                   min_div: float = 0.001, max_div: float = 0.1, klen: int = 20, member_seed: int | None = None):
    """Sketches without sequences (SURVEY.md s8d cfg 3): per cluster an ancestor set of codes; each member keeps an
    ancestor code with probability (1-d)^klen and replaces the rest with fresh random codes.
    Returns (codes uint32 concatenated, index uint64[n+1]); each genome's codes are sorted unique.
    member_seed: same ancestors (they depend on `seed` only), different members -- independent query batches against one
    reference set."""
    ms = seed if member_seed is None else member_seed
    n_clusters = (n_genomes + cluster_size - 1) // cluster_size
    mask = np.uint64((1 << code_bits) - 1)
    chunks, counts = [], []
    g = 0
    for c in range(n_clusters):
        anc = np.unique((_stream(seed * 31 + c, codes_per_genome, salt=20) & mask).astype(np.uint32))
        for m in range(cluster_size):
            if g >= n_genomes:
                break
            if m == 0:
                s = anc
            else:
                d = min_div * (max_div / min_div) ** ((m - 1) / max(cluster_size - 2, 1))
                keep_p = (1.0 - d) ** klen
                r = _stream(ms * 1009 + g, anc.size, salt=21)
                u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
                fresh = (_stream(ms * 2003 + g, anc.size, salt=22) & mask).astype(np.uint32)
                s = np.unique(np.where(u < keep_p, anc, fresh))
            chunks.append(s)
            counts.append(s.size)
            g += 1
    index = np.zeros(n_genomes + 1, dtype=np.uint64)
    np.cumsum(np.asarray(counts, dtype=np.uint64), out=index[1:])
    return np.concatenate(chunks).astype(np.uint32), index


def make_shuf_table_affine(subk: int, seed: int) -> np.ndarray:
    """A permutation of 0..16^subk-1 built from bijective steps (odd multiply, xor-shift, add) instead of an argsort:
    seconds for the 1 GiB subk = 7 table of `-L 4 -k 10` (reference auto rule, command_shuffle.c:154-160).  The
    reference accepts any permutation as .shuf payload."""
    bits = 4 * subk
    mask = np.uint32((1 << bits) - 1) if bits < 32 else np.uint32(0xFFFFFFFF)
    a = np.uint32(((0x9E3779B1 + 2 * seed * 0x632BE5AB) | 1) & 0xFFFFFFFF)
    b = np.uint32((0x85EBCA6B * (seed + 1)) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        x = np.arange(1 << bits, dtype=np.uint32)
        x = (x * a) & mask
        x ^= x >> np.uint32(bits // 2 + 1)
        x = (x * np.uint32(0xC2B2AE3D | 1)) & mask
        x ^= x >> np.uint32(bits // 3 + 1)
        x = (x + b) & mask
    return x.view(np.int32)


def contig_genome(nbases: int, seed: int, contig_len: int = 100_000, width: int = 60, n_frac: float = 0.01,
                  lower_frac: float = 0.2, crlf: bool = False) -> np.ndarray:
    """BASELINE configs[3]-shaped FASTA text, vectorised: many contigs (lengths spread around contig_len, one header
    each), runs of N summing to ~n_frac of the bases (run lengths 1 .. 5000), soft-masked (lower-case) stretches,
    fixed-width lines."""
    bases = random_bases(nbases, seed)
    txt = _ACGT[bases].copy()
    r = _stream(seed, nbases // 512 + 16, salt=31)
    if lower_frac > 0:                                    # soft-masked blocks of 512 bases
        blk = ((r[: (nbases + 511) // 512] >> np.uint64(20)).astype(np.float64) * (1.0 / (1 << 44))) < lower_frac
        m = np.repeat(blk, 512)[:nbases]
        txt[m] |= 0x20
    target, done, i = int(nbases * n_frac), 0, 0
    rn = _stream(seed, 4096, salt=32)
    while done < target and i + 1 < rn.size:
        ln = 1 + int(rn[i] % np.uint64(5000)) if (int(rn[i]) >> 40) & 3 else 1 + int(rn[i] % np.uint64(40))
        st = int(rn[i + 1] % np.uint64(max(nbases - ln, 1)))
        txt[st:st + ln] = ord("N") if (int(rn[i]) >> 50) & 1 else ord("n")
        done += ln
        i += 2
    parts, pos, c = [], 0, 0
    rc = _stream(seed, nbases // max(contig_len // 4, 1) + 8, salt=33)
    eol = np.frombuffer(b"\r\n" if crlf else b"\n", dtype=np.uint8)
    while pos < nbases:
        ln = int(contig_len // 4 + rc[c] % np.uint64(max(contig_len * 3 // 2, 1)))
        ln = min(ln, nbases - pos)
        seg = txt[pos:pos + ln]
        full = ln // width
        body = np.empty(ln + ((ln + width - 1) // width) * eol.size, dtype=np.uint8)
        if full:
            v = body[: full * (width + eol.size)].reshape(full, width + eol.size)
            v[:, :width] = seg[: full * width].reshape(full, width)
            v[:, width:] = eol
        rem = ln - full * width
        if rem:
            t = body[full * (width + eol.size):]
            t[:rem] = seg[full * width:]
            t[rem:] = eol
        parts.append(np.frombuffer(b">ctg%06d len=%d ACGTACGTTGCATGCATGCAAGCT\n" % (c, ln), dtype=np.uint8))
        parts.append(body)
        pos += ln
        c += 1
    return np.concatenate(parts)


def synth_sketches_torch(n_genomes: int, codes_per_genome: int, seed: int, device, code_bits: int = 28, cluster_size: int = 20,
                         min_div: float = 0.001, max_div: float = 0.1, klen: int = 20, member_seed: int | None = None,
                         block_clusters: int = 512):
    """synth_sketches, evaluated with torch on `device` (bench data at configs[2] size in a second instead of 20 s of
    numpy calls per genome): the SAME codes and index, element for element (tests/test_synth_torch.py).
    Returns (codes int32 tensor [bit pattern of the uint32 codes], index int64 tensor[n+1])."""
    import torch

    M64 = 0xFFFFFFFFFFFFFFFF

    def s64(v: int) -> int:                      # python int -> the int64 with the same bit pattern
        v &= M64
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(x, s):                               # logical shift right of int64 bit patterns
        return (x >> s) & ((1 << (64 - s)) - 1)

    def mix(x):
        z = x + s64(0x9E3779B97F4A7C15)
        z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
        z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
        return z ^ lsr(z, 31)

    def streams(seeds, n, salt):                 # rows: _stream(seed, n, salt) for every seed (python ints)
        base = torch.tensor([s64(sd * 0x9E3779B97F4A7C15 + salt * 0xD1B54A32D192ED03) for sd in seeds], dtype=torch.int64, device=device)
        return mix(base[:, None] + torch.arange(n, dtype=torch.int64, device=device)[None, :])

    ms = seed if member_seed is None else member_seed
    n_clusters = (n_genomes + cluster_size - 1) // cluster_size
    mask = (1 << code_bits) - 1
    SENT = 1 << 40                               # sorts after every code
    n = codes_per_genome

    def sort_unique_rows(v):                     # rows sorted, repeats pushed to the end as SENT
        v, _ = torch.sort(v, dim=1)
        dup = torch.zeros_like(v, dtype=torch.bool)
        dup[:, 1:] = (v[:, 1:] == v[:, :-1]) & (v[:, 1:] != SENT)
        v = torch.where(dup, torch.full_like(v, SENT), v)
        v, _ = torch.sort(v, dim=1)
        return v

    keep_ps = []
    for m in range(cluster_size):
        if m == 0:
            keep_ps.append(2.0)
        else:
            d = min_div * (max_div / min_div) ** ((m - 1) / max(cluster_size - 2, 1))
            keep_ps.append((1.0 - d) ** klen)
    keep_row = torch.tensor(keep_ps, dtype=torch.float64, device=device)

    out_codes, out_counts = [], []
    for c0 in range(0, n_clusters, block_clusters):
        c1 = min(c0 + block_clusters, n_clusters)
        nc = c1 - c0
        anc = sort_unique_rows(streams([seed * 31 + c for c in range(c0, c1)], n, 20) & mask)          # [nc, n], SENT-padded
        gids = [c * cluster_size + m for c in range(c0, c1) for m in range(cluster_size)]
        r = streams([ms * 1009 + g for g in gids], n, 21)
        u = lsr(r, 11).to(torch.float64) * (1.0 / (1 << 53))
        fresh = streams([ms * 2003 + g for g in gids], n, 22) & mask
        ancm = anc[:, None, :].expand(nc, cluster_size, n).reshape(nc * cluster_size, n)
        kp = keep_row[None, :].expand(nc, cluster_size).reshape(nc * cluster_size, 1)
        s = torch.where(u < kp, ancm, fresh)
        s = torch.where(ancm == SENT, ancm, s)                                                       # streams are cut at anc.size
        s = sort_unique_rows(s)
        live = torch.tensor([g < n_genomes for g in gids], dtype=torch.bool, device=device)
        s = s[live]
        valid = s != SENT
        out_codes.append(s[valid].to(torch.int32))                                                    # codes < 2^28: same bits
        out_counts.append(valid.sum(dim=1))
    codes = torch.cat(out_codes)
    counts = torch.cat(out_counts)
    index = torch.zeros(n_genomes + 1, dtype=torch.int64, device=device)
    index[1:] = torch.cumsum(counts, dim=0)
    return codes, index
