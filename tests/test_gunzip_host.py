"""csrc/inflate.cuh -- the gzip decoder the GPU runs, one file per thread -- executed on the host (kssd_gunzip_host) against
Python's zlib: every block type, header option and strategy, several members, damage."""
import ctypes as C
import gzip
import zlib

import numpy as np
import pytest

from public_kssd_b200 import capi


def _gunzip(data: bytes, cap: int):
    out = (C.c_uint8 * max(cap, 1))()
    n = C.c_size_t()
    rc = capi.lib().kssd_gunzip_host(data, len(data), out, cap, C.byref(n))
    return rc, bytes(out[:n.value]) if rc == 0 else b""


def _gz(data: bytes, level: int, strategy: int) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    return co.compress(data) + co.flush()


def _fasta(n: int, width: int, seed: int) -> bytes:
    rng = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)]
    rep = rng.integers(0, n - 200, 40)
    for r in rep:                                             # repeats: matches as well as literals
        seq[r + 100:r + 180] = seq[r:r + 80]
    lines = [b">contig_1 synthetic"] + [seq[i:i + width].tobytes() for i in range(0, n, width)]
    return b"\n".join(lines) + b"\n"


INPUTS = [b"", b"A", _fasta(5_000, 60, 1), _fasta(400_000, 80, 2), bytes(np.random.default_rng(3).integers(0, 256, 100_000, dtype=np.uint8)),
          b"A" * 300_000, b"ACGTTGCA\n" * 20_000]


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
def test_decoder_matches_zlib(level, strategy):
    for data in INPUTS:
        z = _gz(data, level, strategy)
        rc, out = _gunzip(z, len(data))
        assert rc == 0 and out == data, (len(data), level, strategy, rc)


def test_header_fields_members_and_damage(tmp_path):
    a, b = _fasta(50_000, 70, 4), _fasta(1_234, 60, 5)
    p = tmp_path / "named.fna.gz"
    with gzip.GzipFile(filename=str(p), mode="wb", compresslevel=6, mtime=12345) as f:      # FNAME in the header
        f.write(a)
    za = p.read_bytes()
    assert za[3] & 8
    rc, out = _gunzip(za, len(a))
    assert rc == 0 and out == a
    zb = _gz(b, 1, zlib.Z_DEFAULT_STRATEGY)
    rc, out = _gunzip(za + zb + b"\0" * 13, len(a) + len(b))             # two members and zero padding
    assert rc == 0 and out == a + b
    assert _gunzip(za, len(a) - 1)[0] == -3                                # output full
    assert _gunzip(za[:-9], len(a))[0] != 0                                # truncated
    bad = bytearray(za)
    bad[len(bad) // 2] ^= 0x10
    rc, out = _gunzip(bytes(bad), len(a))
    assert rc != 0                                                         # data error, wrong size or CRC mismatch
    bad = bytearray(za)
    bad[-6] ^= 1                                                           # the stored CRC itself
    assert _gunzip(bytes(bad), len(a))[0] == -5
    assert _gunzip(b"not a gzip file at all....", 100)[0] == -1


def test_decoder_on_random_mixtures():
    """200 buffers of random structure (alphabet size, repeat density, run lengths, size 0 .. 300 KB), random level / strategy /
    window size: the decoder returns exactly what zlib compressed, and never accepts a buffer with a flipped bit as the original."""
    rng = np.random.default_rng(2024)
    for case in range(200):
        n = int(rng.integers(0, 300_000)) if case % 5 else int(rng.integers(0, 64))
        alpha = int(rng.choice([2, 4, 5, 20, 256]))
        data = rng.integers(0, alpha, n, dtype=np.uint8) + (65 if alpha < 200 else 0)
        for _ in range(int(rng.integers(0, 30))):               # repeats at random distances (short and beyond 8 KiB)
            if n < 100:
                break
            ln = int(rng.integers(3, min(400, n // 2)))
            src = int(rng.integers(0, n - ln))
            dst = int(rng.integers(0, n - ln))
            data[dst:dst + ln] = data[src:src + ln].copy()
        if case % 7 == 0 and n > 10:
            data[: n // 2] = data[0]                           # a long run
        raw = data.astype(np.uint8).tobytes()
        co = zlib.compressobj(int(rng.integers(0, 10)), zlib.DEFLATED, 16 + int(rng.integers(9, 16)), int(rng.integers(1, 10)),
                              int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED])))
        z = co.compress(raw) + co.flush()
        rc, out = _gunzip(z, len(raw))
        assert rc == 0 and out == raw, case
        if len(z) > 30:
            bad = bytearray(z)
            bad[int(rng.integers(10, len(z) - 8))] ^= 1 << int(rng.integers(0, 8))
            rc, out = _gunzip(bytes(bad), len(raw))
            assert rc != 0 or out == raw, case                 # (a flip inside the header's mtime / OS bytes changes nothing)
