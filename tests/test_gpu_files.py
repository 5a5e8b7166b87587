"""GPU parity: Stage I straight from files (kssd_stage1_files) == the batch API on the same bytes, for plain and .gz
inputs, many small batches, one file larger than a batch, FASTQ modes; and against the reference goldens."""
import gzip
import sys
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import capi, synth

GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu


def _write(tmp, files, gz_every=2):
    paths = []
    for i, (name, data) in enumerate(sorted(files.items())):
        raw = data.tobytes()
        if gz_every and i % gz_every == gz_every - 1:
            p = tmp / f"{name}.fa.gz"
            with gzip.open(p, "wb", compresslevel=1) as f:
                f.write(raw)
        else:
            p = tmp / f"{name}.fa"
            p.write_bytes(raw)
        paths.append(p)
    return paths


@pytest.mark.parametrize("batch_bytes,threads", [(0, 0), (300_000, 3), (1 << 20, 1)])
def test_files_match_batch_api_and_reference(gpu_ctx_l3k10, tmp_path, batch_bytes, threads):
    files = cases.fasta_inputs()
    paths = _write(tmp_path, files)
    sk, t = gpu_ctx_l3k10.sketch_files(paths, threads=threads, batch_bytes=batch_bytes)
    want = gpu_ctx_l3k10.sketch([files[n] for n in sorted(files)])
    assert np.array_equal(sk.index[0], want.index[0]) and np.array_equal(sk.ids[0], want.ids[0])
    assert t["bytes"] == sum(v.size for v in files.values()) and t["batches"] >= (1 if batch_bytes == 0 else 3)
    g = np.load(GOLD / "fasta_l3k10.npz", allow_pickle=False)
    for i, n in enumerate(sorted(files)):
        a, b = int(sk.index[0][i]), int(sk.index[0][i + 1])
        assert np.array_equal(sk.ids[0][a:b], np.sort(g[f"{n}.0"]))       # same set as the reference wrote


def test_files_fastq_modes(shuf_s5, tmp_path):
    from public_kssd_b200 import kssd
    ctx = kssd.Context(8, 5, 2, shuf_s5)
    try:
        files = cases.fastq_inputs()
        paths = _write(tmp_path, files)
        names = sorted(files)
        sk, _ = ctx.sketch_files(paths, mode=capi.MODE_FASTQ, Q=40, M=2, batch_bytes=2_000_000, threads=2)
        want = ctx.sketch_fastq([files[n] for n in names], Q=40, M=2)
        assert np.array_equal(sk.index[0], want.index[0]) and np.array_equal(sk.ids[0], want.ids[0])
        sk, _ = ctx.sketch_files(paths, mode=capi.MODE_FASTQ_ABUND, threads=2)
        want = ctx.sketch_fastq([files[n] for n in names], abundance=True)
        assert np.array_equal(sk.ids[0], want.ids[0]) and np.array_equal(sk.abund[0], want.abund[0])
    finally:
        ctx.close()


def test_files_errors(gpu_ctx_l3k10, tmp_path):
    from public_kssd_b200 import kssd
    with pytest.raises(kssd.KssdError):
        gpu_ctx_l3k10.sketch_files([tmp_path / "missing.fa"])
    bad = tmp_path / "trunc.fa"
    bad.write_bytes(b">h\nACGTACGTACGTACGTACGTACGTACGTACGT\n>header without end")
    with pytest.raises(kssd.KssdError):
        gpu_ctx_l3k10.sketch_files([bad])                                    # the reference exits on this file too
    sk, _ = gpu_ctx_l3k10.sketch_files([bad], strict=False)
    assert sk.status[0] == capi.E_HEADER_EOF


def test_files_through_pipe_command_and_list_file(gpu_ctx_l3k10, tmp_path):
    """-P <cmd>: the files are read from the stdout of "<cmd> <file>" (here xz-like: `gzip -dc`); -l list file."""
    from public_kssd_b200 import hostfmt
    files = {n: v for n, v in cases.fasta_inputs().items() if n in ("a_plain80", "d_messy", "f_short_lines")}
    paths = _write(tmp_path, files, gz_every=1)           # every file compressed
    paths = [p if p.suffix == ".gz" else p for p in paths]
    gz_only = [p for p in paths if p.suffix == ".gz"]
    lst = tmp_path / "inputs.list"
    lst.write_text("\n".join(str(p) for p in gz_only) + "\n\n")
    listed = hostfmt.read_list_file(lst)
    assert listed == [str(p) for p in gz_only]
    sk, _ = gpu_ctx_l3k10.sketch_files(listed, pipecmd="gzip -dc", threads=2)
    want, _ = gpu_ctx_l3k10.sketch_files(listed)
    assert np.array_equal(sk.ids[0], want.ids[0]) and np.array_equal(sk.index[0], want.index[0]) and len(sk.ids[0]) > 0
    from public_kssd_b200 import kssd
    with pytest.raises(kssd.KssdError):
        gpu_ctx_l3k10.sketch_files(listed, pipecmd="false")


def test_unreadable_file_fails_loudly(gpu_ctx_l3k10, tmp_path):
    """A path stat() accepts but read() rejects (here: a directory -> EISDIR) must fail the call, as the reference's
    'eof or fread error' does -- never return sketches made from whatever the staging buffer held."""
    from public_kssd_b200 import kssd
    good = tmp_path / "a.fa"
    good.write_bytes(synth.to_fasta(synth.random_bases(200_000, 5), "a", 80).tobytes())
    bad = tmp_path / "b.fa"
    bad.mkdir()
    (bad / "x").write_bytes(b"y" * 10)
    for order in ([good, bad, good], [bad], [good, good, bad]):
        with pytest.raises(kssd.KssdError):
            gpu_ctx_l3k10.sketch_files(order, batch_bytes=150_000)
    sk, _ = gpu_ctx_l3k10.sketch_files([good, good])          # the context is still usable
    assert len(sk.ids[0]) > 0


def _many_fasta(n_files, seed):
    files = {}
    for i in range(n_files):
        src = synth.random_bases(20_000 + 3_000 * (i % 7), seed + i)
        half = src.size // 2
        files[f"g{i:03d}"] = np.concatenate([synth.to_fasta(src[:half], f"contig_{i}_a", width=60 if i % 2 else 80),
                                             synth.to_fasta(src[half:], f"contig_{i}_b", width=70)])
    return files


def test_gz_inflated_on_the_gpu(gpu_ctx_l3k10, tmp_path, monkeypatch):
    """.gz files copied to the device as they are and inflated there (csrc/inflate.cuh, one file per thread): same sketch as zlib on
    the host and as the batch API on the decoded bytes; plain files may sit in the same batch; a handful of files;
    small batches (KSSD_GZ_BATCH_BYTES); a two-member file and a damaged one go through the host path (which reports the damage)."""
    from public_kssd_b200 import kssd
    files = _many_fasta(80, 100)
    files["zz_empty"] = np.frombuffer(b"", dtype=np.uint8)
    names = sorted(files)
    paths = []
    for i, n in enumerate(names):
        raw = files[n].tobytes()
        if i % 9 == 4:
            p = tmp_path / f"{n}.fa"
            p.write_bytes(raw)
        else:
            p = tmp_path / f"{n}.fa.gz"
            with gzip.open(p, "wb", compresslevel=[1, 6, 9][i % 3]) as f:
                f.write(raw)
        paths.append(p)
    want = gpu_ctx_l3k10.sketch([files[n] for n in names])
    host, th = gpu_ctx_l3k10.sketch_files(paths, threads=4)      # 72 .gz files: below the library's threshold (640), zlib on the host
    assert not th["gz_on_gpu"]
    monkeypatch.setenv("KSSD_GZ_GPU", "1")
    for batch in (None, "1048576"):      # (KSSD_GZ_BATCH_BYTES has a floor of 1 MiB: about 2 MB of text -> 2 batches)
        if batch:
            monkeypatch.setenv("KSSD_GZ_BATCH_BYTES", batch)
        sk, t = gpu_ctx_l3k10.sketch_files(paths, threads=4)
        assert t["gz_on_gpu"] and t["bytes"] == sum(v.size for v in files.values()) and (t["batches"] == 1 if not batch else t["batches"] >= 2)
        for got in (sk, host):
            assert np.array_equal(got.index[0], want.index[0]) and np.array_equal(got.ids[0], want.ids[0])
    monkeypatch.delenv("KSSD_GZ_BATCH_BYTES")
    few = [p for p in paths if p.suffix == ".gz"][:5]
    sk5, t5 = gpu_ctx_l3k10.sketch_files(few, threads=2)
    assert t5["gz_on_gpu"]
    idx = [paths.index(p) for p in few]
    for j, i in enumerate(idx):
        assert np.array_equal(sk5.ids[0][int(sk5.index[0][j]):int(sk5.index[0][j + 1])], want.ids[0][int(want.index[0][i]):int(want.index[0][i + 1])])
    # two gzip members in one file: ISIZE is the last member's -> the call falls back to zlib, same result as the decoded bytes
    two = tmp_path / "two_members.fa.gz"
    a, b = files[names[0]].tobytes(), files[names[1]].tobytes()
    two.write_bytes(gzip.compress(a, 6) + gzip.compress(b, 1))
    sk2, t2 = gpu_ctx_l3k10.sketch_files([two] + few, threads=2)
    assert not t2["gz_on_gpu"]
    w2 = gpu_ctx_l3k10.sketch([np.frombuffer(a + b, dtype=np.uint8)])
    assert np.array_equal(sk2.ids[0][:int(sk2.index[0][1])], w2.ids[0])
    # damage inside the stream: an error either way, never a sketch of garbage
    dmg = bytearray(few[0].read_bytes())
    dmg[len(dmg) // 2] ^= 0x5a
    bad = tmp_path / "damaged.fa.gz"
    bad.write_bytes(bytes(dmg))
    with pytest.raises(kssd.KssdError):
        gpu_ctx_l3k10.sketch_files([bad] + few, threads=2)


def test_many_gz_files_switch_the_gpu_decoder_on(gpu_ctx_l3k10, tmp_path):
    """640 or more .gz files in one call: the library inflates on the GPU without being told to."""
    datas, paths = [], []
    for i in range(650):
        d = synth.to_fasta(synth.random_bases(400 + i % 50, 7000 + i), f"s{i}", width=70)
        p = tmp_path / f"s{i:04d}.fa.gz"
        p.write_bytes(gzip.compress(d.tobytes(), 1))
        datas.append(d); paths.append(p)
    sk, t = gpu_ctx_l3k10.sketch_files(paths, threads=4)
    want = gpu_ctx_l3k10.sketch(datas)
    assert t["gz_on_gpu"] and np.array_equal(sk.index[0], want.index[0]) and np.array_equal(sk.ids[0], want.ids[0])
