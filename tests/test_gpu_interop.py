"""GPU: file-format interop with the UNMODIFIED reference binary (drop-in boundary).
  (1) sketch directories written from GPU results are indexed and searched by the reference;
  (2) an index written from GPU results (mco.0 + the 2 GiB dense mco.index.0) is searched by the reference;
both must give the same distance.out as the all-GPU path formatted by hostfmt."""
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref" / "kssd"
REF_MC = ROOT / "oracle" / "_ref" / "kssd_mc"


def _rows(text):
    out = {}
    for ln in text.splitlines()[1:]:
        f = ln.split("\t")
        out[(Path(f[0]).name, Path(f[1]).name)] = f[2:]
    return out


def test_reference_consumes_gpu_written_files(gpu_ctx_l3k10):
    if not REF.exists():
        pytest.skip("oracle/_ref/kssd not built")
    from public_kssd_b200 import hostfmt, kssd
    ctx = gpu_ctx_l3k10
    anc = synth.random_bases(400_000, 71)
    refs = {f"r{i}.fasta": synth.to_fasta(synth.mutate(anc, 0.004 * i, 80 + i), f"r{i}", 80) for i in range(5)}
    qrys = {f"q{i}.fasta": synth.to_fasta(synth.mutate(anc, 0.01 + 0.02 * i, 90 + i), f"q{i}", 70) for i in range(3)}
    rs, qs = ctx.sketch(list(refs.values())), ctx.sketch(list(qrys.values()))
    work = Path(tempfile.mkdtemp(prefix="kssd_interop_", dir="/dev/shm"))
    try:
        rdir, qdir = work / "ref", work / "qry"
        hostfmt.write_sketch_dir(rdir, ctx.shuf_id, ctx.k, ctx.drlevel, rs, list(refs))
        hostfmt.write_sketch_dir(qdir, ctx.shuf_id, ctx.k, ctx.drlevel, qs, list(qrys))
        # all-GPU answer
        ix = ctx.combco2mco(rs.ids[0], rs.index[0])
        job = kssd.DistJob(ctx, qs.ctx_ct(), rs.ctx_ct())
        job.accumulate(ix, qs.ids[0], qs.index[0])
        mine = _rows(hostfmt.distance_out_header(0, 2) + hostfmt.format_stat_rows(job.stats(), list(qrys), list(refs), 0, 2))
        # (2) our index files, the reference's search
        gdir = work / "gpu_index"
        shutil.copytree(rdir, gdir)
        uc, uo, gids = ix.csr()
        hostfmt.write_mco(gdir, 0, gids, ix.dense())
        hostfmt.write_mcofiles_stat(gdir, ctx.shuf_id, 2 * ctx.k, 2 * ctx.drlevel, 1, rs.ctx_ct(), list(refs))
        r = subprocess.run([str(REF), "dist", "-p", "2", "-r", str(gdir), "-o", str(work / "out2"), str(qdir)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-300:]
        assert _rows((work / "out2" / "distance.out").read_text()) == mine
        # (1) the reference indexes our sketch directory itself, then searches
        ok = False
        for binary in (REF, REF, REF_MC):          # the reference's index stage dies now and then (see oracle.RefRun.index)
            for f in rdir.glob("mco*"):
                f.unlink()
            r = subprocess.run([str(binary), "dist", "-p", "2", "-o", str(rdir), str(rdir)], capture_output=True, text=True)
            if r.returncode == 0 and (rdir / "mcofiles.stat").exists():
                ok = True
                break
        assert ok, r.stderr[-300:]
        assert np.array_equal(np.fromfile(rdir / "mco.0", dtype="<u4"), gids)                 # same postings file
        r = subprocess.run([str(REF), "dist", "-p", "2", "-r", str(rdir), "-o", str(work / "out1"), str(qdir)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-300:]
        assert _rows((work / "out1" / "distance.out").read_text()) == mine
        job.close(); ix.close()
    finally:
        shutil.rmtree(work, ignore_errors=True)
