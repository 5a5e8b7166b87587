"""GPU: the C host program (host/kssd_b200_dist.c, the reference's language) drives Stage I -> II -> III through the C-ABI
with the reference's file formats; its distance.out must equal the text the unmodified reference wrote for the same
inputs (tests/golden/index_dist_l3k10.npz), and its sharedk_ct.dat the reference's matrix."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from public_kssd_b200 import hostfmt, kssd, synth

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(GOLD))
import cases  # noqa: E402

pytestmark = pytest.mark.gpu


def _norm(text):
    out = []
    for ln in text.splitlines():
        f = ln.split("\t")
        if f[0] != "Qry":
            f[0] = Path(f[0]).name.rsplit(".", 1)[0]
            f[1] = Path(f[1]).name.rsplit(".", 1)[0]
        out.append("\t".join(f))
    return out


def test_c_host_reproduces_reference_distance_out(tmp_path):
    exe = ROOT / "host" / "kssd_b200_dist"
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    g = np.load(GOLD / "index_dist_l3k10.npz", allow_pickle=False)
    fa = cases.fasta_inputs()
    shuf = tmp_path / "L3K10.shuf"
    kssd.write_shuf_file(shuf, cases.SHUF_ID, 10, 6, 3, synth.make_shuf_table(6, cases.SHUF_SEED_S6))
    seq = tmp_path / "seq"
    seq.mkdir()
    for n, b in fa.items():
        (seq / f"{n}.fasta").write_bytes(b.tobytes())
    refs = [str(seq / f"{n}.fasta") for n in g["ref_names"]]          # the order the reference happened to use
    qrys = [str(seq / f"{n}.fasta") for n in g["qry_names"]]

    def run(*args):
        r = subprocess.run([str(exe), *map(str, args)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return r.stdout

    run("sketch", shuf, tmp_path / "ref", *refs)
    run("sketch", shuf, tmp_path / "qry", *qrys)
    run("index", tmp_path / "ref")
    # the files are the reference's formats: same sets per genome, same postings, same dense table
    st = hostfmt.read_cofiles_stat(tmp_path / "ref")
    assert st["shuf_id"] == cases.SHUF_ID and st["kmerlen"] == 20 and st["dim_rd_len"] == 6 and np.array_equal(st["ctx_ct"], g["ref_ctx_ct"])
    mco, dense = hostfmt.read_mco(tmp_path / "ref", 0)
    assert np.array_equal(mco, g["mco"]) and dense[-1] == g["dense_last"][0]
    assert np.array_equal(dense[g["dense_nonzero_codes"]], g["dense_values_at_nonzero"])
    for tag, extra in {"default": [], "M1_O1": ["-M", "1", "-O", "1"], "corr_O2": ["--correction"], "N2_M1": ["-N", "2", "-M", "1"],
                       "D0.1": ["-D", "0.1"], "O0": ["-O", "0"]}.items():
        out = tmp_path / f"dist_{tag}"
        run("dist", tmp_path / "ref", tmp_path / "qry", out, *extra)
        mine = (out / "distance.out").read_text()
        assert _norm(mine) == _norm(g[f"distance_out.{tag}"].tobytes().decode()), tag
    ct = np.fromfile(tmp_path / "dist_default" / "sharedk_ct.dat", dtype="<u4").reshape(len(qrys), len(refs))
    assert np.array_equal(ct, g["sharedk_ct"])
