"""CPU: include/kssd_b200.h is plain C (compiles with gcc -std=c11 -pedantic) and a C host links against the
library the way INTEGRATION.md describes.  The program only calls kssd_version()/kssd_last_error() -- no GPU."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

C_SRC = r'''
#include <stdio.h>
#include <string.h>
#include "kssd_b200.h"
int main(void) {
    kssd_sketch_opts_t o; memset(&o, 0, sizeof o); o.mode = KSSD_MODE_FASTA;
    kssd_stat_opts_t so; memset(&so, 0, sizeof so); so.metric = KSSD_METRIC_JACCARD;
    kssd_stat_row_t row; (void)row; (void)o; (void)so;
    if (sizeof(kssd_stat_row_t) != 88) return 2;
    /* without a device the context must fail loudly, never fall back */
    kssd_ctx_t *ctx = NULL;
    static int shuf[1 << 20];
    int rc = kssd_ctx_create(&ctx, 0, shuf, 8, 5, 2, 7);
    printf("%s rc=%d err=%s\n", kssd_version(), rc, kssd_last_error());
    if (rc == 0) kssd_ctx_destroy(ctx);
    return 0;
}
'''


def test_header_is_c_and_links(tmp_path):
    from public_kssd_b200 import capi
    capi.build_library()
    src = tmp_path / "host.c"
    src.write_text(C_SRC)
    exe = tmp_path / "host"
    r = subprocess.run(["/usr/bin/gcc", "-std=c11", "-pedantic", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                        "-L", str(capi.PKG_DIR), "-lkssd_b200", f"-Wl,-rpath,{capi.PKG_DIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "sm_100a" in out.stdout


def test_c_host_program_builds(tmp_path):
    """host/kssd_b200_dist.c -- the Stage I / II / III driver in the reference's language -- compiles warning-free and
    links against the library; without arguments it prints its usage (no GPU touched)."""
    from public_kssd_b200 import capi
    capi.build_library()
    exe = tmp_path / "kssd_b200_dist"
    r = subprocess.run(["/usr/bin/gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"),
                        str(ROOT / "host" / "kssd_b200_dist.c"), "-o", str(exe), "-L", str(capi.PKG_DIR), "-lkssd_b200",
                        f"-Wl,-rpath,{capi.PKG_DIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: kssd_b200_dist sketch" in r.stderr
    import torch
    if not torch.cuda.is_available():
        # no device: the host must stop with the library's error, there is no CPU path to fall back to
        import struct
        import numpy as np
        shuf = tmp_path / "t.shuf"
        with open(shuf, "wb") as f:
            f.write(struct.pack("<iiii", 7, 8, 5, 2))
            f.write(np.arange(1 << 20, dtype="<i4").tobytes())
        fa = tmp_path / "a.fa"
        fa.write_text(">a\nACGTACGTACGTACGTACGTACGT\n")
        r = subprocess.run([str(exe), "sketch", str(shuf), str(tmp_path / "out"), str(fa)], capture_output=True, text=True)
        assert r.returncode == 1 and "kssd_ctx_create" in r.stderr
        assert not (tmp_path / "out" / "cofiles.stat").exists()
