# memcheck of the TMA A/B variant through the C host (no python under the sanitizer): one 24 Mbp genome, spans of 256 KiB
set -x
cd $GRAFT_REPO_ROOT
D=/tmp/tma_chk; mkdir -p $D
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from public_kssd_b200 import synth, kssd
tab = synth.make_shuf_table(6, 1)
kssd.write_shuf_file('/tmp/tma_chk/t.shuf', 1, 10, 6, 3, tab)
g = synth.to_fasta(synth.random_bases(24_000_000, 3), 'g0', 80)
open('/tmp/tma_chk/g0.fa', 'wb').write(g.tobytes())
PY
ls -la $D
LD_LIBRARY_PATH=$PWD/public_kssd_b200/variants/tma timeout 200 compute-sanitizer --tool memcheck --print-limit 8 host/kssd_b200_dist sketch $D/t.shuf $D/out $D/g0.fa > gpurun_out/r2_tma_memcheck.log 2>&1
grep -v "^$" gpurun_out/r2_tma_memcheck.log | head -60
