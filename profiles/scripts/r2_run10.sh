set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log)
tail -12 gpurun_out/r2_pytest10.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err
tail -3 gpurun_out/r2_bench10.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench10.json").read().strip().splitlines()[-1])
print("value",j["value"],"ms/step",j["ms_per_step"],"scan",j["roofline"]["kernel_ms"],"frac",j["roofline"]["frac"])
c=j["dist"]["configs2_scale"]
for k,v in c.items():
    if k not in ("roofline","e2e","sharding","timing","content_check","oracle_check"): print(k, v)
PY
KSSD_SPARSE_SHAPE=wide python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, bench_dist
from public_kssd_b200 import kssd, synth
ctx = kssd.Context(10, 6, 3, synth.make_shuf_table(6, 1))
out = bench_dist.run(ctx, 1, 0, torch.device("cuda",0), 6545.3, batches=4)
print("WIDE ms_per_batch", out["ms_per_batch"], out["rank0_kernel_ms_per_batch_untimed_pass"], out["content_ok"])
PY
