set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --no-fastq --no-files --no-cpu-baseline > gpurun_out/r2_bench52.json 2> gpurun_out/r2_bench52.err
tail -3 gpurun_out/r2_bench52.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench52.json").read().strip().splitlines()[-1])
print(j["dist"]["configs2_scale"]["e2e"], j["dist"]["configs2_scale"]["ms_per_batch"])
PY
