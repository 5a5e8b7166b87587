set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_index_dist.py tests/test_gpu_chost.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
python profiles/dist_multi.py 8 0 2>gpurun_out/r2_dist_multi20_n1.err | tail -1 | tee gpurun_out/r2_dist_multi20_n1.log
tail -3 gpurun_out/r2_dist_multi20_n1.err
timeout 1200 python bench.py > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
tail -3 gpurun_out/r2_bench20.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench20.json").read().strip().splitlines()[-1])
print("value",j["value"],"ms/step",j["ms_per_step"],"scan",j["roofline"]["kernel_ms"],"frac",j["roofline"]["frac"])
print(json.dumps(j["e2e"],indent=1)[:2500])
print(json.dumps(j["fastq"],indent=1)[:1500])
print(json.dumps(j["cpu_baseline"],indent=1)[:1500])
c=j["dist"]["configs2_scale"]
for k,v in c.items():
    if k not in ("sharding","timing","content_check","oracle_check"): print(k, v)
PY
