set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench39_n8.json 2> gpurun_out/r2_bench39_n8.err
tail -3 gpurun_out/r2_bench39_n8.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench39_n8.json").read().strip().splitlines()[-1])
print("value",j["value"],"ms/step",j["ms_per_step"],"frac",j["roofline"]["frac"])
print(json.dumps(j["e2e"])[:600])
c=j["dist"]["configs2_scale"]
for k,v in c.items():
    if k not in ("sharding","timing","content_check","oracle_check","baseline_note"): print(k, v)
PY
