set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench9_n2.json 2> gpurun_out/r2_bench9_n2.err
tail -5 gpurun_out/r2_bench9_n2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench9_n2.json").read().strip().splitlines()[-1])
print("value",j["value"],"ms/step",j["ms_per_step"],"e2e",j["e2e"]["value"])
print(json.dumps(j["dist"],indent=1)[:3000])
PY
