set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py tests/test_gpu_files.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log)
tail -5 gpurun_out/r2_pytest8.log
for v in "" t640; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep "scan " | tee gpurun_out/r2_ab8.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
tail -5 gpurun_out/r2_bench8.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_bench8.json").read().strip().splitlines()[-1])
print("value",j["value"],"ms/step",j["ms_per_step"],"scan",j["roofline"]["kernel_ms"],"frac",j["roofline"]["frac"],"e2e",j["e2e"]["value"], j["e2e"]["matches_device_path"])
print(json.dumps(j["dist"].get("configs2_scale"),indent=1)[:3500])
print(json.dumps(j["cpu_baseline"],indent=1)[:1500])
PY
