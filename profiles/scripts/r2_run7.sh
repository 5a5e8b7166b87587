set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log)
tail -5 gpurun_out/r2_pytest7.log
for v in "" t640; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep "scan " | tee gpurun_out/r2_ab7.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench7.json",):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value",j["value"],"ms/step",j["ms_per_step"],"scan",j["roofline"]["kernel_ms"],"frac",j["roofline"]["frac"],"e2e",j["e2e"]["value"], j["e2e"]["matches_device_path"], "launches", j["gpu_launches"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3_kernel -c 1 -o gpurun_out/r2_sketch_v15 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v15.log 2>&1
