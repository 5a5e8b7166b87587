set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log)
tail -15 gpurun_out/r2_pytest6.log
python profiles/ab_scan.py 1000 2>&1 | grep "scan " | tee gpurun_out/r2_ab6.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
KSSD_NO_BUCKETS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_bench6_list.json 2> gpurun_out/r2_bench6_list.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench6.json","gpurun_out/r2_bench6_list.json"):
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value",j["value"],"ms/step",j["ms_per_step"],"scan",j["roofline"]["kernel_ms"],"frac",j["roofline"]["frac"],"e2e",j["e2e"]["value"], j["e2e"]["matches_device_path"], "launches", j["gpu_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kssd|cub|Device" -c 120 --csv --log-file gpurun_out/r2_launches6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_launches6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3_kernel -c 1 -o gpurun_out/r2_sketch_v14 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v14.log 2>&1
