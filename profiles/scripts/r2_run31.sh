set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest31.log)
tail -6 gpurun_out/r2_pytest31.log
timeout 1200 python bench.py > gpurun_out/r2_bench31.json 2> gpurun_out/r2_bench31.err
tail -3 gpurun_out/r2_bench31.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3 -s 2 -c 1 -o gpurun_out/r2_sketch_v17 python profiles/ab_scan.py 200 > gpurun_out/r2_ncu_v17.log 2>&1
tail -2 gpurun_out/r2_ncu_v17.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches31.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-files > gpurun_out/r2_bench_under_ncu31.log 2>&1
tail -2 gpurun_out/r2_bench_launches31.csv | cut -c1-200
