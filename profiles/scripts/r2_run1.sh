set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log)
tail -30 gpurun_out/r2_pytest1.log
(timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_bench_new.json 2> gpurun_out/r2_bench_new.err; echo rc=$?)
(KSSD_SCAN_IMPL=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_bench_old.json 2> gpurun_out/r2_bench_old.err; echo rc=$?)
cat gpurun_out/r2_bench_new.json | head -c 1500; echo
cat gpurun_out/r2_bench_old.json | head -c 600; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3 -c 1 -o gpurun_out/r2_sketch_v9 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v9.log 2>&1
ls -la gpurun_out | tail -8
