set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py tests/test_gpu_files.py tests/test_gpu_tutorial.py tests/test_gpu_set.py tests/test_gpu_index_dist.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log)
tail -15 gpurun_out/r2_pytest4.log
for v in "" 512 576 768; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_t$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep "scan " | tee gpurun_out/r2_ab4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3_kernel -c 1 -o gpurun_out/r2_sketch_v12 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v12.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_launches4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_launches4.log 2>&1
