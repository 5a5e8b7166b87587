set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py tests/test_gpu_files.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log)
tail -8 gpurun_out/r2_pytest5.log
for v in "" pf2 t576 t672 t768; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep "scan " | tee gpurun_out/r2_ab5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3_kernel -c 1 -o gpurun_out/r2_sketch_v13 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v13.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_launches5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_launches5.log 2>&1
