set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py tests/test_gpu_files.py tests/test_gpu_tutorial.py tests/test_gpu_interop.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log)
tail -15 gpurun_out/r2_pytest2.log
for v in "" public_kssd_b200/variants/lib_t640.so public_kssd_b200/variants/lib_t704.so; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/$v python profiles/ab_scan.py 1000; fi
done 2>&1 | grep -v Warning | tee gpurun_out/r2_ab2.log
KSSD_SCAN_IMPL=2 python profiles/ab_scan.py 1000 2>&1 | tail -1 | tee -a gpurun_out/r2_ab2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3_kernel -c 1 -o gpurun_out/r2_sketch_v10 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v10.log 2>&1
ls -la gpurun_out | tail -5
