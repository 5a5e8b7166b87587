set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py tests/test_gpu_files.py tests/test_gpu_interop.py tests/test_gpu_tutorial.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest28.log)
tail -6 gpurun_out/r2_pytest28.log
for v in "" r2a ""  r2a; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep -E "scan |rror" | tee gpurun_out/r2_ab28.log
KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq|fastq2co|^-A|parity" | tail -7 | tee gpurun_out/r2_fastq28.log
