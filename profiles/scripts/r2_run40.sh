set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_configs.py tests/test_gpu_composite.py tests/test_gpu_files.py tests/test_gpu_sketch.py tests/test_gpu_byread.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest40.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest40.log)
tail -25 gpurun_out/r2_pytest40.log | cut -c1-200
KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq|fastq2co|^-A|parity" | tail -8 | tee gpurun_out/r2_fastq40.log
KSSD_FASTQ_THREAD_WALK=1 KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "fastq2co|^-A|parity" | tail -3 | tee -a gpurun_out/r2_fastq40.log
