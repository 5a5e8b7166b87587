set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_composite.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6)
KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq|fastq2co|^-A|parity" | tail -8 | tee gpurun_out/r2_fastq45.log
