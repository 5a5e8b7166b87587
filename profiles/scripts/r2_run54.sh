set -x
cd $GRAFT_REPO_ROOT
KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq|fastq2co|^-A|parity" | tail -8
timeout 600 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_configs.py tests/test_gpu_composite.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
