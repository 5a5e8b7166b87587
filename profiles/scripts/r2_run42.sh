set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq|fastq2co|^-A|parity" | tail -8 | tee gpurun_out/r2_fastq42.log
timeout 600 python -m pytest tests/test_gpu_fastq.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
