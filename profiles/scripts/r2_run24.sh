set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq|fastq2co|^-A" | tail -12 | tee gpurun_out/r2_fastq24.log
