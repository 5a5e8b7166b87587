set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
KSSD_FASTQ_TIMING=1 KSSD_NLX_NOLB=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq" | head -6 | tee gpurun_out/r2_fastq26.log
