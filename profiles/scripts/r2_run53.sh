set -x
cd $GRAFT_REPO_ROOT
for m in 0 2 1; do KSSD_NLX_NOLB=$m KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2>&1 | grep -E "kssd fastq" | sed -n 2,3p; done
