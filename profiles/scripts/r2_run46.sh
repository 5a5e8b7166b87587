set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" nlx5 nlx6; do
  if [ -z "$v" ]; then KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2000000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so KSSD_FASTQ_TIMING=1 python profiles/fastq_scale.py 2000000; fi 2>&1 | grep -E "kssd fastq" | sed -n 2,3p
done | tee gpurun_out/r2_fastq46.log
