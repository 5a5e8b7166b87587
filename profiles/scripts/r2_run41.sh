set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fastq3 -s 1 -c 1 -o gpurun_out/r2_fastq3_v2 python profiles/fastq_scale.py > gpurun_out/r2_ncu_fastq3_v1.log 2>&1
tail -2 gpurun_out/r2_ncu_fastq3_v1.log | cut -c1-200
