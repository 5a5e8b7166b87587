set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py tests/test_gpu_fastq.py tests/test_gpu_files.py tests/test_gpu_interop.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log)
tail -8 gpurun_out/r2_pytest23.log
for v in "" r2a ""  r2a; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep -E "scan |rror" | tee gpurun_out/r2_ab23.log
python profiles/fastq_scale.py 2>&1 | tail -3 | tee gpurun_out/r2_fastq23.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"nl_|sketch_fastq" --csv --log-file gpurun_out/r2_fastq_launches23.csv python profiles/fastq_scale.py > /dev/null 2>&1
grep -E "nl_index|sketch_fastq" gpurun_out/r2_fastq_launches23.csv | grep duration | tail -4 | cut -c1-60,200-400
