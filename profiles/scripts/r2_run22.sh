set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_configs.py tests/test_gpu_composite.py tests/test_gpu_files.py tests/test_gpu_tutorial.py -m gpu -q --tb=short -x -p no:cacheprovider > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest22.log)
tail -15 gpurun_out/r2_pytest22.log
python profiles/fastq_scale.py 2>&1 | tail -3 | tee gpurun_out/r2_fastq22.log
KSSD_FASTQ_TWO_PASS=1 python profiles/fastq_scale.py 2>&1 | tail -3 | tee -a gpurun_out/r2_fastq22.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"nl_|sketch_fastq|DeviceScan" --csv --log-file gpurun_out/r2_fastq_launches22.csv python profiles/fastq_scale.py > /dev/null 2>&1
tail -8 gpurun_out/r2_fastq_launches22.csv | cut -c1-300
