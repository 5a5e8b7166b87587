set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_formatter.py tests/test_gpu_index_dist.py tests/test_gpu_fastq.py -m gpu -q --tb=short -p no:cacheprovider -k "text or sparse_job_rows or golden" > gpurun_out/r2_pytest58.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest58.log)
tail -25 gpurun_out/r2_pytest58.log
timeout 300 python profiles/text_time.py 2>&1 | tail -6 | tee gpurun_out/r2_text58.log
