set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_files.py tests/test_gpu_chost.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest62.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest59.log)
tail -25 gpurun_out/r2_pytest62.log
timeout 400 python profiles/gz_time.py 1000 2>&1 | tail -8 | tee gpurun_out/r2_gz62.log
