set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_files.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3)
timeout 400 python profiles/gz_time.py 3600 1000,3600 2>&1 | tail -4 | tee gpurun_out/r2_gz65.log
