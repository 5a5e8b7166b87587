set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_index_dist.py tests/test_gpu_chost.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest55.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest55.log)
tail -15 gpurun_out/r2_pytest55.log
timeout 300 python profiles/topn_time.py 2>&1 | tail -6 | tee gpurun_out/r2_topn55.log
