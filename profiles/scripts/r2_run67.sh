cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_files.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest67.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest67.log
tail -5 gpurun_out/r2_pytest67.log
