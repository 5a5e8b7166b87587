set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest50.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest50.log)
tail -4 gpurun_out/r2_pytest50.log
timeout 1200 python bench.py > gpurun_out/r2_bench50.json 2> gpurun_out/r2_bench50.err
tail -3 gpurun_out/r2_bench50.err
