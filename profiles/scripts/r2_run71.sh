cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 85 python -m pytest tests -m gpu -q -x --tb=line -p no:cacheprovider > gpurun_out/r2_pytest71.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest71.log
tail -4 gpurun_out/r2_pytest71.log
