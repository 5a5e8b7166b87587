cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_files.py tests/test_formatter.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest69.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest69.log
tail -4 gpurun_out/r2_pytest69.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
