set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest66.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest66.log)
tail -6 gpurun_out/r2_pytest66.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench66.json 2> gpurun_out/r2_bench66.err; tail -c 3000 gpurun_out/r2_bench66.json
