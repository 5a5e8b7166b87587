set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 540 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_index_dist.py -q -m gpu -p no:cacheprovider -k "sparse" > gpurun_out/r2_racecheck2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_racecheck2.log)
grep -E "passed|failed|RACECHECK SUMMARY|rc=" gpurun_out/r2_racecheck2.log | tail -4
