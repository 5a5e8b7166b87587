set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_fastq.py tests/test_gpu_byread.py tests/test_gpu_set.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_memcheck1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_memcheck1.log)
grep -E "passed|failed|ERROR SUMMARY|rc=" gpurun_out/r2_memcheck1.log | tail -4
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_index_dist.py tests/test_gpu_files.py tests/test_gpu_composite.py -x -q -m gpu -p no:cacheprovider > gpurun_out/r2_memcheck2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_memcheck2.log)
grep -E "passed|failed|ERROR SUMMARY|rc=" gpurun_out/r2_memcheck2.log | tail -4
(timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_fastq.py -q -m gpu -p no:cacheprovider -k "test_fasta_uniq_parity or test_header_eof or mixed_read_lengths or line_index_paths" > gpurun_out/r2_racecheck1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_racecheck1.log)
grep -E "passed|failed|RACECHECK SUMMARY|rc=" gpurun_out/r2_racecheck1.log | tail -4
