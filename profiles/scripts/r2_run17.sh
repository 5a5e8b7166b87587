set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python profiles/dist_multi.py 8 0 2>gpurun_out/r2_dist_multi17_n1.err | tail -1 | tee gpurun_out/r2_dist_multi17_n1.log
tail -3 gpurun_out/r2_dist_multi17_n1.err
timeout 300 python -m pytest tests/test_gpu_index_dist.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_tma.so timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python profiles/ab_scan.py 3 > gpurun_out/r2_tma_memcheck.log 2>&1
grep -v "^$" gpurun_out/r2_tma_memcheck.log | head -40
