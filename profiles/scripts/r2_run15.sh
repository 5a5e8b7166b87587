set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_index_dist.py tests/test_gpu_configs.py tests/test_gpu_composite.py tests/test_gpu_chost.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log)
tail -8 gpurun_out/r2_pytest15.log
python profiles/dist_multi.py 8 0 2>/dev/null | tail -1 | tee gpurun_out/r2_dist_multi_n1.log
python profiles/ab_scan.py 1000 2>&1 | tail -3
KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_tma.so python profiles/ab_scan.py 1000 2>&1 | tail -5
