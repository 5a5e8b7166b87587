set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" r2d "" r2d; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep -E "scan |rror" | tee gpurun_out/r2_ab36.log
timeout 600 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_configs.py tests/test_gpu_byread.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
