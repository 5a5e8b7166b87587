set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" tma; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep -E "scan |Error|error" | tee gpurun_out/r2_ab16.log
KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_tma.so timeout 600 python -m pytest tests/test_gpu_sketch.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 | tee -a gpurun_out/r2_ab16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3 -s 2 -c 1 -o gpurun_out/r2_sketch_v16 python profiles/ab_scan.py 200 > gpurun_out/r2_ncu_v16.log 2>&1
KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_tma.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3 -s 2 -c 1 -o gpurun_out/r2_sketch_v16tma python profiles/ab_scan.py 200 > gpurun_out/r2_ncu_v16tma.log 2>&1
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest16.log)
tail -6 gpurun_out/r2_pytest16.log
