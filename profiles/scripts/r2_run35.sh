set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "" pf4 pf6 ""; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep -E "scan |rror" | tee gpurun_out/r2_ab35.log
