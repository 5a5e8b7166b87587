set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for s in narrow wide; do KSSD_SPARSE_SHAPE=$s python profiles/dist_shard.py 8 2>&1 | tail -2; done | tee gpurun_out/r2_dist_shard.log
KSSD_SPARSE_SHAPE=narrow timeout 300 ncu --set full --clock-control none --import-source on -k regex:dist_sparse_kernel -s 2 -c 1 -o gpurun_out/r2_dist_sparse_shard python profiles/dist_shard.py 8 > gpurun_out/r2_ncu_shard.log 2>&1
python profiles/index_scale.py 2>&1 | tail -1 | tee gpurun_out/r2_index_scale.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kssd|cub|Device" --csv --log-file gpurun_out/r2_index_launches.csv python profiles/index_scale.py > /dev/null 2>&1
for v in "" tma; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep "scan " | tee gpurun_out/r2_ab13.log
KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_tma.so timeout 600 python -m pytest tests/test_gpu_sketch.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2 | tee -a gpurun_out/r2_ab13.log
