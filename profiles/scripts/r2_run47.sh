set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_sparse -s 1 -c 1 -o gpurun_out/r2_dist_sparse_n1 python profiles/dist_multi.py 2 0 > gpurun_out/r2_ncu_dist_sparse_n1.log 2>&1
tail -2 gpurun_out/r2_ncu_dist_sparse_n1.log | cut -c1-200
