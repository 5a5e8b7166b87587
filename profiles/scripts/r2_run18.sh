set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 profiles/dist_multi.py 8 0 2> gpurun_out/r2_dist_multi18_n2.err | tail -1 | tee gpurun_out/r2_dist_multi18_n2.log
tail -3 gpurun_out/r2_dist_multi18_n2.err
bash profiles/tma_memcheck.sh
