set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 profiles/dist_multi.py 8 0 2> gpurun_out/r2_dist_multi_n4.err | tail -1 | tee -a gpurun_out/r2_dist_multi_n4.log
tail -5 gpurun_out/r2_dist_multi_n4.err
