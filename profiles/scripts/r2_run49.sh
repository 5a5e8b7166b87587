set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 profiles/dist_multi.py 8 0 2> gpurun_out/r2_dist_multi49_n8.err | tail -1 | tee gpurun_out/r2_dist_multi49_n8.log | cut -c1-900
tail -2 gpurun_out/r2_dist_multi49_n8.err
