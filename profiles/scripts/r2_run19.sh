set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 profiles/dist_multi.py 8 1 2> gpurun_out/r2_dist_multi21_n8.err | tail -1 | tee gpurun_out/r2_dist_multi21_n8.log
tail -3 gpurun_out/r2_dist_multi21_n8.err
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
