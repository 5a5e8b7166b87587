set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for shape in narrow wide; do
KSSD_SPARSE_SHAPE=$shape timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 profiles/dist_multi.py 8 0 2>/dev/null | tail -1 | tee -a gpurun_out/r2_dist_multi_n8.log
done
