set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for shape in wide narrow; do
KSSD_SPARSE_SHAPE=$shape timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 profiles/dist_multi.py 8 0 2>/dev/null | tail -1 | tee -a gpurun_out/r2_dist_multi_n2.log
done
