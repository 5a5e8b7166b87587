set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_index_dist.py tests/test_gpu_configs.py tests/test_gpu_chost.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4)
python profiles/dist_multi.py 8 0 2>gpurun_out/r2_dist_multi48_n1.err | tail -1 | tee gpurun_out/r2_dist_multi48_n1.log | cut -c1-700
