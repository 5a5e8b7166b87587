set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_index_dist.py tests/test_gpu_configs.py tests/test_gpu_composite.py tests/test_gpu_chost.py tests/test_gpu_interop.py tests/test_gpu_tutorial.py -m gpu -q --tb=short -x -p no:cacheprovider 2>&1 | tail -4)
python profiles/index_scale.py 2>&1 | tail -1 | tee gpurun_out/r2_index_scale38.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tag_gids|rb_build|Device|Radix|Scan|RunLength|RLE|Reduce" --csv --log-file gpurun_out/r2_index_launches38.csv python profiles/index_scale.py > /dev/null 2>&1
tail -16 gpurun_out/r2_index_launches38.csv | awk -F'","' '{print $5, $NF}' | cut -c1-140
