set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python profiles/gz_time.py 3600 2>&1 | tail -9 | tee gpurun_out/r2_gz63.log
