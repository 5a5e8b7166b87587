set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python profiles/gz_time.py 3600 1800,3600 2>&1 | tail -4 | tee gpurun_out/r2_gz64.log
