cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dist_text|name_len|stats_rows' -c 60 --csv --log-file gpurun_out/r2_text_launches.csv python profiles/text_time.py > gpurun_out/r2_text68.log 2>&1
tail -4 gpurun_out/r2_text68.log
grep -c dist_text gpurun_out/r2_text_launches.csv
