set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.build(); g.smoke()' 2>&1 | tail -2
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest44.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest44.log)
tail -5 gpurun_out/r2_pytest44.log
timeout 1200 python bench.py > gpurun_out/r2_bench44.json 2> gpurun_out/r2_bench44.err
tail -3 gpurun_out/r2_bench44.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench44_ref.json 2> gpurun_out/r2_bench44_ref.err
tail -2 gpurun_out/r2_bench44_ref.err; cut -c1-400 gpurun_out/r2_bench44_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sketch_|bucket_|span_|plan|nl_|dist_|stats_|tag_gids|rb_build|group_|set_|Device|Radix|cub" -c 400 --csv --log-file gpurun_out/r2_bench_launches44.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-files --no-fastq --no-dist-scale > gpurun_out/r2_bench_under_ncu44.log 2>&1
tail -2 gpurun_out/r2_bench_launches44.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:dist_count_rows -c 1 -o gpurun_out/r2_dist_count_rows python profiles/dist_multi.py 2 0 > gpurun_out/r2_ncu_dist_rows.log 2>&1
tail -2 gpurun_out/r2_ncu_dist_rows.log | cut -c1-200
