set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log)
tail -15 gpurun_out/r2_pytest3.log
for v in "" 512 576 704 768; do
  if [ -z "$v" ]; then python profiles/ab_scan.py 1000; else KSSD_B200_LIB=$PWD/public_kssd_b200/variants/lib_t$v.so python profiles/ab_scan.py 1000; fi
done 2>&1 | grep "scan " | tee gpurun_out/r2_ab3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sketch_fasta3_kernel -c 1 -o gpurun_out/r2_sketch_v11 python bench.py --genomes 200 --steps 1 --warmup 0 --no-cpu-baseline --no-dist-scale > gpurun_out/r2_ncu_v11.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
head -c 3000 gpurun_out/r2_bench3.json
python profiles/dist_scale.py > gpurun_out/r2_dist_scale3.log 2>&1; tail -20 gpurun_out/r2_dist_scale3.log
