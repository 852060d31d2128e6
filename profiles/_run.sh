cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/final2_launches.csv python bench.py --steps 1 --warmup 1 --extras '' --no-cpu-baseline > gpurun_out/final2_launches_bench.log 2>&1; echo "launch list rc=$?"
