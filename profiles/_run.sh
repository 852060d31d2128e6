set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r2s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_tests.log
tail -4 gpurun_out/r2s_tests.log
timeout 200 python profiles/op_bench.py > gpurun_out/r2s_opbench.log 2>&1; tail -18 gpurun_out/r2s_opbench.log
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
RVT_TMA_LONE_KB=0 timeout 400 $B > gpurun_out/r2s_bench_nolone.json 2> gpurun_out/r2s_bench_nolone.err
for f in gpurun_out/r2s_bench*.json; do echo $f; cut -c1-120 $f; done
