set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python profiles/trace_v2.py stem > gpurun_out/r2q_trace_stem.log 2>&1; sed -n 10,20p gpurun_out/r2q_trace_stem.log
export RVT_FAST_GATES=1
timeout 200 python profiles/op_bench.py --only lstm > gpurun_out/r2q_opbench_fg.log 2>&1; tail -5 gpurun_out/r2q_opbench_fg.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py tests/test_gpu_parity_envelope.py -q -x -m gpu > gpurun_out/r2q_tests_fg.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_tests_fg.log
tail -12 gpurun_out/r2q_tests_fg.log
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2q_bench_fg.json 2> gpurun_out/r2q_bench_fg.err
for f in gpurun_out/r2q_bench*.json; do echo $f; cut -c1-120 $f; done
