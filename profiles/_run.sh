set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
RVT_STEM_V2=1 timeout 200 python profiles/trace_v2.py stem > gpurun_out/r2j_trace_stem.log 2>&1; head -40 gpurun_out/r2j_trace_stem.log
export RVT_CONV_TMA=1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q -x -m gpu > gpurun_out/r2j_tests_ctma.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests_ctma.log
tail -12 gpurun_out/r2j_tests_ctma.log
timeout 200 python profiles/op_bench.py --only conv > gpurun_out/r2j_opbench_ctma.log 2>&1; cat gpurun_out/r2j_opbench_ctma.log | tail -6
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2j_bench_ctma.json 2> gpurun_out/r2j_bench_ctma.err
for f in gpurun_out/r2j_bench*.json; do echo $f; cut -c1-120 $f; done
