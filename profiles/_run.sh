set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
timeout 120 python profiles/trace_v2.py attn > gpurun_out/r2e_trace_attn.log 2>&1
timeout 120 python profiles/trace_v2.py mlp > gpurun_out/r2e_trace_mlp.log 2>&1
timeout 200 python profiles/op_bench.py > gpurun_out/r2e_opbench.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
# C = 128 MLP
RVT_MLP_V2=2 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q -x > gpurun_out/r2e_tests_mlpx.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tests_mlpx.log
RVT_MLP_V2=2 timeout 200 python profiles/op_bench.py --only mlp > gpurun_out/r2e_opbench_mlpx.log 2>&1
RVT_MLP_V2=2 timeout 120 python profiles/trace_v2.py mlp --stage 1 > gpurun_out/r2e_trace_mlpx.log 2>&1
RVT_MLP_V2=2 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2e_bench_mlpx.json 2> gpurun_out/r2e_bench_mlpx.err
tail -5 gpurun_out/r2e_tests.log; tail -5 gpurun_out/r2e_tests_mlpx.log; head -14 gpurun_out/r2e_trace_attn.log; head -9 gpurun_out/r2e_trace_mlp.log; head -12 gpurun_out/r2e_trace_mlpx.log; grep -E "S1|S2" gpurun_out/r2e_opbench.log; cat gpurun_out/r2e_opbench_mlpx.log
cut -c1-200 gpurun_out/r2e_bench.json gpurun_out/r2e_bench_mlpx.json
