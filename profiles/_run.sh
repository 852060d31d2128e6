set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python profiles/amp_envelope.py --batch 2 --steps 21 > gpurun_out/r2a_amp.log 2>&1
RVT_GELU_F16X2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q > gpurun_out/r2a_tests_f16x2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests_f16x2.log
mkdir -p gpurun_out/f16x2 && cp gpurun_out/op_parity_*.json gpurun_out/f16x2/ 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
RVT_GELU_F16X2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_f16x2.json 2> gpurun_out/r2a_bench_f16x2.err
# ---- attn_v2 ----
RVT_ATTN_V2=1 timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -x -q > gpurun_out/r2a_tests_v2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests_v2.log
mkdir -p gpurun_out/v2 && cp gpurun_out/op_parity_*.json gpurun_out/v2/ 2>/dev/null
RVT_ATTN_V2=0 timeout 200 python profiles/op_bench.py --only attn > gpurun_out/r2a_opbench_attn_v1.log 2>&1
RVT_ATTN_V2=1 timeout 200 python profiles/op_bench.py --only attn > gpurun_out/r2a_opbench_attn_v2.log 2>&1
RVT_ATTN_V2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_v2.json 2> gpurun_out/r2a_bench_v2.err
# ---- mlp_v2 ----
RVT_MLP_V2=1 RVT_ATTN_V2=1 timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -x -q > gpurun_out/r2a_tests_mlpv2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests_mlpv2.log
mkdir -p gpurun_out/mlpv2 && cp gpurun_out/op_parity_*.json gpurun_out/mlpv2/ 2>/dev/null
RVT_MLP_V2=0 timeout 200 python profiles/op_bench.py --only mlp > gpurun_out/r2a_opbench_mlp_v1.log 2>&1
RVT_MLP_V2=1 timeout 200 python profiles/op_bench.py --only mlp > gpurun_out/r2a_opbench_mlp_v2.log 2>&1
RVT_MLP_V2=1 RVT_GELU_F16X2=1 timeout 200 python profiles/op_bench.py --only mlp > gpurun_out/r2a_opbench_mlp_v2h.log 2>&1
RVT_MLP_V2=1 RVT_ATTN_V2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_v2both.json 2> gpurun_out/r2a_bench_v2both.err
RVT_MLP_V2=1 RVT_ATTN_V2=1 RVT_GELU_F16X2=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_v2both_h2.json 2> gpurun_out/r2a_bench_v2both_h2.err
tail -3 gpurun_out/r2a_tests.log; tail -12 gpurun_out/r2a_amp.log; tail -3 gpurun_out/r2a_tests_f16x2.log; cat gpurun_out/r2a_bench.json gpurun_out/r2a_bench_f16x2.json | cut -c1-300
tail -15 gpurun_out/r2a_tests_v2.log; cat gpurun_out/r2a_opbench_attn_v1.log gpurun_out/r2a_opbench_attn_v2.log; cut -c1-300 gpurun_out/r2a_bench_v2.json
tail -15 gpurun_out/r2a_tests_mlpv2.log; cat gpurun_out/r2a_opbench_mlp_v1.log gpurun_out/r2a_opbench_mlp_v2.log gpurun_out/r2a_opbench_mlp_v2h.log; cut -c1-300 gpurun_out/r2a_bench_v2both.json gpurun_out/r2a_bench_v2both_h2.json
