set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
# ---- attn_v2 / mlp_v2 parity ----
RVT_ATTN_V2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -x -q > gpurun_out/r2a_tests_v2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests_v2.log
mkdir -p gpurun_out/v2 && cp gpurun_out/op_parity_*.json gpurun_out/v2/ 2>/dev/null
RVT_MLP_V2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -x -q > gpurun_out/r2a_tests_mlpv2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests_mlpv2.log
mkdir -p gpurun_out/mlpv2 && cp gpurun_out/op_parity_*.json gpurun_out/mlpv2/ 2>/dev/null
# ---- op timings ----
RVT_ATTN_V2=0 RVT_MLP_V2=0 timeout 200 python profiles/op_bench.py > gpurun_out/r2a_opbench_v1.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 200 python profiles/op_bench.py > gpurun_out/r2a_opbench_v2.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_GELU_F16X2=1 timeout 200 python profiles/op_bench.py --only mlp > gpurun_out/r2a_opbench_v2h.log 2>&1
# ---- bench ----
timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2a_bench_v1.json 2> gpurun_out/r2a_bench_v1.err
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_v2.json 2> gpurun_out/r2a_bench_v2.err
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_GELU_F16X2=1 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2a_bench_v2h.json 2> gpurun_out/r2a_bench_v2h.err
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_PERSIST_SMS=132 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2a_bench_v2_p132.json 2> gpurun_out/r2a_bench_v2_p132.err
# ---- envelopes ----
timeout 600 python profiles/amp_envelope.py --batch 2 --steps 21 > gpurun_out/r2a_amp.log 2>&1
RVT_GELU_F16X2=1 RVT_MLP_V2=1 timeout 600 python -m pytest tests/test_gpu_ops.py -q > gpurun_out/r2a_tests_f16x2.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests_f16x2.log
mkdir -p gpurun_out/f16x2 && cp gpurun_out/op_parity_*.json gpurun_out/f16x2/ 2>/dev/null
tail -5 gpurun_out/r2a_tests.log; tail -8 gpurun_out/r2a_tests_v2.log; tail -8 gpurun_out/r2a_tests_mlpv2.log
cat gpurun_out/r2a_opbench_v1.log gpurun_out/r2a_opbench_v2.log gpurun_out/r2a_opbench_v2h.log
cut -c1-250 gpurun_out/r2a_bench_v1.json gpurun_out/r2a_bench_v2.json gpurun_out/r2a_bench_v2h.json gpurun_out/r2a_bench_v2_p132.json
tail -8 gpurun_out/r2a_amp.log; tail -3 gpurun_out/r2a_tests_f16x2.log
