set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
timeout 120 python profiles/trace_v2.py attn > gpurun_out/r2c_trace_attn.log 2>&1
timeout 120 python profiles/trace_v2.py mlp > gpurun_out/r2c_trace_mlp.log 2>&1
timeout 120 python profiles/trace_v2.py attn --stage 1 > gpurun_out/r2c_trace_attn_s2.log 2>&1
timeout 200 python profiles/op_bench.py > gpurun_out/r2c_opbench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
RVT_PERSIST_SMS=132 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2c_bench_p132.json 2> gpurun_out/r2c_bench_p132.err
RVT_PERSIST_SMS=140 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2c_bench_p140.json 2> gpurun_out/r2c_bench_p140.err
tail -5 gpurun_out/r2c_tests.log; head -14 gpurun_out/r2c_trace_attn.log; head -9 gpurun_out/r2c_trace_mlp.log; cat gpurun_out/r2c_opbench.log
cut -c1-200 gpurun_out/r2c_bench.json gpurun_out/r2c_bench_p132.json gpurun_out/r2c_bench_p140.json
