set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
timeout 120 python profiles/trace_v2.py attn > gpurun_out/r2d_trace_attn.log 2>&1
timeout 120 python profiles/trace_v2.py mlp > gpurun_out/r2d_trace_mlp.log 2>&1
timeout 200 python profiles/op_bench.py > gpurun_out/r2d_opbench.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
# PDL
RVT_PDL=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py tests/test_gpu_train.py -q -x > gpurun_out/r2d_tests_pdl.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_tests_pdl.log
RVT_PDL=1 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline > gpurun_out/r2d_bench_pdl.json 2> gpurun_out/r2d_bench_pdl.err
RVT_PDL=1 timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline --no-wavefront > gpurun_out/r2d_bench_pdl_nowf.json 2> gpurun_out/r2d_bench_pdl_nowf.err
timeout 400 python bench.py --steps 10 --warmup 3 --extras "" --no-cpu-baseline --no-wavefront > gpurun_out/r2d_bench_nowf.json 2> gpurun_out/r2d_bench_nowf.err
tail -5 gpurun_out/r2d_tests.log; tail -5 gpurun_out/r2d_tests_pdl.log; head -14 gpurun_out/r2d_trace_attn.log; head -9 gpurun_out/r2d_trace_mlp.log; cat gpurun_out/r2d_opbench.log
cut -c1-200 gpurun_out/r2d_bench.json gpurun_out/r2d_bench_pdl.json gpurun_out/r2d_bench_pdl_nowf.json gpurun_out/r2d_bench_nowf.json
tail -5 gpurun_out/r2d_bench_pdl.err
