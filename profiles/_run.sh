set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python profiles/op_bench.py --only conv --stage 0 > gpurun_out/r2r_opbench_stem.log 2>&1; tail -2 gpurun_out/r2r_opbench_stem.log
timeout 200 python profiles/trace_v2.py stem > gpurun_out/r2r_trace_stem.log 2>&1; sed -n 11,20p gpurun_out/r2r_trace_stem.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q -x -m gpu > gpurun_out/r2r_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_tests.log
tail -8 gpurun_out/r2r_tests.log
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
for f in gpurun_out/r2r_bench*.json; do echo $f; cut -c1-120 $f; done
