set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r2i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests.log
tail -4 gpurun_out/r2i_tests.log
timeout 200 python profiles/op_bench.py > gpurun_out/r2i_opbench.log 2>&1; cat gpurun_out/r2i_opbench.log | tail -18
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
export RVT_STEM_V2=1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q -x -m gpu > gpurun_out/r2i_tests_stem.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests_stem.log
tail -3 gpurun_out/r2i_tests_stem.log
timeout 200 python profiles/op_bench.py --only conv --stage 0 > gpurun_out/r2i_opbench_stem.log 2>&1; cat gpurun_out/r2i_opbench_stem.log | tail -3
timeout 400 $B > gpurun_out/r2i_bench_stem.json 2> gpurun_out/r2i_bench_stem.err
for f in gpurun_out/r2i_bench*.json; do echo $f; cut -c1-120 $f; done
