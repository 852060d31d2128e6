set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export RVT_STEM_V2=1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q -x -m gpu > gpurun_out/r2h_tests_stem.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests_stem.log
tail -5 gpurun_out/r2h_tests_stem.log
timeout 200 python profiles/op_bench.py --only conv --stage 0 > gpurun_out/r2h_opbench_stem.log 2>&1; cat gpurun_out/r2h_opbench_stem.log | tail -3
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 400 $B > gpurun_out/r2h_bench_stem.json 2> gpurun_out/r2h_bench_stem.err
timeout 300 python profiles/wavefront_timeline.py > gpurun_out/r2h_timeline.log 2>&1
timeout 300 python profiles/wavefront_timeline.py --no-wavefront > gpurun_out/r2h_timeline_seq.log 2>&1
for f in gpurun_out/r2h_bench*.json; do echo $f; cut -c1-120 $f; done
head -8 gpurun_out/r2h_timeline.log; head -8 gpurun_out/r2h_timeline_seq.log
