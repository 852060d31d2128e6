cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -x -m gpu > gpurun_out/final2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final2_tests.log
timeout 300 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?"; head -c 200 gpurun_out/final2_bench.json; echo; grep -v Warning gpurun_out/final2_bench.err | tail -3
