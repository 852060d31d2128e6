cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/final_bench.json; echo; grep -v Warning gpurun_out/final_bench.err | tail -30
