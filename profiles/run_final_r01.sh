#!/bin/bash
# final round-1 evidence: every command has its own tight timeout (a hung kernel must not eat the GPU budget)
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 100 -s > gpurun_out/final_tests.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/final_tests.log
timeout 110 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; head -c 400 gpurun_out/final_bench.json; echo
timeout 70 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/final_bench_train.json 2> gpurun_out/final_bench_train.err; echo "train exit $?"; head -c 300 gpurun_out/final_bench_train.json; echo
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/final_smoke.log
RVT_ATTN_BWD=1 timeout 45 python -m pytest tests/test_gpu_train_ops.py -k attn_core_bwd -q -m gpu -p no:cacheprovider --timeout 30 > gpurun_out/final_attn_tc1.log 2>&1; e1=$?; echo "attn tc pair exit $e1"; tail -n 2 gpurun_out/final_attn_tc1.log
if [ $e1 -ne 0 ]; then RVT_ATTN_BWD=2 timeout 45 python -m pytest tests/test_gpu_train_ops.py -k attn_core_bwd -q -m gpu -p no:cacheprovider --timeout 30 > gpurun_out/final_attn_tc2.log 2>&1; echo "attn tc split exit $?"; tail -n 2 gpurun_out/final_attn_tc2.log; fi
if [ $e1 -eq 0 ]; then RVT_ATTN_BWD=1 timeout 70 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/final_bench_train_tc.json 2> gpurun_out/final_bench_train_tc.err; echo "train tc exit $?"; head -c 300 gpurun_out/final_bench_train_tc.json; echo; RVT_ATTN_BWD=1 timeout 45 python -m pytest tests/test_gpu_train.py -k golden -q -m gpu -p no:cacheprovider --timeout 30 > gpurun_out/final_train_tc.log 2>&1; echo "train golden tc exit $?"; tail -n 2 gpurun_out/final_train_tc.log; fi
