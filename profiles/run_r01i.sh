#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 -s > gpurun_out/t_$name.log 2>&1; echo "$name exit $?" >> gpurun_out/t_summary.log; tail -n 3 gpurun_out/t_$name.log >> gpurun_out/t_summary.log; }
rm -f gpurun_out/t_summary.log
RVT_ATTN_BWD=1 run attn_tc_pair tests/test_gpu_train_ops.py -k attn_core_bwd
RVT_ATTN_BWD=2 run attn_tc_split tests/test_gpu_train_ops.py -k attn_core_bwd
RVT_ATTN_BWD=0 run attn_simt tests/test_gpu_train_ops.py -k attn_core_bwd
run train tests/test_gpu_train.py
RVT_ATTN_BWD=2 run train_split tests/test_gpu_train.py -k golden
cat gpurun_out/t_summary.log
for m in 1 2 0; do RVT_ATTN_BWD=$m timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_attn$m.json 2> gpurun_out/bench_train_attn$m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_attn$m.json')); print('train attn_bwd mode $m', d['value'], d['ms_per_step'], d['final_loss'], d['grads_finite'])"; grep -v Warning gpurun_out/bench_train_attn$m.err | tail -n 3; done
