#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 -s > gpurun_out/t_$name.log 2>&1; echo "$name exit $?" >> gpurun_out/t_summary.log; tail -n 3 gpurun_out/t_$name.log >> gpurun_out/t_summary.log; }
rm -f gpurun_out/t_summary.log
run ops tests/test_gpu_train_ops.py
RVT_TN_EPI=0 run ops_atomics tests/test_gpu_train_ops.py -k gemm_tn_mn
run train tests/test_gpu_train.py
cat gpurun_out/t_summary.log
timeout 600 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_n1.json')); print('train', d['value'], d['ms_per_step'], d['phases_ms'])"; tail -n 5 gpurun_out/bench_train_n1.err
for p in 0 1 2 3; do RVT_STREAM_PRIO=$p timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prio$p.json 2> gpurun_out/bench_prio$p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_prio$p.json')); print('prio$p', d['value'], d['e2e']['value'])"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv python profiles/train_small.py 2 > gpurun_out/train_small.log 2>&1; tail -n 2 gpurun_out/train_small.log
