#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 -s > gpurun_out/t_$name.log 2>&1; echo "$name exit $?" >> gpurun_out/t_summary.log; tail -n 3 gpurun_out/t_$name.log >> gpurun_out/t_summary.log; }
rm -f gpurun_out/t_summary.log
run ops tests/test_gpu_train_ops.py
run train tests/test_gpu_train.py
cat gpurun_out/t_summary.log
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_n1.json')); print('train graph+wavefront', d['value'], d['ms_per_step'], d['phases_ms'], d['final_loss'], d['grads_finite'])"; grep -v Warning gpurun_out/bench_train_n1.err | tail -n 12
timeout 600 python bench.py --mode train --no-wavefront --steps 5 --warmup 3 > gpurun_out/bench_train_nowf.json 2> gpurun_out/bench_train_nowf.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_nowf.json')); print('train graph, 1 stream', d['value'], d['ms_per_step'], d['final_loss'])"; grep -v Warning gpurun_out/bench_train_nowf.err | tail -n 5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv python profiles/train_small.py 2 > gpurun_out/train_small.log 2>&1; tail -n 2 gpurun_out/train_small.log
