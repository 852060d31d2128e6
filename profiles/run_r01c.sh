#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 -s > gpurun_out/t_$name.log 2>&1; echo "$name exit $?" >> gpurun_out/t_summary.log; tail -n 3 gpurun_out/t_$name.log >> gpurun_out/t_summary.log; }
rm -f gpurun_out/t_summary.log
run ops tests/test_gpu_train_ops.py
run train tests/test_gpu_train.py
cat gpurun_out/t_summary.log
timeout 600 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; tail -c 700 gpurun_out/bench_train_n1.json; tail -n 5 gpurun_out/bench_train_n1.err
timeout 600 python bench.py --impl reference-gpu --steps 3 --warmup 3 > gpurun_out/bench_refgpu.json 2> gpurun_out/bench_refgpu.err; cat gpurun_out/bench_refgpu.json; tail -n 3 gpurun_out/bench_refgpu.err
timeout 600 python bench.py --impl reference-gpu --mode train --steps 2 --warmup 3 > gpurun_out/bench_refgpu_train.json 2> gpurun_out/bench_refgpu_train.err; cat gpurun_out/bench_refgpu_train.json; tail -n 3 gpurun_out/bench_refgpu_train.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv python profiles/train_small.py 2 > gpurun_out/train_small.log 2>&1; tail -n 2 gpurun_out/train_small.log
