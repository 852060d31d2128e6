#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_train_n1.json')); print('train graph', d['value'], d['ms_per_step'], d['phases_ms'], d['final_loss'], d['grads_finite'])"; grep -v Warning gpurun_out/bench_train_n1.err | tail -n 12
