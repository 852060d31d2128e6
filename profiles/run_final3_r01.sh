#!/bin/bash
mkdir -p gpurun_out
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; echo "bench exit $?"; python -c "
import json; l=open('gpurun_out/final3_bench.json').read().strip().splitlines(); print(len(l), 'stdout line(s)'); d=json.loads(l[0]); print(d['value'], d['e2e']['value'], d['roofline']['frac'])"
timeout 60 python bench.py --mode train --steps 3 --warmup 3 > gpurun_out/final3_train.json 2> gpurun_out/final3_train.err; echo "train exit $?"; python -c "
import json; l=open('gpurun_out/final3_train.json').read().strip().splitlines(); print(len(l), 'stdout line(s)'); d=json.loads(l[0]); print(d['value'])"
