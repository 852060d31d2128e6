#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_ops.py -q -x -m gpu -p no:cacheprovider --timeout 60 > gpurun_out/final2_tests.log 2>&1; echo "train tests (tcgen05 attn bwd default) exit $?"; tail -n 2 gpurun_out/final2_tests.log
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --mode train --gpus 2 --steps 5 --warmup 3 > gpurun_out/final2_bench_train_n2.json 2> gpurun_out/final2_bench_train_n2.err; echo "train n2 exit $?"; tail -c 900 gpurun_out/final2_bench_train_n2.json; echo; grep -v Warning gpurun_out/final2_bench_train_n2.err | tail -n 4
timeout 60 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/final2_bench_train_n1.json 2> gpurun_out/final2_bench_train_n1.err; echo "train n1 exit $?"; head -c 250 gpurun_out/final2_bench_train_n1.json; echo
