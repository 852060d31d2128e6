#!/bin/bash
# one pytest process per risk group so a faulting kernel cannot poison the CUDA context of the others
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 600 -s > gpurun_out/t_$name.log 2>&1; echo "$name exit $?" >> gpurun_out/t_summary.log; tail -n 3 gpurun_out/t_$name.log >> gpurun_out/t_summary.log; }
rm -f gpurun_out/t_summary.log
run tn_mn tests/test_gpu_train_ops.py -k gemm_tn_mn_major
run tn_k tests/test_gpu_train_ops.py -k gemm_tn_k_major
run ops_rest tests/test_gpu_train_ops.py -k "not gemm_tn"
run train tests/test_gpu_train.py
RVT_TN_MODE=1 run train_kmajor tests/test_gpu_train.py
run old tests/test_gpu_backbone.py tests/test_gpu_linear.py tests/test_gpu_ops.py tests/test_gpu_voxel.py
cat gpurun_out/t_summary.log
