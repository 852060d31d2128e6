set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export RVT_CONV_TMA=1
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
timeout 200 python profiles/op_bench.py > gpurun_out/r2k_opbench.log 2>&1
for kb in 150 200; do RVT_TMA_SMEM_KB=$kb timeout 400 $B > gpurun_out/r2k_bench_tma$kb.json 2> gpurun_out/r2k_bench_tma$kb.err; RVT_TMA_SMEM_KB=$kb timeout 200 python profiles/op_bench.py > gpurun_out/r2k_opbench_tma$kb.log 2>&1; done
RVT_GEMM_SMEM_KB=200 timeout 400 $B > gpurun_out/r2k_bench_gemm200.json 2> gpurun_out/r2k_bench_gemm200.err
RVT_TMA_SMEM_KB=70 timeout 400 $B > gpurun_out/r2k_bench_tma70.json 2> gpurun_out/r2k_bench_tma70.err
timeout 400 $B > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
for f in gpurun_out/r2k_bench*.json; do echo $f; cut -c1-120 $f; done
tail -18 gpurun_out/r2k_opbench.log; tail -18 gpurun_out/r2k_opbench_tma200.log
