cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "n2 rc=$?"; head -c 400 gpurun_out/final_bench_n2.json; echo; grep -v Warning gpurun_out/final_bench_n2.err | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/final_bench_ref_n2.json 2> gpurun_out/final_bench_ref_n2.err; echo "ref n2 rc=$?"; head -c 300 gpurun_out/final_bench_ref_n2.json; echo; grep -v Warning gpurun_out/final_bench_ref_n2.err | tail -4
