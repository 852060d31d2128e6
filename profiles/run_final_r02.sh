#!/bin/bash
# Round-2 evidence run (one GPU).  Every command has its own timeout; outputs land in gpurun_out/ and are summarised into profiles/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/final_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final_tests.log; tail -3 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/final_bench.json; echo; grep -v Warning gpurun_out/final_bench.err | tail -5
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/final_bench_ref.json; echo
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --extras '' --no-cpu-baseline > gpurun_out/final_launches_bench.log 2>&1; echo "launch list rc=$?"
# every kernel of one timestep, full sections
timeout 900 ncu --set full --clock-control none -k 'regex:^(gemm_fused|attn_v2|mlp_v2|lstm_v2|stem_v2|ln_rows|cast_xh|attention_core|attn_fused|mlp_fused)' -s 96 -c 48 -f -o gpurun_out/r02_step python profiles/one_timestep.py --steps 3 > gpurun_out/final_ncu_step.log 2>&1; echo "ncu step rc=$?"; tail -2 gpurun_out/final_ncu_step.log
# the dominant kernel with source correlation, and the voxelizer
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^attn_v2 -s 4 -c 1 -f -o gpurun_out/r02_attn_s1 python profiles/one_timestep.py --steps 3 > gpurun_out/final_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:^voxel -c 2 -f -o gpurun_out/r02_voxel python profiles/one_timestep.py --steps 1 --voxel > gpurun_out/final_ncu_voxel.log 2>&1; echo "ncu voxel rc=$?"
for w in attn mlp; do timeout 120 python profiles/trace_v2.py $w > gpurun_out/final_trace_$w.log 2>&1; done
timeout 120 python profiles/trace_v2.py stem > gpurun_out/final_trace_stem.log 2>&1
timeout 200 python profiles/op_bench.py --json gpurun_out/final_opbench.json > gpurun_out/final_opbench.log 2>&1
# summarise on the box; the 48-kernel report is too large to travel (gpurun_out/ is capped at 64 MiB), its raw CSV export is not
python profiles/ncu_table.py gpurun_out/r02_step.ncu-rep gpurun_out/r02_voxel.ncu-rep --csv gpurun_out/ncu_r02_csv > gpurun_out/ncu_r02_table.md 2> gpurun_out/ncu_table.err; echo "ncu table rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/ncu_r02_csv
rm -f gpurun_out/r02_step.ncu-rep
du -sh gpurun_out
