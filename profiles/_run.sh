set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# remaining test modules (first run stopped at the 21-step taps case)
timeout 900 python -m pytest tests -m gpu -q --deselect "tests/test_gpu_backbone.py::test_operator_taps_match_oracle[rvt_b_1mpx_bs8_l21]" > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q --deselect "tests/test_gpu_backbone.py::test_operator_taps_match_oracle[rvt_b_1mpx_bs8_l21]" > gpurun_out/r2b_tests_fastln.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests_fastln.log
for f in 0 1; do
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_V2_FAST_LN=$f timeout 200 python profiles/op_bench.py --only attn > gpurun_out/r2b_attn_fastln$f.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_V2_FAST_LN=$f timeout 200 python profiles/op_bench.py --only mlp > gpurun_out/r2b_mlp_fastln$f.log 2>&1
done
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 120 python profiles/trace_v2.py attn > gpurun_out/r2b_trace_attn.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 120 python profiles/trace_v2.py mlp > gpurun_out/r2b_trace_mlp.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_GELU_F16X2=1 timeout 120 python profiles/trace_v2.py mlp > gpurun_out/r2b_trace_mlp_h2.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 120 python profiles/trace_v2.py attn --stage 1 > gpurun_out/r2b_trace_attn_s2.log 2>&1
# ncu: attn_v2 S1 and mlp_v2 S1 (one launch each, full set + source)
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_v2 -s 1 -c 1 -o gpurun_out/prof_attn_v2_s1 -f python profiles/op_bench.py --only attn --stage 0 --iters 1 > gpurun_out/r2b_ncu_attn.log 2>&1
RVT_ATTN_V2=1 RVT_MLP_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_v2 -s 1 -c 1 -o gpurun_out/prof_mlp_v2_s1 -f python profiles/op_bench.py --only mlp --stage 0 --iters 1 > gpurun_out/r2b_ncu_mlp.log 2>&1
RVT_LSTM_V2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backbone.py -q --deselect "tests/test_gpu_backbone.py::test_operator_taps_match_oracle[rvt_b_1mpx_bs8_l21]" > gpurun_out/r2b_tests_lstmv2.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests_lstmv2.log
RVT_LSTM_V2=0 timeout 200 python profiles/op_bench.py --only lstm > gpurun_out/r2b_lstm_v1.log 2>&1
RVT_LSTM_V2=1 timeout 200 python profiles/op_bench.py --only lstm > gpurun_out/r2b_lstm_v2.log 2>&1
RVT_LSTM_V2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_v2 -s 1 -c 1 -o gpurun_out/prof_lstm_v2_s1 -f python profiles/op_bench.py --only lstm --stage 0 --iters 1 > gpurun_out/r2b_ncu_lstm.log 2>&1
# CTA-count sensitivity of the persistent kernels
for n in 148 296 512 1024; do
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_V2_CTAS=$n timeout 200 python profiles/op_bench.py --only attn --stage 0 > gpurun_out/r2b_attn_ctas$n.log 2>&1
done
for n in 74 148 480 960; do
RVT_ATTN_V2=1 RVT_MLP_V2=1 RVT_V2_CTAS=$n timeout 200 python profiles/op_bench.py --only mlp --stage 0 > gpurun_out/r2b_mlp_ctas$n.log 2>&1
done
head -40 gpurun_out/r2b_trace_attn.log; head -24 gpurun_out/r2b_trace_mlp.log; head -24 gpurun_out/r2b_trace_mlp_h2.log; tail -25 gpurun_out/r2b_tests.log; tail -5 gpurun_out/r2b_tests_fastln.log; cat gpurun_out/r2b_attn_fastln*.log gpurun_out/r2b_mlp_fastln*.log; tail -5 gpurun_out/r2b_tests_lstmv2.log; cat gpurun_out/r2b_lstm_v1.log gpurun_out/r2b_lstm_v2.log
grep -h "attn\|mlp" gpurun_out/r2b_attn_ctas*.log gpurun_out/r2b_mlp_ctas*.log
