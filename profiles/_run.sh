cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export RVT_CONV_SPLITK=0
B="python bench.py --steps 10 --warmup 3 --extras '' --no-cpu-baseline"
run() { n=$1; shift; env "$@" timeout 300 $B > gpurun_out/r2v_$n.json 2> gpurun_out/r2v_$n.err; python - gpurun_out/r2v_$n.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), round(d['e2e']['value']))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
}
run base A=1
run base2 A=1
run wbn64 RVT_WIDE_BN=64
run wbn256 RVT_WIDE_BN=256
run prio1 RVT_STREAM_PRIO=1
run prio3 RVT_STREAM_PRIO=3
run psm132 RVT_PERSIST_SMS=132
run cast256 RVT_LSTM_CAST_DIM=256
run tma70 RVT_TMA_SMEM_KB=70
run tma90 RVT_TMA_SMEM_KB=90
run gemm70 RVT_GEMM_SMEM_KB=70
run nogates RVT_FAST_GATES=0
run stemst0 RVT_STEM_TMA_STORE=0
run stem1 RVT_STEM_V2=1
