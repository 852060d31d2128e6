cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -q -x -m gpu > gpurun_out/r2t_tests_variants.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2t_tests_variants.log
timeout 300 python profiles/wavefront_timeline.py > gpurun_out/r2t_timeline.log 2>&1; head -6 gpurun_out/r2t_timeline.log
