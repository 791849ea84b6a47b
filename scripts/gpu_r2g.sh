# round 2, call g: new tests (configs, fp32, dBias, debug32) + full suite
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -n 40 gpurun_out/r2g_tests.log | cut -c1-400
