cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout 300 -k "graph" > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/r2y_tests.log | cut -c1-300
