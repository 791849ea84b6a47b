# round 2, session 2: full GPU suite with the native-low-precision criterion live at every call site
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -n 30 gpurun_out/r2v_tests.log | cut -c1-300
