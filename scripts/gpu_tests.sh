cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_bringup.py -m gpu -q --timeout 120 > gpurun_out/t_bringup.log 2>&1; echo "bringup rc=$?"
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout 180 > gpurun_out/t_fwd.log 2>&1; echo "fwd rc=$?"
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 180 > gpurun_out/t_bwd.log 2>&1; echo "bwd rc=$?"
tail -n 25 gpurun_out/t_bringup.log; tail -n 40 gpurun_out/t_fwd.log; tail -n 40 gpurun_out/t_bwd.log
