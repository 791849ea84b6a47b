# round 2, session 2: timeline of one persistent backward CTA over its items; fp32 envelope on both libraries
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
timeout 300 python scripts/timeline.py r2o 0 0 2>&1 | tail -2
timeout 300 python scripts/timeline.py r2o_c100 100 0 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -q --timeout 300 -k float32 2>&1 | tail -3
FASN_LIBRARY=$LIBDIR/libfasn_prev.so timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -q --timeout 300 -k float32 2>&1 | tail -3
