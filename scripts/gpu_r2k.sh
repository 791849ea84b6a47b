cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
FASN_LIBRARY=$LIBDIR/libfasn_dk6.so timeout 600 python -m pytest tests -m gpu -q --timeout 180 -k "backward or random or properties" > gpurun_out/r2k_tests.log 2>&1; echo "tests(dk6) rc=$?"; tail -n 3 gpurun_out/r2k_tests.log | cut -c1-300
bash scripts/gpu_ab.sh "smoke_nothing_selected" "c3 c3nd" prev dk2 dk4 dk6 2>&1 | grep -v "^tests\|deselected\|no tests"
