# round 2, session 2: forward with the exponent phases of the two Q tiles alternating (named-barrier hand-off between the two softmax warps of a
# scheduler): ap1 = at D = 64 only, ap2 = at both head dims; parity of the forward suites on ap2, then A/B
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
FASN_LIBRARY=$LIBDIR/libfasn_ap2.so timeout 600 python -m pytest tests -m gpu -q --timeout 180 -x -k "forward or random or configs or smoke" > gpurun_out/r2t_tests.log 2>&1; echo "tests(ap2) rc=$?"; tail -n 3 gpurun_out/r2t_tests.log | cut -c1-300
bash scripts/gpu_ab.sh "smoke_nothing_selected" "c2 c5 c3nd c3 c4" base ap1 ap2 2>&1 | grep -v "^tests\|deselected\|no tests"
