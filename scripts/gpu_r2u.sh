# round 2, session 2: forward with the exponentials of the first half tile issued before the row maximum is known (base: D = 128 only, spec2: both
# head dims; nospec: the previous order); parity of the forward suites on spec2 and base, then A/B
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
FASN_LIBRARY=$LIBDIR/libfasn_spec2.so timeout 600 python -m pytest tests -m gpu -q --timeout 180 -k "forward or random or configs or smoke or properties" > gpurun_out/r2u_tests_spec2.log 2>&1; echo "tests(spec2) rc=$?"; tail -n 3 gpurun_out/r2u_tests_spec2.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q --timeout 180 -k "forward or random or configs or smoke or properties" > gpurun_out/r2u_tests.log 2>&1; echo "tests(base) rc=$?"; tail -n 3 gpurun_out/r2u_tests.log | cut -c1-300
bash scripts/gpu_ab.sh "smoke_nothing_selected" "c3 c3nd c4 c2 c5" base nospec spec2 2>&1 | grep -v "^tests\|deselected\|no tests"
