# round 2, session 2: forward with Q tile 1 of an item started half a step after Q tile 0 (one named-barrier hand-off per item): sk1 = at D = 64,
# sk2 = at both head dims
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
FASN_LIBRARY=$LIBDIR/libfasn_sk2.so timeout 600 python -m pytest tests -m gpu -q --timeout 180 -x -k "forward or random or configs or smoke" > gpurun_out/r2x_tests.log 2>&1; echo "tests(sk2) rc=$?"; tail -n 3 gpurun_out/r2x_tests.log | cut -c1-300
bash scripts/gpu_ab.sh "smoke_nothing_selected" "c2 c5 c3 c3nd c4" base sk1 sk2 2>&1 | grep -v "^tests\|deselected\|no tests"
