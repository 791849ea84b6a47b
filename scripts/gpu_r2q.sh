# round 2, session 2: persistent backward with K / dO swapping buffers per item (base) against the first persistent cut (p1) and the per-tile-CTA backward (prev)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_ab.sh "backward or random or properties or configs or smoke" "c3 c3nd" base p1 prev 2>&1
cp gpurun_out/ab_tests.log gpurun_out/r2q_tests.log
timeout 300 python scripts/timeline.py r2q 0 0 2>&1 | tail -1
