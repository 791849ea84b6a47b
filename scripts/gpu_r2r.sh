# round 2, session 2: persistent backward, dK / dV stored from registers, V / K / dO of the next item prefetched in that order (base)
# against the first persistent cut (p1) and the per-tile-CTA backward (prev)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_ab.sh "backward or random or properties or configs or smoke" "c3 c3nd" base p1 prev 2>&1
cp gpurun_out/ab_tests.log gpurun_out/r2r_tests.log
timeout 300 python scripts/timeline.py r2r 0 0 2>&1 | tail -1
