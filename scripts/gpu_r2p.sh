# round 2, session 2: persistent backward, second cut (K / V barriers split, epilogue stores on warp 15) against the per-tile-CTA backward (prev)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_ab.sh "backward or random or properties or configs or smoke" "c3 c3nd" base prev 2>&1
cp gpurun_out/ab_tests.log gpurun_out/r2p_tests.log
timeout 300 python scripts/timeline.py r2p 0 0 2>&1 | tail -1
