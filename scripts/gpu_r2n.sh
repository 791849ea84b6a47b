# round 2, session 2: persistent backward (base) against the per-tile-CTA backward (prev)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_ab.sh "backward or random or properties or configs or smoke" "c3 c3nd" base prev 2>&1
cp gpurun_out/ab_tests.log gpurun_out/r2n_tests.log
