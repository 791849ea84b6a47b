# round 2, session 2: backward barrier experiments (per-warp arrivals = base, e0 = per-thread arrivals, e2 = spinning waits on the chain,
# e5 = dQ staging deferred behind dP^T, e25 = both)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_ab.sh "backward or random or properties" "c3 c3nd" base e0 e2 e5 e25 2>&1
cp gpurun_out/ab_tests.log gpurun_out/r2m_tests.log
