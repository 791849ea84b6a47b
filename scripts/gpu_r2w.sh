# round 2, session 2: share of the exponentials computed as a polynomial at D = 64: base 25 %, pp4 12.5 %, np64 none
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash scripts/gpu_ab.sh "smoke_nothing_selected" "c2 c5" base pp4 np64 2>&1 | grep -v "^tests\|deselected\|no tests"
