# usage: bash scripts/gpu_ncu.sh [tag] -- launch list + one full ncu capture of the fwd and bwd kernels on the headline workload
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; TAG=${1:-r1}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fasn_.*_kernel -s 12 -c 4 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -n 3 gpurun_out/ncu_full_$TAG.log | cut -c1-300
