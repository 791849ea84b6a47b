# usage: bash scripts/gpu_bench.sh [tag]   -- tests that failed last time, smoke, bench lines, ncu launch list + full capture
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; TAG=${1:-r1}
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 180 > gpurun_out/t_bwd.log 2>&1; echo "bwd rc=$?"; tail -n 3 gpurun_out/t_bwd.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 6 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_c3_$TAG.json; tail -n 5 gpurun_out/bench_c3_$TAG.err
for wl in c2 c4 c5; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; cat gpurun_out/bench_${wl}_$TAG.json; tail -n 3 gpurun_out/bench_${wl}_$TAG.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fasn_.*_kernel -s 12 -c 4 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -n 5 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/
