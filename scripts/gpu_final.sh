# usage: bash scripts/gpu_final.sh [tag]  -- full GPU suite, smoke, every bench line (ours + reference arm), HF route bench, ncu launch list + full capture
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; TAG=${1:-r1f}
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/t_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_all_$TAG.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_c3_$TAG.json; tail -n 3 gpurun_out/bench_c3_$TAG.err
for wl in c2 c4 c5 c3nd; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?"; cut -c1-260 gpurun_out/bench_${wl}_$TAG.json | cut -c100-260
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_reference_$TAG.json
timeout 300 python scripts/bench_hf.py > gpurun_out/hf_$TAG.log 2>&1; echo "hf rc=$?"; tail -n 3 gpurun_out/hf_$TAG.log; cp gpurun_out/hf_bench.json gpurun_out/hf_bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fasn_.*_kernel -s 12 -c 4 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -n 2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
