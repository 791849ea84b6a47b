# usage: bash scripts/gpu_steps_sweep.sh  -- headline workload at several timed-region lengths (clock / power behaviour under sustained load)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=power.limit,power.default_limit,power.max_limit,clocks.max.sm,temperature.gpu --format=csv
for cfg in "c3 20" "c3 50" "c3 200" "c3 600" "c3bf16 200" "c3nd 200" "c3 20"; do
  set -- $cfg
  timeout 300 python bench.py --workload $1 --steps $2 --warmup 3 --no-cpu --no-e2e > gpurun_out/sweep_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$1_$2.json")); r=d["roofline"]
print("$1 steps $2: %.1f TFLOP/s  %.3f ms  fwd %.3f  main %.3f  clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"], d["clocks"]))
PY
done
