# usage: bash scripts/gpu_impl_sweep.sh  -- single-CTA vs CTA-pair backward at short and sustained timed regions
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for cfg in "0 50" "2 50" "0 300" "2 300" "0 50" "2 50"; do
  set -- $cfg
  FASN_BWD_IMPL=$1 timeout 300 python bench.py --workload c3 --steps $2 --warmup 3 --no-cpu --no-e2e > gpurun_out/impl_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/impl_$1_$2.json")); r=d["roofline"]
print("impl $1 steps $2: %.1f TFLOP/s  %.3f ms  fwd %.3f  main %.3f  clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"], d["clocks"]))
PY
done
