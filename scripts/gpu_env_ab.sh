# usage: bash scripts/gpu_env_ab.sh VAR "v1 v2 ..." "<pytest -k expr>"   -- A/B an environment switch: tests with the LAST value, then bench $WLS for each value (twice)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
VAR=$1; VALS=$2; K=$3
last=$(echo $VALS | awk '{print $NF}')
env $VAR=$last timeout 900 python -m pytest tests -m gpu -q --timeout 180 -x -k "$K" 2>&1 | tail -n 4
for rep in 1 2; do for v in $VALS; do for wl in ${WLS:-c3}; do
  env $VAR=$v timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_${v}_$wl.json 2>gpurun_out/ab_${v}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${v}_$wl.json")); r=d["roofline"]; print("$VAR=$v $wl: %.1f TFLOP/s  %.3f ms  fwd %.3f ms  bwd-main %.3f ms" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"]))
except Exception as e:
    print("$VAR=$v $wl failed", e); print(open("gpurun_out/ab_${v}_$wl.err").read()[-800:])
PY
done; done; done
