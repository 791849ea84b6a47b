cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 -k "mask or bias or golden" 2>&1 | tail -3
for wl in c3pad c3alibi; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/aux_$wl.json 2>gpurun_out/aux_$wl.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/aux_$wl.json")); r=d["roofline"]; print("$wl: %.1f TFLOP/s  %.3f ms  (fwd kernel %.3f ms, main kernel %.3f ms)" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"]))
except Exception as e:
    print("$wl failed", e)
PY
done
