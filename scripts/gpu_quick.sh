# usage: bash scripts/gpu_quick.sh [tag]  -- full GPU test suite + headline bench lines (no profiler)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; TAG=${1:-q}
timeout 1200 python -m pytest tests -m gpu -q --timeout 180 -x > gpurun_out/t_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/t_all_$TAG.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c3_$TAG.json"))
    r=d["roofline"]
    print("c3: value %.1f TFLOP/s  ms/step %.3f  fwd %.3f ms (%.0f TF)  bwd-main %.3f ms (%.0f TF) e2e %.1f clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["fwd_kernel_tflops"], r["kernel_ms"], r["achieved"], (d["e2e"] or {}).get("value", 0), d["clocks"]))
except Exception as e:
    print("bench parse failed", e)
PY
tail -n 5 gpurun_out/bench_c3_$TAG.err
for wl in c3nd c3pad c3alibi c2 c4 c5; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}_$TAG.json")); r=d["roofline"]; print("$wl: %.1f TFLOP/s  %.3f ms  (fwd kernel %.3f ms, main kernel %.3f ms)" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"]))
except Exception as e:
    print("$wl failed", e)
PY
done
