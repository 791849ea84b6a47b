# round 2, call b: reducers drain the whole dQ tile at once (+ split dS hand-off variant), A/B against round 1
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"; tail -n 8 gpurun_out/r2b_tests.log | cut -c1-300
FASN_LIBRARY=$LIBDIR/libfasn_split.so timeout 600 python -m pytest tests -m gpu -q --timeout 180 -k "backward or random or properties" > gpurun_out/r2b_tests_split.log 2>&1; echo "tests(split) rc=$?"; tail -n 4 gpurun_out/r2b_tests_split.log | cut -c1-300
for rep in 1 2; do
for v in base split r1; do
  export FASN_LIBRARY=$LIBDIR/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$LIBDIR/libfasn.so
  for wl in c3 c3nd; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_${v}_${wl}_$rep.json 2>gpurun_out/r2b_${v}_${wl}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_${v}_${wl}_$rep.json")); r=d["roofline"]; print("$v $wl #$rep: %.1f TFLOP/s  %.3f ms  fwd %.3f ms  bwd-main %.3f ms  clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$v $wl failed", e); print(open("gpurun_out/r2b_${v}_${wl}_$rep.err").read()[-800:])
PY
  done
done
done
unset FASN_LIBRARY
timeout 200 python scripts/timeline.py r2b 4 70 2>&1 | tail -2
