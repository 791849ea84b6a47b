# round 2, call d: persistent forward kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
timeout 900 python -m pytest tests -m gpu -q --timeout 180 -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2d_tests.log | cut -c1-300
for rep in 1 2; do
for v in base r1; do
  export FASN_LIBRARY=$LIBDIR/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$LIBDIR/libfasn.so
  for wl in c3 c3nd c2 c4 c5; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2d_${v}_${wl}_$rep.json 2>gpurun_out/r2d_${v}_${wl}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_${v}_${wl}_$rep.json")); r=d["roofline"]; print("$v $wl #$rep: %.1f TFLOP/s  %.3f ms  fwd %.3f ms  bwd-main %.3f ms  clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$v $wl failed", e); print(open("gpurun_out/r2d_${v}_${wl}_$rep.err").read()[-800:])
PY
  done
done
done
unset FASN_LIBRARY
timeout 200 python scripts/cta_profile.py r2d 0.1 2>&1 | tail -24
