# usage: bash scripts/gpu_ab.sh "<pytest -k expr>" "<workloads>" variant...   -- GPU tests on libfasn.so, then A/B bench lines (20 steps) per library variant, twice
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
K="$1"; WLS="$2"; shift; shift
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "$K" > gpurun_out/ab_tests.log 2>&1; else timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/ab_tests.log 2>&1; fi
echo "tests rc=$?"; tail -n 6 gpurun_out/ab_tests.log | cut -c1-300
for rep in 1 2; do
for v in "$@"; do
  export FASN_LIBRARY=$LIBDIR/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$LIBDIR/libfasn.so
  for wl in $WLS; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e --sustained-seconds 0 > gpurun_out/ab_${v}_${wl}_$rep.json 2>gpurun_out/ab_${v}_${wl}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${v}_${wl}_$rep.json").read().strip().splitlines()[-1]); r=d["roofline"]; print("$v $wl #$rep: %.1f TFLOP/s  %.3f ms  fwd %.3f ms  bwd-main %.3f ms  clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$v $wl failed", e); print(open("gpurun_out/ab_${v}_${wl}_$rep.err").read()[-800:])
PY
  done
done
done
