# usage: bash scripts/gpu_iter.sh "<pytest -k expr or empty for all>" variant1 variant2 ...  -- tests on the base lib, then bench c3 (+c3nd) per variant ("base" = libfasn.so)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K="$1"; shift
if [ -n "$K" ]; then timeout 900 python -m pytest tests -m gpu -q --timeout 180 -x -k "$K" 2>&1 | tail -n 4; else timeout 1200 python -m pytest tests -m gpu -q --timeout 180 -x 2>&1 | tail -n 4; fi
for v in "$@"; do
  export FASN_LIBRARY=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn.so
  for wl in ${WLS:-c3}; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/iter_${v}_$wl.json 2>gpurun_out/iter_${v}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/iter_${v}_$wl.json")); r=d["roofline"]; print("$v $wl: %.1f TFLOP/s  %.3f ms  fwd %.3f ms  bwd-main %.3f ms" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"]))
except Exception as e:
    print("$v $wl failed", e); print(open("gpurun_out/iter_${v}_$wl.err").read()[-1500:])
PY
  done
done
