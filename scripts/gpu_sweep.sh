# usage: bash scripts/gpu_sweep.sh lib1 lib2 ...   -- bench c3/c4/c2 for each libfasn_<name>.so variant
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "$@"; do
  export FASN_LIBRARY=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n/libfasn.so
  for wl in c3 c4 c2 c5; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/sweep_${v}_$wl.json 2>/dev/null
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${v}_$wl.json")); r=d["roofline"]; print("$v $wl: %.1f TFLOP/s  %.3f ms  fwd %.3f ms" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"]))
except Exception as e:
    print("$v $wl failed", e)
PY
  done
done
