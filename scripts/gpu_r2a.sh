# round 2, first GPU call: full GPU suite on the quad-layout backward, the fold variant's first run, A/B of the three libraries,
# phase timeline and ncu of the new backward kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LIBDIR=$GRAFT_REPO_ROOT/flash-attention-softmax-n_b200/flash_attention_softmax_n
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/r2a_tests.log | cut -c1-400
FASN_LIBRARY=$LIBDIR/libfasn.so timeout 300 python scripts/fold_check.py > gpurun_out/r2a_check_base.log 2>&1; echo "check base rc=$?"; tail -n 6 gpurun_out/r2a_check_base.log | cut -c1-300
FASN_LIBRARY=$LIBDIR/libfasn_fold.so timeout 300 python scripts/fold_check.py > gpurun_out/r2a_check_fold.log 2>&1; echo "check fold rc=$?"; tail -n 6 gpurun_out/r2a_check_fold.log | cut -c1-300
for rep in 1 2; do
for v in base r1 fold; do
  export FASN_LIBRARY=$LIBDIR/libfasn_$v.so
  [ "$v" = "base" ] && export FASN_LIBRARY=$LIBDIR/libfasn.so
  for wl in c3 c3nd; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_${v}_${wl}_$rep.json 2>gpurun_out/r2a_${v}_${wl}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2a_${v}_${wl}_$rep.json")); r=d["roofline"]; print("$v $wl #$rep: %.1f TFLOP/s  %.3f ms  fwd %.3f ms  bwd-main %.3f ms  clocks %s" % (d["value"], d["ms_per_step"], r["fwd_kernel_ms"], r["kernel_ms"], d["clocks"]))
except Exception as e:
    print("$v $wl failed", e); print(open("gpurun_out/r2a_${v}_${wl}_$rep.err").read()[-800:])
PY
  done
done
done
unset FASN_LIBRARY
timeout 200 python scripts/timeline.py r2a 4 70 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_r2a.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fasn_.*_kernel -s 12 -c 4 -o gpurun_out/prof_r2a -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_r2a.log 2>&1; echo "ncu full rc=$?"; tail -n 3 gpurun_out/ncu_full_r2a.log | cut -c1-300
bash scripts/gpu_smem_metrics.sh base r1 2>&1 | tail -30
