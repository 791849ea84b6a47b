# final sanity on the committed tree: full GPU suite, smoke, the driver's bench command (both arms)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/last_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/last_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/last_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/last_bench_reference.json 2> gpurun_out/last_bench_reference.err; echo "reference rc=$?"; cut -c1-260 gpurun_out/last_bench_reference.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/last_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value %.1f TFLOP/s  %.3f ms/step  frac_of_burst %.3f  roofline frac %.3f (%s) kernel_ms %.3f fwd %.3f  e2e %.1f  cpu %.3f  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["frac_of_burst_peak"], r["frac"], r["regime"], r["kernel_ms"], r["fwd_kernel_ms"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"]))
print("sustained", d["sustained"]["value"], d["sustained"]["clocks"]["sm_mhz"])
PY
