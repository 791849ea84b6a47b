# round 2, call e: bench.py smoke on 1 GPU (new JSON fields) after the split-dS removal
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 -x -k "backward or smoke or properties" > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/r2e_tests.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -n 5 gpurun_out/r2e_bench.err | cut -c1-400
python - <<PY
import json
d=json.load(open("gpurun_out/r2e_bench.json"))
for k in ("value","ms_per_step","frac_of_peak","frac_of_burst_peak","frac_of_sustained_peak","sustained","clocks","cpu_baseline"): print(k, d.get(k))
print("roofline", {k:v for k,v in d["roofline"].items() if k in ("achieved","peak","frac","regime","kernel_ms","fwd_kernel_ms","frac_of_burst_peak","frac_of_sustained_peak")})
print("e2e", d["e2e"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-600
