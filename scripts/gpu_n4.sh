cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/scale_r2z_n4.json 2> gpurun_out/scale_r2z_n4.err; echo "n=4 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_r2z_n4.json") if l.startswith("{")][-1])
print("n=%d value %.1f TFLOP/s per-gpu %.1f ms/step %.3f e2e %s" % (d["n_gpus"], d["value"], d["per_gpu_tflops"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
print("sharded", json.dumps(d.get("sharded"))[:600])
PY
