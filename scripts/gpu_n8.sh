# usage (under gpurun --gpus 8): bash scripts/gpu_n8.sh [tag] -- bench.py at 8 GPUs exactly as the driver launches it (weak-scaling value + sharded C4 leg)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; TAG=${1:-r2z}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/scale_${TAG}_n8.json 2> gpurun_out/scale_${TAG}_n8.err; echo "n=8 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/scale_${TAG}_n8.json") if l.startswith("{")][-1])
print("n=%d value %.1f TFLOP/s per-gpu %.1f ms/step %.3f e2e %s" % (d["n_gpus"], d["value"], d["per_gpu_tflops"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
print("sharded", json.dumps(d.get("sharded"))[:900])
print("clocks", d["clocks"])
PY
tail -n 3 gpurun_out/scale_${TAG}_n8.err | cut -c1-300
