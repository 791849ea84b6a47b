# usage (under gpurun --gpus N): bash scripts/gpu_scale.sh N [tag]  -- bench.py at 1 and N GPUs (both arms), JSON lines kept under gpurun_out/
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; N=${1:-8}; TAG=${2:-s}
nvidia-smi -L | head -8
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err; echo "n=1 rc=$?"
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 30 --warmup 3 > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err; echo "n=$n rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_${TAG}_n*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print("n=%d value %.1f TFLOP/s per-gpu %.1f ms/step %.3f e2e %s" % (d["n_gpus"], d["value"], d["per_gpu_tflops"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
    except Exception as e:
        print(f, "parse failed", e)
PY
