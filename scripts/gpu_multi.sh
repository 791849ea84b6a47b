# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh N [tag]
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; N=${1:-2}; TAG=${2:-m}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 > gpurun_out/t_multi_$TAG.log 2>&1; echo "multi test rc=$?"; tail -n 3 gpurun_out/t_multi_$TAG.log
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  fi
  echo "bench n=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_${TAG}_n$n.json") if l.startswith("{")][-1]); print("n=%d value %.1f TFLOP/s per-gpu %.1f ms/step %.3f e2e %s" % (d["n_gpus"], d["value"], d["per_gpu_tflops"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
except Exception as e:
    print("parse failed", e)
PY
  tail -n 3 gpurun_out/scale_${TAG}_n$n.err
done
