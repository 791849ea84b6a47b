# round 2, call f (2 GPUs): NCCL / IPC sharded leg of bench.py and the 2-GPU bit-identity test
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20 > gpurun_out/r2f_topo.txt; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/r2f_topo.txt; cat gpurun_out/r2f_topo.txt
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 2>&1 | tail -n 3
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err; echo "bench rc=$?"; tail -n 12 gpurun_out/r2f_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_bench_n$N.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "sustained", d["sustained"] and d["sustained"]["value"])
    print("sharded", json.dumps(d["sharded"], indent=1))
    print("e2e", json.dumps(d["e2e"], indent=1))
except Exception as e: print("parse failed", e)
PY
