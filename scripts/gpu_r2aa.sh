# round 2, session 2: shared-K/V head reduction inside the backward epilogue (FasnParams.dk_accum / dv_accum): full GPU suite + headline check
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2aa_tests.log | cut -c1-300
for rep in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('c3: %.1f TFLOP/s %.3f ms fwd %.3f bwd-main %.3f' % (d['value'], d['ms_per_step'], r['fwd_kernel_ms'], r['kernel_ms']))"; done
python - <<'PY'
# multi-query shape: H = 32 query heads on one shared K/V head, S = 2048, D = 128: time of the backward with the in-kernel head sum
import torch, time, sys
sys.path[:0] = ['.', 'flash-attention-softmax-n_b200']
from flash_attention_softmax_n import flash_attention_n
B, H, S, D = 4, 32, 2048, 128
q = torch.randn(B, H, S, D, device='cuda', dtype=torch.bfloat16).mul_(0.5).requires_grad_()
k = torch.randn(B, S, D, device='cuda', dtype=torch.bfloat16).mul_(0.5).requires_grad_()
v = torch.randn(B, S, D, device='cuda', dtype=torch.bfloat16).mul_(0.5).requires_grad_()
do = torch.randn(B, H, S, D, device='cuda', dtype=torch.bfloat16)
for i in range(3):
    o = flash_attention_n(q, k, v, softmax_n_param=1.0, is_causal=True); o.backward(do)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20):
    q.grad = k.grad = v.grad = None
    o = flash_attention_n(q, k, v, softmax_n_param=1.0, is_causal=True); o.backward(do)
b.record(); torch.cuda.synchronize()
print("MQA B4 H32 (1 K/V head) S2048 D128 causal bf16 fwd+bwd: %.3f ms/step, peak mem %.0f MB" % (a.elapsed_time(b) / 20, torch.cuda.max_memory_allocated() / 2**20))
PY
