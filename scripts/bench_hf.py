"""Hugging Face encoder routed to the fused kernels: ms per training step (forward + backward) of a randomly initialised
BERT-large-shaped encoder (24 layers, hidden 1024, 16 heads of 64, fp16) at B=16, L=512 with right-padded sequences, for
the library's own SDPA route (softmax_0 only), the eager softmax_n route (what the reference's surgery computes,
surgery_functions/_bert.py:73-111) and the fused softmax_n route.  Writes gpurun_out/hf_bench.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-softmax-n_b200")]
import torch
from transformers import BertConfig, BertModel
from flash_attention_softmax_n.surgery import EAGER, FUSED, apply_attention_softmax_n

layers = int(os.environ.get("HF_LAYERS", "24"))
B, L = 16, 512
torch.manual_seed(0)
cfg = BertConfig(hidden_size=1024, num_attention_heads=16, num_hidden_layers=layers, intermediate_size=4096, vocab_size=30522,
                 max_position_embeddings=512, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.1)
model = BertModel(cfg, add_pooling_layer=False).cuda().half().train()
ids = torch.randint(0, 30522, (B, L), device="cuda")
am = torch.ones(B, L, dtype=torch.long, device="cuda")
for b in range(B):
    am[b, L - 16 * b:] = 0                      # 0 .. 240 padded positions
res = {"config": f"BERT-large shape, {layers} layers, fp16, B={B} L={L}, attention dropout 0.1, right padding 0..240"}
for name, impl, n in (("sdpa_softmax0", "sdpa", None), ("eager_softmax_n", EAGER, 1.0), ("fused_softmax_n", FUSED, 1.0)):
    if n is None:
        model.config._attn_implementation = impl
    else:
        apply_attention_softmax_n(model, n, implementation=impl)
    def step():
        model.zero_grad(set_to_none=True)
        model(input_ids=ids, attention_mask=am).last_hidden_state.float().square().mean().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record(); torch.cuda.synchronize()
    res[name + "_ms_per_step"] = e0.elapsed_time(e1) / 10
    print(name, res[name + "_ms_per_step"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "hf_bench.json"), "w"), indent=1)
