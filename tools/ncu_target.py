"""Small target for ncu captures: one generate() of a few tokens at the bench shape (GPT-L, B=64, cfg 4.0), eager launches
(LG_NO_GRAPH=1 recommended so every kernel is an ordinary launch). Usage: python tools/ncu_target.py [tokens] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llamagen_b200 import GPT_models, generate
S = int(sys.argv[1]) if len(sys.argv) > 1 else 24
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
m = GPT_models["GPT-L"](block_size=256, vocab_size=16384)
m.output.weight.data.normal_(std=0.02)
m = m.to("cuda", torch.bfloat16).eval()
labels = torch.randint(0, 1000, (B,), device="cuda")
generate(m, labels, S, cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True)
torch.cuda.synchronize()
print("ok")
