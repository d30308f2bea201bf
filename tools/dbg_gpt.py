import os, sys, time, faulthandler
faulthandler.dump_traceback_later(240, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
t0 = time.time()
from util import build_gpt, load_golden
from llamagen_b200 import generate
print("import", time.time() - t0, flush=True)
g = load_golden("gpt_c2i.pt")
m = build_gpt(g["cfg"], g["state_dict"], torch.float32)
print("built", time.time() - t0, flush=True)
for cfg in (1.0, 4.0):
    toks, logits = generate(m, g["cond"].cuda(), g["S"], cfg_scale=cfg, sample_logits=False, return_logits=True)
    torch.cuda.synchronize()
    print("generated", cfg, time.time() - t0, flush=True)
    print(torch.equal(toks.cpu(), g[f"tokens_cfg{cfg}"]), (logits.cpu() - g[f"logits_cfg{cfg}"]).abs().max().item(), flush=True)
