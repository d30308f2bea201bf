"""Phase timing of the persistent small-row decode kernel (decode_persist.cu): %globaltimer stamps of CTA 0 and CTA G-1 for the
LAST token of a generate() call. Usage: LG_PERSIST=1 python tools/persist_probe.py [GPT-L] [B]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LG_PERSIST", "1")
import torch
from llamagen_b200 import GPT_models, generate, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "GPT-L"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = 256
lib = _lib.load()
lib.lg_debug_set_pd_trace.argtypes = [ctypes.c_void_p]
torch.manual_seed(0)
m = GPT_models[name](block_size=S, vocab_size=16384)
m.output.weight.data.normal_(std=0.02)
m = m.to("cuda", torch.bfloat16).eval()
L = m.config.n_layer
cond = torch.randint(0, 1000, (B,), device="cuda")
kw = dict(cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True)
for _ in range(2):
    generate(m, cond, S, **kw)
trace = torch.zeros(2 * (L + 1) * 16, dtype=torch.int64, device="cuda")
lib.lg_debug_set_pd_trace(ctypes.c_void_p(trace.data_ptr()))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
generate(m, cond, S, **kw)
ev1.record()
torch.cuda.synchronize()
lib.lg_debug_set_pd_trace(ctypes.c_void_p(0))
t = trace.view(2, L + 1, 16).cpu().double()
names = ["P1 stage", "P1 gemm+epi", "barrier1", "P2 attn", "barrier2", "P3 merge", "P3 gemm+epi", "barrier3", "P4 stage", "P4 gemm+epi",
         "barrier4", "P5 stage", "P5 gemm+epi", "barrier5"]
print(f"{name} B={B}: generate {ev0.elapsed_time(ev1) * 1000 / S:.1f} us/token (events, incl. prefill + sampling)")
for c in range(2):
    lay = t[c, :L]
    d = lay[:, 1:15] - lay[:, 0:14]
    print(f"CTA {'0' if c == 0 else 'G-1'}: per-layer mean ns (layers 1..L-1) / layer 0:")
    for i, n in enumerate(names):
        print(f"   {n:14s} {d[1:, i].mean():8.0f}   {d[0, i]:8.0f}")
    print(f"   layer total    {(lay[1:, 14] - lay[1:, 0]).mean():8.0f}   token total {(t[c, L, 2] - t[c, 0, 0]):8.0f}   head stage {t[c, L, 1] - t[c, L, 0]:.0f} head gemm {t[c, L, 2] - t[c, L, 1]:.0f}")
