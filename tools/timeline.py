"""Raw kernel timeline (name, stream, start, end in us) of a window of decode steps in the middle of one generate() call,
from CUPTI via torch.profiler. PDL and the dual-chain split stay as configured by the environment, so this shows how the two
chains' kernels actually overlap. Usage: python tools/timeline.py --out gpurun_out/timeline.json [--batch 64] [--window 1200]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from llamagen_b200 import GPT_models, generate

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--gpt-model", default="GPT-L")
ap.add_argument("--tokens", type=int, default=256)
ap.add_argument("--window", type=int, default=1200, help="number of kernel records kept (from the middle of the loop)")
ap.add_argument("--out", required=True)
args = ap.parse_args()
torch.manual_seed(0)
dev = "cuda"
gpt = GPT_models[args.gpt_model](block_size=args.tokens, vocab_size=16384)
gpt.output.weight.data.normal_(std=0.02)
gpt = gpt.to(dev, torch.bfloat16).eval()
labels = torch.randint(0, 1000, (args.batch,), device=dev)
kw = dict(cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True)
for _ in range(2):
    generate(gpt, labels, args.tokens, **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    generate(gpt, labels, args.tokens, **kw)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
mid = len(evs) // 2
win = evs[max(0, mid - args.window // 2): mid + args.window // 2]
t0 = win[0].time_range.start
rows = []
for e in win:
    nm = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:48]
    stream = getattr(e, "stream", None)
    if stream is None:
        stream = getattr(e, "device_resource_id", -1)
    rows.append([nm, int(stream) if stream is not None else -1, round(e.time_range.start - t0, 2), round(e.time_range.end - t0, 2)])
span = evs[-1].time_range.end - evs[0].time_range.start
json.dump({"total_kernels": len(evs), "span_us": span, "events": rows}, open(args.out, "w"))
print(json.dumps({"total_kernels": len(evs), "span_us": span, "window": len(rows)}))
