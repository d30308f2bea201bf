"""Per-kernel CUPTI breakdown of VQModel.encode / decode_code (torch.profiler), top kernels by total time.
  python tools/trace_encode.py [--batch 64] [--size 256] [--what encode|decode]"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--what", default="encode")
    args = ap.parse_args()
    from llamagen_b200 import VQ_models
    torch.manual_seed(0)
    m = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).cuda().eval()
    g = args.size // 16
    x = torch.rand(args.batch, 3, args.size, args.size, device="cuda") * 2 - 1
    idx = torch.randint(0, 16384, (args.batch, g * g), device="cuda")
    fn = (lambda: m.encode(x)) if args.what == "encode" else (lambda: m.decode_code(idx, [args.batch, 8, g, g]))
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
            dur = getattr(ev, "device_time_total", 0) or getattr(ev, "cuda_time_total", 0)
            agg[ev.name[:90]][0] += dur
            agg[ev.name[:90]][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"{args.what} B={args.batch} {args.size}px: kernel time {total / 1e3:.2f} ms")
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"{t / 1e3:9.3f} ms {n:5d}x {100 * t / total:5.1f}%  {name}")


if __name__ == "__main__":
    main()
