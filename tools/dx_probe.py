"""Phase timing of the direct-epilogue GEMM (gemm_dx.cu) through its %globaltimer stamps, stand-alone calls at decode shapes.
Usage: python tools/dx_probe.py [rows]   (env LG_DX_RBLK / LG_DX_STAGES as for the engine)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from llamagen_b200 import _lib
from util import gemm_dx
lib = _lib.load()
lib.lg_debug_set_dx_trace.argtypes = [ctypes.c_void_p]
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, K, mode, norm, tag) in ((1024, 1024, 1, False, "wo' resid"), (2816, 1024, 2, True, "w13' norm+swiglu"), (1024, 1024, 0, False, "plain f32"),
                                (1024, 1024, 0, True, "plain f32 + norm")):
    x = (torch.randn(R, K, device="cuda") * 0.5).bfloat16()
    wa = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    wb = (torch.randn(N, K, device="cuda") * 0.05).bfloat16() if mode == 2 else None
    g = torch.ones(K, device="cuda").bfloat16() if norm else None
    h = torch.randn(R, N, device="cuda").bfloat16() if mode == 1 else None
    for cold in (False, True):
        for _ in range(3):
            gemm_dx(x, wa, wb, mode=mode, normw=g, h=h)
        if cold:
            flush.fill_(1)
        trace = torch.zeros(4096 * 12, dtype=torch.int64, device="cuda")
        lib.lg_debug_set_dx_trace(ctypes.c_void_p(trace.data_ptr()))
        torch.cuda.synchronize()
        gemm_dx(x, wa, wb, mode=mode, normw=g, h=h)
        torch.cuda.synchronize()
        lib.lg_debug_set_dx_trace(ctypes.c_void_p(0))
        t = trace.view(-1, 12).cpu()
        t = t[t[:, 0] > 0].double()
        g0 = t[:, 0].min()
        names = ["entry->setup", "setup->preloads_issued", "pdl_wait", "x_issued->all_w_issued(3->4)", "x_issued->xfull(3->5)", "xfull/norm->mma_start(5|3->6)",
                 "mma_start->last_full(6->7)", "last_full->tmem_full(7->8)", "tmem_full->drained(8->9)", "drained->exit(9->10)", "total(0->10)"]
        def dd(a, b):
            v = (t[:, b] - t[:, a])
            ok = (t[:, a] > 0) & (t[:, b] > 0)
            v = v[ok]
            return f"{float(v.mean()):.0f}/{float(v.max()):.0f}" if v.numel() else "-"
        vals = [dd(0, 1), dd(1, 2), dd(2, 3), dd(3, 4), dd(3, 5), dd(5, 6) if norm else dd(3, 6), dd(6, 7), dd(7, 8), dd(8, 9), dd(9, 10), dd(0, 10)]
        print(f"{tag} N={N} K={K} R={R} ctas={t.shape[0]} cold_l2={cold} rblk={os.environ.get('LG_DX_RBLK')} stages={os.environ.get('LG_DX_STAGES')} "
              f"kernel_span_ns={float(t[:, 10].max() - g0):.0f}")
        print("   " + "  ".join(f"{k}: {v}" for k, v in zip(names, vals)))
