"""Drop-in for autoregressive/models/generate.py: same `generate()` signature and return value
(int32 [B, max_new_tokens] on cond.device, generate.py:126-176), executed by ONE C-ABI call: prefill,
the S-1 KV-cached decode steps, CFG mixing and top-k/top-p sampling all stay on the device
(lg_generate in include/llamagen_b200.h); the reference's Python loop of S host iterations is gone.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .gpt import Transformer


def _draw_seed() -> int:
    # torch.multinomial consumes torch's global generator in the reference (generate.py:63); we take one
    # 62-bit draw from the same generator so `torch.manual_seed(s)` keeps runs reproducible.
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def sample(logits, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0, sample_logits=True, seed=None):
    """generate.py:57-66 on the device: logits [B, T, V] (last position used) -> (idx int64 [B,1], probs [B,V])."""
    lib = _lib.load()
    x = _lib.require_cuda(logits, "sample")[:, -1, :].float().contiguous()
    B, V = x.shape
    idx = torch.empty(B, dtype=torch.int32, device=x.device)
    probs = torch.empty(B, V, dtype=torch.float32, device=x.device)
    sc = _lib.SampleCfg(1.0, -1, float(temperature), int(top_k), float(top_p), 0 if sample_logits else 1,
                        _draw_seed() if seed is None else int(seed))
    _lib.check(lib.lg_sample(_lib.ptr(x), B, V, 0, _lib.LG_DTYPE_F32, ctypes.byref(sc), 0, _lib.ptr(idx),
                             _lib.ptr(probs), _lib.current_stream(x.device)), "lg_sample")
    return idx.long().unsqueeze(-1), probs


@torch.no_grad()
def generate(model, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1, **sampling_kwargs):
    """Same contract as the reference generate() (generate.py:126-176).

    Extra keyword-only knobs (ignored by the reference signature, all optional):
      seed=int             explicit RNG seed for the multinomial draw
      return_logits=True   also return the per-step CFG-mixed logits [S, B, V] (fp32) for parity tests
      teacher=int tensor   [B, S] tokens fed back instead of the sampled ones (teacher forcing)
    """
    if not isinstance(model, Transformer):
        raise TypeError("llamagen_b200.generate needs a llamagen_b200 GPT_models[...] instance")
    temperature = float(sampling_kwargs.pop("temperature", 1.0))
    top_k = int(sampling_kwargs.pop("top_k", 0) or 0)
    top_p = float(sampling_kwargs.pop("top_p", 1.0))
    sample_logits = bool(sampling_kwargs.pop("sample_logits", True))
    seed = sampling_kwargs.pop("seed", None)
    return_logits = bool(sampling_kwargs.pop("return_logits", False))
    teacher = sampling_kwargs.pop("teacher", None)
    if sampling_kwargs:
        raise TypeError(f"unexpected sampling kwargs: {sorted(sampling_kwargs)}")

    dev = model.tok_embeddings.weight.device
    _lib.require_cuda(model.tok_embeddings.weight, "generate")
    if model.model_type == "c2i":
        T = 1
        cond_dev = cond.to(device=dev, dtype=torch.int32).contiguous()
        if cond_dev.dim() != 1:
            raise ValueError("c2i cond must be a 1-D tensor of class labels")
    elif model.model_type == "t2i":
        T = cond.shape[1]
        cond_dev = cond.to(device=dev, dtype=model.tok_embeddings.weight.dtype).contiguous()
    else:
        raise Exception("please check model type")            # generate.py:143
    B = cond.shape[0]
    S = int(max_new_tokens)
    use_cfg = cfg_scale > 1.0
    rows = 2 * B if use_cfg else B

    em = None
    if emb_masks is not None:
        assert emb_masks.shape[0] == B                          # generate.py:155-156
        assert emb_masks.shape[-1] == T
        em = emb_masks.to(device=dev, dtype=torch.float32).contiguous()

    model.setup_caches(max_batch_size=rows, max_seq_length=T + S, dtype=model.tok_embeddings.weight.dtype)
    lib, handle = _lib.load(), model.engine()
    out = torch.empty((B, S), dtype=torch.int32, device=dev)
    dbg = torch.empty((S, B, model.vocab_size), dtype=torch.float32, device=dev) if return_logits else None
    tf = teacher.to(device=dev, dtype=torch.int32).contiguous() if teacher is not None else None
    if tf is not None and tuple(tf.shape) != (B, S):
        raise ValueError("teacher must be [B, max_new_tokens]")
    sc = _lib.SampleCfg(float(cfg_scale), int(cfg_interval), temperature, top_k, top_p, 0 if sample_logits else 1,
                        _draw_seed() if seed is None else int(seed))
    _lib.check(lib.lg_generate(handle, _lib.ptr(cond_dev), _lib.ptr(em), B, T, S, ctypes.byref(sc), _lib.ptr(out),
                               _lib.ptr(dbg), _lib.ptr(tf), _lib.current_stream(dev)), "lg_generate")
    if cond.device != dev:
        out = out.to(cond.device)
    return (out, dbg) if return_logits else out
