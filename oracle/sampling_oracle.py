"""Oracle for CFG mixing + top-k/top-p filtering + sampling (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows autoregressive/models/generate.py: top_k_top_p_filtering :16-54, sample :57-66, CFG mix :81-82,95-99.
"""
from __future__ import annotations

import torch


def cfg_mix_oracle(logits_2b: torch.Tensor, cfg_scale: float) -> torch.Tensor:
    """generate.py:95-97 — rows [0,B) are conditional, [B,2B) unconditional."""
    half = logits_2b.shape[0] // 2
    cond, uncond = logits_2b[:half], logits_2b[half:]
    return uncond + (cond - uncond) * cfg_scale


def top_k_top_p_oracle(logits: torch.Tensor, top_k: int = 0, top_p: float = 1.0) -> torch.Tensor:
    """generate.py:16-54 on a [B, V] tensor (returns a filtered copy; -inf marks removed tokens)."""
    x = logits.clone()
    neg = -float("inf")
    if top_k > 0:
        k = min(max(top_k, 1), x.size(-1))                                   # :32
        kth = torch.topk(x, k).values[..., -1:]                              # :35 value of the k-th largest
        x = torch.where(x < kth, torch.full_like(x, neg), x)                 # ties with kth survive
    if top_p < 1.0:
        order = torch.sort(x, descending=True)                               # :39
        cum = torch.cumsum(torch.softmax(order.values, dim=-1), dim=-1)      # :40
        drop_sorted = cum > top_p                                            # :43
        drop_sorted = torch.cat([torch.zeros_like(drop_sorted[..., :1]), drop_sorted[..., :-1]], dim=-1)  # :48-49
        drop = torch.zeros_like(drop_sorted).scatter(1, order.indices, drop_sorted)                        # :52
        x = torch.where(drop, torch.full_like(x, neg), x)
    return x


def sample_oracle(logits_last: torch.Tensor, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0,
                  sample_logits: bool = True, generator=None):
    """generate.py:57-66 given the last-position logits [B, V]; returns (idx [B,1] int64, probs [B,V])."""
    x = logits_last / max(temperature, 1e-5)
    if top_k > 0 or top_p < 1.0:
        x = top_k_top_p_oracle(x, top_k=top_k, top_p=top_p)
    probs = torch.softmax(x, dim=-1)
    if sample_logits:
        idx = torch.multinomial(probs, num_samples=1, generator=generator)
    else:
        idx = torch.topk(probs, k=1, dim=-1).indices
    return idx, probs
