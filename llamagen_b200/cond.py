"""Text-condition front end of the t2i samplers: T5 feature packing and left-padding on the device.

Replaces the per-prompt Python loop of `autoregressive/sample/sample_t2i.py:92-106` / `sample_t2i_ddp.py:140-156`
(one `.item()` sync and one `torch.cat` per prompt) with batched tensor ops, and reads the feature files written by
`language/extract_t5_feature.py:103-108` (fp32 `[1, valid_len, 2048]` .npy, one per caption). The Flan-T5 encoder itself
is upstream of the path and stays HuggingFace."""
from __future__ import annotations

import numpy as np
import torch


def pack_t5_features(features, max_len: int = 120, dim: int = 2048):
    """features: iterable of arrays/tensors shaped [valid_len, dim] or [1, valid_len, dim] (the .npy layout). Returns
    (embs fp32 [B,max_len,dim] right-padded with zeros, masks fp32 [B,max_len]) — the layout T5Embedder.get_text_embeddings
    hands to the sampler (language/t5.py), truncated at max_len like `dataset/t2i.py:117-118`."""
    feats = [torch.as_tensor(np.asarray(f) if not torch.is_tensor(f) else f).reshape(-1, dim)[:max_len].float() for f in features]
    embs = torch.zeros(len(feats), max_len, dim)
    masks = torch.zeros(len(feats), max_len)
    for i, f in enumerate(feats):
        embs[i, : f.shape[0]] = f
        masks[i, : f.shape[0]] = 1
    return embs, masks


def load_t5_feature_files(paths, max_len: int = 120, dim: int = 2048):
    return pack_t5_features([np.load(p) for p in paths], max_len, dim)


def left_pad_features(caption_embs: torch.Tensor, emb_masks: torch.Tensor):
    """sample_t2i.py:92-103 for the whole batch at once: row i is rotated left by valid_i (= cat([emb[valid:], emb[:valid]]))
    so the valid tokens sit at the right end, and the mask is flipped. No host sync. Returns (embs, masks)."""
    B, T = emb_masks.shape
    valid = emb_masks.sum(dim=-1).to(torch.long)                                  # [B]
    src = (torch.arange(T, device=emb_masks.device)[None, :] + valid[:, None]) % T
    rolled = torch.gather(caption_embs, 1, src[:, :, None].expand(B, T, caption_embs.shape[-1]))
    return rolled, torch.flip(emb_masks, dims=[-1])


def prepare_condition(caption_embs: torch.Tensor, emb_masks: torch.Tensor, left_padding: bool = True):
    """The (c_indices, c_emb_masks) pair the samplers pass to generate() (sample_t2i.py:104-106)."""
    if left_padding:
        caption_embs, emb_masks = left_pad_features(caption_embs, emb_masks)
    return caption_embs * emb_masks[:, :, None].to(caption_embs.dtype), emb_masks


def synthetic_features(batch: int, max_len: int, dim: int, seed: int, device, dtype):
    """Seeded random T5-shaped features with ragged valid lengths (benchmarks / tests: no T5 weights offline)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(8, max_len, (batch,), generator=g)
    masks = (torch.arange(max_len)[None, :] < lens[:, None]).float()
    return (torch.randn(batch, max_len, dim, generator=g) * masks[:, :, None]).to(device, dtype), masks.to(device)


class HFT5Encoder:
    """The T5Embedder.get_text_embeddings call of language/t5.py as the samplers use it: tokenizer padded to max_len,
    encoder last_hidden_state, attention mask. Plain HuggingFace — upstream of the path, not reimplemented."""

    def __init__(self, t5_path, model_type, max_len, device, dtype):
        import os
        from transformers import AutoTokenizer, T5EncoderModel
        path = os.path.join(t5_path, model_type)
        self.tok = AutoTokenizer.from_pretrained(path)
        self.enc = T5EncoderModel.from_pretrained(path, torch_dtype=dtype).to(device).eval()
        self.max_len, self.device = max_len, device

    @torch.no_grad()
    def __call__(self, prompts):
        t = self.tok(list(prompts), max_length=self.max_len, padding="max_length", truncation=True, return_attention_mask=True,
                     add_special_tokens=True, return_tensors="pt")
        ids, mask = t["input_ids"].to(self.device), t["attention_mask"].to(self.device)
        return self.enc(input_ids=ids, attention_mask=mask)["last_hidden_state"].detach(), mask
