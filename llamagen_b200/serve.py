"""Request-queue front end for class-conditional sampling with the call surface of the reference's
vLLM-fork entry point: `LLM.generate(prompt_token_ids=..., sampling_params=...)` (autoregressive/serve/llm.py:138-221,
used by autoregressive/serve/sample_c2i.py:39-63) returning `RequestOutput`-shaped objects sorted by request id (:266).

Every image request has the same length (cls_token_num condition tokens + S image tokens), so iteration-level scheduling
degenerates to packing waiting requests into engine batches: `step()` takes up to `max_num_seqs` waiting requests that
share their sampling parameters and runs them through `generate()` (one CUDA-graph decode loop); requests queued while a
batch runs are served by the next `step()`. Classifier-free guidance follows the reference's serving protocol
(serve/sample_c2i.py:35-37, serve/sampler.py:54-58): the caller appends one `[num_classes]` (null-class) prompt per
conditional prompt; the second half is recognised as the unconditional twins, the pair is sampled once, and both
requests receive the same tokens (the reference samples the two copies of the mixed logits independently and the caller
discards the second half, serve/sample_c2i.py:66-67)."""
from __future__ import annotations

import itertools
from collections import deque
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from .generate import generate


@dataclass(frozen=True)
class SamplingParams:
    """The vllm.SamplingParams fields the reference sets (serve/sample_c2i.py:47-49). top_k=-1 / 0 disables top-k."""
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    max_tokens: int = 256


@dataclass
class CompletionOutput:
    index: int
    token_ids: List[int]
    finish_reason: str = "length"


@dataclass
class RequestOutput:
    request_id: str
    prompt_token_ids: List[int]
    outputs: List[CompletionOutput] = field(default_factory=list)
    finished: bool = False


@dataclass
class _Request:
    request_id: str
    prompt_token_ids: List[int]
    params: SamplingParams
    twin_of: Optional[str] = None      # request id of the conditional request this null-class request mirrors


class LLM:
    def __init__(self, gpt_model, cfg_scale: float = 1.0, num_classes: int = 1000, max_num_seqs: int = 64, seed: Optional[int] = None):
        if gpt_model.model_type != "c2i":
            raise ValueError("serve.LLM handles class-conditional models (the reference's serve path is c2i only)")
        self.model = gpt_model
        self.cfg_scale = float(cfg_scale)
        self.num_classes = int(num_classes)
        self.max_num_seqs = int(max_num_seqs)
        self.request_counter = itertools.count()
        self._waiting: deque[_Request] = deque()
        self._twins: dict[str, _Request] = {}
        self._seed = seed
        self._batches = 0

    # ------------------------------------------------------------------ engine-style surface (llm_engine.py)
    def add_request(self, prompt_token_ids: Sequence[int], sampling_params: Optional[SamplingParams] = None, twin_of=None) -> str:
        if len(prompt_token_ids) != 1 or not (0 <= int(prompt_token_ids[0]) <= self.num_classes):
            raise ValueError(f"a c2i prompt is one class id in [0, {self.num_classes}], got {list(prompt_token_ids)}")
        rid = str(next(self.request_counter))
        req = _Request(rid, [int(prompt_token_ids[0])], sampling_params or SamplingParams(), twin_of)
        if twin_of is None:
            self._waiting.append(req)
        else:
            self._twins[twin_of] = req
        return rid

    def has_unfinished_requests(self) -> bool:
        return bool(self._waiting)

    def get_num_unfinished_requests(self) -> int:
        return len(self._waiting) + len(self._twins)

    @torch.no_grad()
    def step(self) -> List[RequestOutput]:
        """Run one engine batch: the longest prefix run of waiting requests with identical sampling parameters, capped at
        max_num_seqs. Returns the finished RequestOutputs (conditional requests and their null-class twins)."""
        if not self._waiting:
            return []
        params = self._waiting[0].params
        batch = []
        while self._waiting and len(batch) < self.max_num_seqs and self._waiting[0].params == params:
            batch.append(self._waiting.popleft())
        dev = self.model.tok_embeddings.weight.device
        cond = torch.tensor([r.prompt_token_ids[0] for r in batch], dtype=torch.int32, device=dev)
        kw = {}
        if self._seed is not None:
            kw["seed"] = self._seed + self._batches
        self._batches += 1
        tokens = generate(self.model, cond, params.max_tokens, cfg_scale=self.cfg_scale, temperature=params.temperature,
                          top_k=max(0, params.top_k), top_p=params.top_p, sample_logits=True, **kw).cpu().tolist()
        outs = []
        for req, toks in zip(batch, tokens):
            outs.append(RequestOutput(req.request_id, req.prompt_token_ids, [CompletionOutput(0, toks)], True))
            twin = self._twins.pop(req.request_id, None)
            if twin is not None:
                outs.append(RequestOutput(twin.request_id, twin.prompt_token_ids, [CompletionOutput(0, list(toks))], True))
        return outs

    # ------------------------------------------------------------------ LLM.generate (llm.py:138-266)
    def generate(self, prompts=None, sampling_params=None, prompt_token_ids: Optional[List[List[int]]] = None,
                 use_tqdm: bool = False) -> List[RequestOutput]:
        if prompts is not None:
            raise ValueError("prompts must be None: the image models have no tokenizer (skip_tokenizer_init=True)")
        if prompt_token_ids is None:
            raise ValueError("Either prompts or prompt_token_ids must be provided.")
        n = len(prompt_token_ids)
        if isinstance(sampling_params, list) and len(sampling_params) != n:
            raise ValueError("The lengths of prompts and sampling_params must be the same.")
        per = sampling_params if isinstance(sampling_params, list) else [sampling_params or SamplingParams()] * n
        half = n // 2
        paired = (self.cfg_scale > 1.0 and n % 2 == 0 and n > 0 and
                  all(list(p) == [self.num_classes] for p in prompt_token_ids[half:]) and
                  all(per[i] == per[i + half] for i in range(half)))
        if self.cfg_scale > 1.0 and not paired:
            raise ValueError("cfg_scale > 1 expects the serving protocol of serve/sample_c2i.py:35-37: the conditional prompts "
                             "followed by one [num_classes] prompt each")
        ids = [self.add_request(prompt_token_ids[i], per[i]) for i in range(half if paired else n)]
        if paired:
            for i in range(half):
                self.add_request(prompt_token_ids[half + i], per[half + i], twin_of=ids[i])
        outputs: List[RequestOutput] = []
        while self.has_unfinished_requests():
            outputs.extend(o for o in self.step() if o.finished)
        return sorted(outputs, key=lambda o: int(o.request_id))
