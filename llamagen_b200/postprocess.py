"""Pixel finishing of the samplers: fp32 NCHW decoder output -> uint8 NHWC, on the GPU in one kernel.

Replaces, in `autoregressive/sample/sample_c2i_ddp.py:141-148`,
    samples = F.interpolate(samples, size=(eval, eval), mode='bicubic')           # only when image_size_eval differs
    samples = torch.clamp(127.5 * samples + 128.0, 0, 255).permute(0, 2, 3, 1).to("cpu", dtype=torch.uint8).numpy()
    Image.fromarray(sample).save(...)                                             # one PNG per image, serial on the host
with `lg_pixels_to_u8` (resize + affine + clamp + convert + NHWC in one pass) and a thread pool of PNG encoders that runs
while the GPU samples the next batch."""
from __future__ import annotations

import ctypes
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib


@torch.no_grad()
def to_uint8_nhwc(samples: torch.Tensor, size=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """samples: CUDA fp32 NCHW [B,C,H,W]. Returns CUDA uint8 NHWC [B,size,size,C] (size=None keeps H x W)."""
    _lib.require_cuda(samples, "to_uint8_nhwc")
    if samples.dtype != torch.float32 or samples.dim() != 4:
        raise ValueError(f"to_uint8_nhwc expects fp32 NCHW, got {samples.dtype} {tuple(samples.shape)}")
    x = samples.contiguous()
    B, C, H, W = x.shape
    oh, ow = (H, W) if size is None else ((size, size) if isinstance(size, int) else tuple(size))
    if out is None:
        out = torch.empty(B, oh, ow, C, dtype=torch.uint8, device=x.device)
    elif tuple(out.shape) != (B, oh, ow, C) or out.dtype != torch.uint8 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("to_uint8_nhwc: `out` must be a contiguous CUDA uint8 [B,H,W,C] tensor on the input's device")
    _lib.check(_lib.load().lg_pixels_to_u8(_lib.ptr(x), B, C, H, W, oh, ow, _lib.ptr(out), _lib.current_stream(x.device)),
               "lg_pixels_to_u8")
    return out


class AsyncPngWriter:
    """Host side of sample_c2i_ddp.py:145-148: PNG-encode uint8 HWC arrays on worker threads (zlib releases the GIL) so the
    sampler's GPU loop is not serialised behind the encoder. `close()` waits for every file and re-raises the first error."""

    def __init__(self, workers: int = 8):
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._futures = []

    @staticmethod
    def _save(arr, path):
        from PIL import Image
        Image.fromarray(arr).save(path)

    def submit(self, arr, path):
        self._futures.append(self._pool.submit(self._save, arr, path))

    def drain(self):
        futures, self._futures = self._futures, []
        for f in futures:
            f.result()

    def close(self):
        try:
            self.drain()
        finally:
            self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
