"""Mirror of the image-processor call the reference's dataset makes on decoded video frames
(`self.video_processor.preprocess(frames, return_tensors='pt')['pixel_values']`, dataset/quick_start_dataset.py:303-315,
where `video_processor` is the HF `CLIPImageProcessor` of the vision tower, scripts/quick_start.py:560).

The reference decodes video frames at `image_size x image_size` (VideoReader(height=224, width=224)), so for them the
processor's shortest-edge resize and centre crop are identities and what remains is rescale (1/255) and normalisation with
the OpenAI CLIP statistics.  Here that arithmetic runs on the GPU, either to the reference's fp32 `pixel_values`
(`preprocess`) or — when the uint8 frames are handed to the engine directly — fused into the patch-embed im2col so the
fp32 tensor never exists (`CrabEngine.clip_forward` on a uint8 input).

Inputs of any other size (the `<image>` tasks open arbitrary files) go through the processor's resize first.  The
reference's pinned transformers (4.37.2) implements it as `PIL.Image.resize((w, h), resample=BICUBIC)` followed by a centre
crop; Pillow's 8-bit resampler is integer arithmetic (22-bit fixed-point coefficients, one horizontal and one vertical pass,
each rounded to uint8), so it is reproduced BIT-EXACTLY: the coefficient tables below follow Pillow's `precompute_coeffs` /
`normalize_coeffs_8bpc` (src/libImaging/Resample.c) in float64, the two passes run in `crab_resample_u8`, and the crop is folded
into the passes (only the kept 224 x 224 window is computed).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import ops
from ..lib import CrabError

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


PRECISION_BITS = 32 - 8 - 2  # Pillow: 8-bit pixels, 2 guard bits


def _bicubic(x: np.ndarray) -> np.ndarray:
    """Pillow's bicubic kernel (a = -0.5), float64."""
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def pil_bicubic_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """(bounds int32 [out, 2] = first tap / tap count, kk int32 [out, ksize] fixed-point coefficients) for resampling
    `in_size` -> `out_size` samples, as Pillow's precompute_coeffs + normalize_coeffs_8bpc build them."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    centers = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((centers - support + 0.5).astype(np.int64), 0)         # C (int) cast truncates; values are >= -0.x
    xmax = np.minimum((centers + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    taps = np.arange(ksize, dtype=np.float64)[None, :]
    w = _bicubic((taps + xmin[:, None] - centers[:, None] + 0.5) * (1.0 / fscale))
    w = np.where(taps < cnt[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for t in range(ksize):  # Pillow accumulates the normaliser tap by tap, left to right
        ww = ww + w[:, t]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    fixed = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
    kk = np.trunc(fixed).astype(np.int32)
    kk[taps.repeat(out_size, 0) >= cnt[:, None]] = 0
    return np.stack([xmin, cnt], 1).astype(np.int32), kk


def resize_output_size(h: int, w: int, shortest_edge: int) -> Tuple[int, int]:
    """(new_h, new_w) of HF's shortest-edge resize (image_transforms.get_resize_output_image_size, default_to_square=False)."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = shortest_edge, int(shortest_edge * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def frames_to_uint8_thwc(frames) -> torch.Tensor:
    """list of PIL images / HWC uint8 arrays, or a (t, H, W, 3) uint8 array/tensor -> uint8 tensor (t, H, W, 3)."""
    import numpy as np

    if isinstance(frames, torch.Tensor):
        t = frames
    elif isinstance(frames, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(frames))
    else:
        t = torch.stack([torch.from_numpy(np.ascontiguousarray(np.asarray(f.convert("RGB") if hasattr(f, "convert") else f)))
                         for f in frames], 0)
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
        raise CrabError(f"expected uint8 frames (t, H, W, 3), got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


class ClipImageProcessorB200:
    """`preprocess(frames, return_tensors='pt') -> {'pixel_values': fp32 (t, 3, 224, 224)}` on the device."""

    def __init__(self, image_size: int = 224, image_mean: Sequence[float] = OPENAI_CLIP_MEAN,
                 image_std: Sequence[float] = OPENAI_CLIP_STD, rescale_factor: float = 1.0 / 255.0,
                 device: Optional[torch.device] = None):
        self.crop_size = {"height": image_size, "width": image_size}
        self.size = {"shortest_edge": image_size}
        self.image_mean, self.image_std, self.rescale_factor = tuple(image_mean), tuple(image_std), rescale_factor
        self.device = device
        self._coeffs = {}

    def _tables(self, in_size: int, out_size: int, dev: torch.device):
        key = (in_size, out_size, dev.index or 0)
        if key not in self._coeffs:
            b, k = pil_bicubic_tables(in_size, out_size)
            self._coeffs[key] = (torch.from_numpy(b).to(dev), torch.from_numpy(k).to(dev))
        return self._coeffs[key]

    def to_device_uint8(self, frames) -> torch.Tensor:
        """-> uint8 (t, S, S, 3) on the device, S = crop size: as-is for S x S inputs, otherwise Pillow-exact bicubic
        shortest-edge resize + centre crop (two `crab_resample_u8` passes computing only the kept window)."""
        t = frames_to_uint8_thwc(frames)
        s = self.crop_size["height"]
        dev = self.device or (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        if dev.type != "cuda":
            raise CrabError("ClipImageProcessorB200 runs on the GPU only (no CPU path)")
        t = t.to(dev, non_blocking=True)
        h, w = int(t.shape[1]), int(t.shape[2])
        if (h, w) == (s, s):
            return t
        nh, nw = resize_output_size(h, w, self.size["shortest_edge"])
        if nh < s or nw < s:
            raise CrabError(f"resized image {nh}x{nw} is smaller than the {s}x{s} crop")
        top, left = (nh - s) // 2, (nw - s) // 2
        if nw != w:      # horizontal pass first, as Pillow does; only the kept columns are computed
            bx, kx = self._tables(w, nw, dev)
            t = ops.resample_u8(t, s, 1, bx, kx, left)
        elif w != s:     # pure crop along x: Pillow skips the pass, so do we
            t = t[:, :, left:left + s].contiguous()
        if nh != h:
            by, ky = self._tables(h, nh, dev)
            t = ops.resample_u8(t, s, 0, by, ky, top)
        elif h != s:
            t = t[:, top:top + s].contiguous()
        return t

    def preprocess(self, frames, return_tensors: str = "pt") -> Dict[str, torch.Tensor]:
        u8 = self.to_device_uint8(frames)
        return {"pixel_values": ops.normalize_u8(u8, self.image_mean, self.image_std, self.rescale_factor)}

    __call__ = preprocess
