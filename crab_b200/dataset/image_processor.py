"""Mirror of the image-processor call the reference's dataset makes on decoded video frames
(`self.video_processor.preprocess(frames, return_tensors='pt')['pixel_values']`, dataset/quick_start_dataset.py:303-315,
where `video_processor` is the HF `CLIPImageProcessor` of the vision tower, scripts/quick_start.py:560).

The reference decodes every frame at `image_size x image_size` (VideoReader(height=224, width=224)), so the processor's
shortest-edge resize and centre crop are identities on this path; what remains is rescale (1/255) and normalisation with
the OpenAI CLIP statistics.  Here that arithmetic runs on the GPU, either to the reference's fp32 `pixel_values`
(`preprocess`) or — when the uint8 frames are handed to the engine directly — fused into the patch-embed im2col so the
fp32 tensor never exists (`CrabEngine.clip_forward` on a uint8 input).  Frames of another size raise: resizing belongs to
the decoder stage here, exactly as in the reference dataset.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

import torch

from .. import ops
from ..lib import CrabError

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


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

    def to_device_uint8(self, frames) -> torch.Tensor:
        t = frames_to_uint8_thwc(frames)
        s = self.crop_size["height"]
        if t.shape[1] != s or t.shape[2] != s:
            raise CrabError(f"frames are {t.shape[1]}x{t.shape[2]}; decode them at {s}x{s} as the reference dataset does "
                            "(VideoReader(height=image_size, width=image_size)) — the GPU front-end does not resize")
        dev = self.device or (t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        if dev.type != "cuda":
            raise CrabError("ClipImageProcessorB200 runs on the GPU only (no CPU path)")
        return t.to(dev, non_blocking=True)

    def preprocess(self, frames, return_tensors: str = "pt") -> Dict[str, torch.Tensor]:
        u8 = self.to_device_uint8(frames)
        return {"pixel_values": ops.normalize_u8(u8, self.image_mean, self.image_std, self.rescale_factor)}

    __call__ = preprocess
