"""TEST INFRASTRUCTURE — deterministic synthetic weights and inputs.

The reference's checkpoints are not available offline and its constructors' random inits cannot travel to the GPU
box, so parity cases use weights regenerated from (key name, shape, seed) alone: the golden generator loads them
into the reference's modules, the tests load the same tensors into the oracle and into crab_b200.
Magnitudes are chosen so activations stay O(1) through the stack (SURVEY.md §7 "realistic magnitudes").
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Tuple

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return torch.Generator(device="cpu").manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)


def synth_tensor(key: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    if "relative_attention_bias" in key:
        # one Parameter shared by all BEATs layers (backbone.py:78-81) but listed once per layer in the state dict
        import re

        key = re.sub(r"layers\.\d+\.", "layers.0.", key)
    g = _gen(key, seed)
    shape = tuple(shape)
    last = key.rsplit(".", 1)[-1]
    n = 1
    for s in shape:
        n *= s
    r = torch.randn(shape, generator=g, dtype=torch.float32) if n > 0 else torch.zeros(shape)
    if key.endswith("position_ids"):
        return torch.arange(shape[-1]).expand(shape).clone()
    if "lora_B" in key:
        return 0.05 * r  # the reference zero-inits lora_B (lora.py:305); non-zero so the side path is exercised
    if "lora_A" in key or "lora_route" in key:
        return r / math.sqrt(shape[-1])
    if last == "weight_g":
        return 0.5 + 0.1 * r.abs()
    if last == "grep_a":
        return 1.0 + 0.2 * r
    if "relative_attention_bias" in key:
        return 0.5 * r
    if "query_tokens" in key:
        return 0.5 * r
    if "class_embedding" in key:
        return 0.5 * r
    if "embed_tokens" in key or "position_embedding" in key or "word_embeddings" in key or "position_embeddings" in key:
        return 0.5 * r if "embed_tokens" not in key else r
    if len(shape) == 1:
        if last == "bias":
            return 0.05 * r
        return 1.0 + 0.1 * r  # norm scales
    fan_in = n // shape[0]
    return r / math.sqrt(fan_in)


def synth_state_dict(manifest: Dict[str, Tuple[int, ...]], seed: int) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, tuple(shp), seed) for k, shp in manifest.items()}


def synth_inputs(seed: int, *, frames: int, image: int, audio_segs: int, audio_len: int, prompt_len: int,
                 base_vocab: int, video_id: int, audio_id: int, video_at: int = 10, audio_at: int = 20):
    """One sample: video (frames,3,image,image) ~ N(0,1), fbank (segs, audio_len, 128) ~ 0.5 N(0,1), prompt ids with
    one <video> and one <audio> placeholder."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    video = torch.randn(frames, 3, image, image, generator=g)
    audio = 0.5 * torch.randn(audio_segs, audio_len, 128, generator=g)
    ids = torch.randint(3, base_vocab, (prompt_len,), generator=g)
    ids[video_at] = video_id
    ids[audio_at] = audio_id
    return video, audio, ids
