"""Mirror of the reference's `models/unified_qwen.py` (Qwen2 backbone): same surface as unified_llama, with the
Qwen2 differences the engine needs — q/k/v projections carry a bias (models/qwen/modeling_qwen2.py:234-237), grouped
KV heads (`repeat_kv`, :190-199) and rope_theta / vocab from the checkpoint config.

The reference's multimodal entry for Qwen is stale (it passes keyword arguments `prepare_multimodal_inputs` does not
accept — SURVEY.md §2.1); here `generate(batch_*)` uses the same working path as the LLaMA wrapper.
"""
from __future__ import annotations

from transformers import Qwen2Config

from .unified_llama import UnifiedForCausalLM as _LlamaWrapper
from .unified_llama import UnifiedModel as _LlamaModel


class UnifiedConfig(Qwen2Config):
    model_type = "unified_llm_qwen"


class UnifiedModel(_LlamaModel):
    config_class = UnifiedConfig
    qkv_bias = True


class UnifiedForCausalLM(_LlamaWrapper):
    config_class = UnifiedConfig
    _model_cls = UnifiedModel
