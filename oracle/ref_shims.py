"""TEST INFRASTRUCTURE — import the *real* reference (`/root/reference`) in the build container.

The reference pins transformers 4.37.2 and needs packages that are absent here; SURVEY.md §8(c) lists the shims
that let its own modules run unmodified on the installed transformers 5.5 / torch 2.11.  Nothing in this file is
used on the GPU box (the reference checkout does not exist there): it exists only to (a) pin the restatement in
`oracle/crab_oracle.py` against the reference's actual code and (b) generate `tests/golden/*.pt`.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types
from pathlib import Path

import torch

REF_ROOT = Path(os.environ.get("CRAB_REFERENCE_ROOT", "/root/reference"))


def reference_available() -> bool:
    return (REF_ROOT / "models" / "unified_arch.py").exists()


_installed = False


def install_shims() -> None:
    """Monkey-patch the import environment so `models.*` / `peft_hyper` of the reference import cleanly."""
    global _installed
    if _installed:
        return
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    # (1) helpers that moved out of modeling_utils after 4.x  (models/Qformer.py:40-45)
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = pu.prune_linear_layer
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def _fphi(*a, **k):
            raise NotImplementedError("head pruning is not on the hot path")
        mu.find_pruneable_heads_and_indices = _fphi

    # (2) models/unified_arch.py:13 imports a package that is not in the repo
    pkg = types.ModuleType("models.video_llama2")
    pkg.__path__ = []
    proj = types.ModuleType("models.video_llama2.projector")

    class STCConnectorV35(torch.nn.Module):  # never constructed on this path (model == 'unified')
        def __init__(self, *a, **k):
            super().__init__()

    proj.STCConnectorV35 = STCConnectorV35
    sys.modules.setdefault("models.video_llama2", pkg)
    sys.modules.setdefault("models.video_llama2.projector", proj)

    # (5) accelerate is import-only in peft_hyper/peft_model.py:24-26
    if "accelerate" not in sys.modules:
        try:
            import accelerate  # noqa: F401
        except Exception:
            acc = types.ModuleType("accelerate")
            acc.__path__ = []
            acc.dispatch_model = lambda m, *a, **k: m
            acc.infer_auto_device_map = lambda *a, **k: {}
            hooks = types.ModuleType("accelerate.hooks")

            class AlignDevicesHook:  # noqa: D401
                pass

            hooks.AlignDevicesHook = AlignDevicesHook
            hooks.add_hook_to_module = lambda *a, **k: None
            hooks.remove_hook_from_submodules = lambda *a, **k: None
            utils = types.ModuleType("accelerate.utils")
            utils.get_balanced_memory = lambda *a, **k: {}
            sys.modules["accelerate"] = acc
            sys.modules["accelerate.hooks"] = hooks
            sys.modules["accelerate.utils"] = utils
            acc.hooks = hooks
            acc.utils = utils

    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))

    # (3) Q-Former: v5 PreTrainedModel.init_weights / get_head_mask changed
    import models.Qformer as Q  # noqa: E402

    def _init_weights_compat(self):
        try:
            self.post_init()
        except Exception:
            self.apply(self._init_weights)

    Q.BertPreTrainedModel.init_weights = _init_weights_compat
    Q.BertModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n

    # (4) the BertConfig path is hard-coded (models/multimodal_encoder.py:90,192) -> bert-base-uncased defaults
    import models.multimodal_encoder as ME  # noqa: E402

    class _BertConfigShim(ME.BertConfig):
        @classmethod
        def from_pretrained(cls, *a, **k):
            cfg = ME.BertConfig()
            ov = getattr(_BertConfigShim, "_overrides", None)
            if ov:
                for kk, vv in ov.items():
                    setattr(cfg, kk, vv)
            return cfg

    ME.BertConfig = _BertConfigShim
    _installed = True


def set_bert_overrides(**kw) -> None:
    """Shrink the Q-Former BertConfig (hidden_size, heads, intermediate_size...) for small golden cases."""
    install_shims()
    import models.multimodal_encoder as ME

    ME.BertConfig._overrides = dict(kw)


def make_clip_dir(tmp: Path, hidden=1024, inter=4096, layers=24, heads=16, image=224, patch=14) -> Path:
    """(6) a synthetic CLIP checkpoint directory so VisualEncoder's from_pretrained calls work unmodified."""
    from transformers import CLIPVisionConfig, CLIPVisionModel

    d = tmp / "clip"
    d.mkdir(parents=True, exist_ok=True)
    cfg = CLIPVisionConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                           num_attention_heads=heads, image_size=image, patch_size=patch)
    m = CLIPVisionModel(cfg)
    m.save_pretrained(d)
    (d / "preprocessor_config.json").write_text(json.dumps({
        "crop_size": image, "do_center_crop": True, "do_normalize": True, "do_resize": True,
        "image_mean": [0.48145466, 0.4578275, 0.40821073], "image_std": [0.26862954, 0.26130258, 0.27577711],
        "resample": 3, "size": image, "image_processor_type": "CLIPImageProcessor"}))
    return d


BEATS_CFG_PUBLIC = {
    # public BEATs_iter3_plus_AS2M_finetuned_on_AS2M_cpt2 cfg (values come from the released checkpoint)
    "input_patch_size": 16, "embed_dim": 512, "conv_bias": False, "encoder_layers": 12, "encoder_embed_dim": 768,
    "encoder_ffn_embed_dim": 3072, "encoder_attention_heads": 12, "activation_fn": "gelu",
    "layer_wise_gradient_decay_ratio": 0.6, "layer_norm_first": False, "deep_norm": True, "dropout": 0.0,
    "attention_dropout": 0.0, "activation_dropout": 0.0, "encoder_layerdrop": 0.05, "dropout_input": 0.0,
    "conv_pos": 128, "conv_pos_groups": 16, "relative_position_embedding": True, "num_buckets": 320,
    "max_distance": 800, "gru_rel_pos": True, "finetuned_model": True, "predictor_dropout": 0.0,
    "predictor_class": 527,
}


def make_beats_ckpt(tmp: Path, cfg: dict) -> Path:
    install_shims()
    from models.beats.BEATs import BEATs, BEATsConfig

    m = BEATs(BEATsConfig(cfg))
    p = tmp / "beats.pt"
    torch.save({"cfg": cfg, "model": m.state_dict()}, p)
    return p


class FakeTokenizer:
    """Minimal tokenizer stand-in for initialize_MM_tokenizer (no LLaMA tokenizer files are available)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def add_tokens(self, toks, special_tokens=False):
        self.n += len(toks)
        return len(toks)


def build_reference_model(*, llama_cfg: dict, d_model: int, clip: dict, beats_cfg: dict, bert: dict | None,
                          image_size=224, patch_size=14, select_layer_list=(14, 22, 23), lora=True,
                          dtype=torch.float32, seed=42):
    """Construct the reference's UnifiedForCausalLM (+ hyper-LoRA wrap + encoders) exactly as
    scripts/quick_start.py:453-529 does, on synthetic checkpoints.  Returns (peft_or_plain_model, tokenizer)."""
    install_shims()
    set_bert_overrides(**(bert or {}))
    torch.manual_seed(seed)
    from transformers import LlamaConfig
    from models.unified_llama import UnifiedForCausalLM

    tmp = Path(tempfile.mkdtemp(prefix="crab_ref_"))
    clip_dir = make_clip_dir(tmp, image=image_size, patch=patch_size, **clip)
    beats_pt = make_beats_ckpt(tmp, beats_cfg)
    cfg = LlamaConfig(**llama_cfg)
    cfg._attn_implementation = "eager"
    model = UnifiedForCausalLM(cfg).to(dtype)
    if lora:
        from peft_hyper import LoraConfig, TaskType, get_peft_model

        lcfg = LoraConfig(task_type=TaskType.CAUSAL_LM, inference_mode=False, r=8, lora_alpha=16, lora_dropout=0.05,
                          lora_nums=3, target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj",
                                                       "down_proj", "up_proj"])
        model = get_peft_model(model, lcfg)
    tok = FakeTokenizer(cfg.vocab_size)
    model.get_model().pad_token_id = 0
    model.get_model().init_multimodal_modules(
        d_model=d_model, vit_ckpt_path=str(clip_dir), select_layer_list=list(select_layer_list),
        select_feature="patch", image_size=image_size, patch_size=patch_size, visual_query_token_nums=32,
        BEATs_ckpt_path=str(beats_pt), audio_query_token_nums=32, visual_branch=True, audio_branch=True,
        segment_branch=False)
    model.initialize_MM_tokenizer(tok, mask_token_nums=6)
    model.eval()
    return model, tok


def build_reference_qwen(*, qwen_cfg: dict, lora=True, dtype=torch.float32, seed=42):
    """The reference's Qwen2 wrapper (models/unified_qwen.py) + hyper-LoRA, decoder only: its multimodal entry is stale
    (SURVEY.md §2.1), so the contract for the Qwen backbone is forward/generate over `inputs_embeds`."""
    install_shims()
    torch.manual_seed(seed)
    from transformers import Qwen2Config
    from models.unified_qwen import UnifiedForCausalLM

    cfg = Qwen2Config(**qwen_cfg)
    cfg._attn_implementation = "eager"
    model = UnifiedForCausalLM(cfg).to(dtype)
    if lora:
        from peft_hyper import LoraConfig, TaskType, get_peft_model

        lcfg = LoraConfig(task_type=TaskType.CAUSAL_LM, inference_mode=False, r=8, lora_alpha=16, lora_dropout=0.05,
                          lora_nums=3, target_modules=["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj",
                                                       "down_proj", "up_proj"])
        model = get_peft_model(model, lcfg)
    model.eval()
    return model
