"""Host-side mirror of the reference's `models/unified_arch.py` module boundary, backed by the B200 engine.

Same public names and call semantics as the reference classes (so `scripts/quick_start.py` can import this module in
place of the reference's), but none of its implementation: the modules created here are *parameter containers*
whose names and shapes reproduce the reference state-dict layout (so `load_state_dict(finetune_weights.bin)` and
`peft_hyper.get_peft_model` keep working), while all arithmetic is done by `crab_b200.engine.CrabEngine`.

  UnifiedMetaModel.init_multimodal_modules      <- models/unified_arch.py:31-110
  UnifiedMetaModel.encode_video / encode_audio  <- models/unified_arch.py:113-155
  UnifiedMetaForCausalLM.prepare_multimodal_inputs   <- models/unified_arch.py:217-406 (generation branch)
  UnifiedMetaForCausalLM.initialize_MM_tokenizer     <- models/unified_arch.py:409-459
The segmentation branch (`segment_branch=True`) creates the reference SegModule's parameter names; its
arithmetic lives in crab_b200/seg.py.  The VQGAN mask tokeniser (`use_vqgan`) is not built and raises NotImplementedError.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from ..engine import BeatsConfig, ClipConfig, CrabConfig, CrabEngine, DecoderConfig, QformerConfig

# token order is part of the checkpoint contract: ids are assigned consecutively after the base vocabulary
_IMAGE_TOKENS = ["<image>", "<image_start>", "<image_end>"]
_VIDEO_TOKENS = ["<video>", "<video_start>", "<video_end>"]
_AUDIO_TOKENS = ["<audio>", "<audio_start>", "<audio_end>"]
_MASK_TOKENS = ["<mask_start>", "<mask_end>"]


class ParamTree(nn.Module):
    """A module tree built from {dotted.name: shape}; leaves are nn.Parameter.  Used to expose the reference's
    encoder / projector parameter names without instantiating HF / BEATs / Q-Former modules."""

    def __init__(self, manifest: Dict[str, Sequence[int]], dtype=torch.float32):
        super().__init__()
        for name, shape in manifest.items():
            node = self
            parts = name.split(".")
            for p in parts[:-1]:
                if not hasattr(node, p):
                    node.add_module(p, ParamTree({}))
                node = getattr(node, p)
            node.register_parameter(parts[-1], nn.Parameter(torch.zeros(tuple(shape), dtype=dtype), requires_grad=False))


def clip_manifest(hidden: int, inter: int, layers: int, patch: int, image: int) -> Dict[str, Tuple[int, ...]]:
    """Parameter names/shapes of HF CLIPVisionModel (`vision_tower.vision_model.*`)."""
    m: Dict[str, Tuple[int, ...]] = {}
    p = "vision_tower.vision_model."
    n_pos = (image // patch) ** 2 + 1
    m[p + "embeddings.class_embedding"] = (hidden,)
    m[p + "embeddings.patch_embedding.weight"] = (hidden, 3, patch, patch)
    m[p + "embeddings.position_embedding.weight"] = (n_pos, hidden)
    for nm in ("pre_layrnorm", "post_layernorm"):
        m[p + nm + ".weight"] = (hidden,)
        m[p + nm + ".bias"] = (hidden,)
    for i in range(layers):
        lp = f"{p}encoder.layers.{i}."
        for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
            m[lp + f"self_attn.{pr}.weight"] = (hidden, hidden)
            m[lp + f"self_attn.{pr}.bias"] = (hidden,)
        for ln in ("layer_norm1", "layer_norm2"):
            m[lp + ln + ".weight"] = (hidden,)
            m[lp + ln + ".bias"] = (hidden,)
        m[lp + "mlp.fc1.weight"] = (inter, hidden)
        m[lp + "mlp.fc1.bias"] = (inter,)
        m[lp + "mlp.fc2.weight"] = (hidden, inter)
        m[lp + "mlp.fc2.bias"] = (hidden,)
    return m


def beats_manifest(c: BeatsConfig) -> Dict[str, Tuple[int, ...]]:
    """Parameter names/shapes of the reference BEATs module (`audio_encoder.*`, models/beats/BEATs.py:75-103)."""
    m: Dict[str, Tuple[int, ...]] = {}
    p = "audio_encoder."
    m[p + "patch_embedding.weight"] = (c.embed, 1, c.patch, c.patch)
    m[p + "layer_norm.weight"] = (c.embed,)
    m[p + "layer_norm.bias"] = (c.embed,)
    m[p + "post_extract_proj.weight"] = (c.dim, c.embed)
    m[p + "post_extract_proj.bias"] = (c.dim,)
    m[p + "encoder.pos_conv.0.bias"] = (c.dim,)
    m[p + "encoder.pos_conv.0.weight_g"] = (1, 1, c.conv_pos)
    m[p + "encoder.pos_conv.0.weight_v"] = (c.dim, c.dim // c.conv_groups, c.conv_pos)
    m[p + "encoder.layer_norm.weight"] = (c.dim,)
    m[p + "encoder.layer_norm.bias"] = (c.dim,)
    hd = c.dim // c.heads
    for i in range(c.layers):
        lp = f"{p}encoder.layers.{i}."
        for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
            m[lp + f"self_attn.{pr}.weight"] = (c.dim, c.dim)
            m[lp + f"self_attn.{pr}.bias"] = (c.dim,)
        m[lp + "self_attn.grep_linear.weight"] = (8, hd)
        m[lp + "self_attn.grep_linear.bias"] = (8,)
        m[lp + "self_attn.grep_a"] = (1, c.heads, 1, 1)
        m[lp + "self_attn.relative_attention_bias.weight"] = (c.num_buckets, c.heads)
        for ln in ("self_attn_layer_norm", "final_layer_norm"):
            m[lp + ln + ".weight"] = (c.dim,)
            m[lp + ln + ".bias"] = (c.dim,)
        m[lp + "fc1.weight"] = (c.ffn, c.dim)
        m[lp + "fc1.bias"] = (c.ffn,)
        m[lp + "fc2.weight"] = (c.dim, c.ffn)
        m[lp + "fc2.bias"] = (c.dim,)
    return m


def projector_manifest(kind: str, q: QformerConfig, enc_width: int, d_model: int, n_query: int) -> Dict[str, Tuple[int, ...]]:
    """VLProjector / ALProjector parameters on the executed path (models/multimodal_encoder.py:87-144, 189-262): the
    Q-Former text branch, word/position embeddings and `cls` head exist in reference checkpoints but are never
    executed — they are accepted and ignored by load_state_dict(strict=False)."""
    ln, qf, qt, proj = {"visual": ("visual_ln", "visual_Qformer", "visual_query_tokens", "visual_proj"),
                        "audio": ("audio_ln", "audio_Qformer", "audio_query_tokens", "audio_proj")}[kind]
    m: Dict[str, Tuple[int, ...]] = {ln + ".weight": (enc_width,), ln + ".bias": (enc_width,), qt: (1, n_query, q.hidden)}
    b = qf + ".bert."
    m[b + "embeddings.LayerNorm.weight"] = (q.hidden,)
    m[b + "embeddings.LayerNorm.bias"] = (q.hidden,)

    def lin(n, o, i):
        m[n + ".weight"] = (o, i)
        m[n + ".bias"] = (o,)

    def lnorm(n):
        m[n + ".weight"] = (q.hidden,)
        m[n + ".bias"] = (q.hidden,)

    for i in range(q.layers):
        lp = f"{b}encoder.layer.{i}."
        for blk, kv_in in (("attention", q.hidden), ("crossattention", enc_width)):
            lin(lp + blk + ".self.query", q.hidden, q.hidden)
            lin(lp + blk + ".self.key", q.hidden, kv_in)
            lin(lp + blk + ".self.value", q.hidden, kv_in)
            lin(lp + blk + ".output.dense", q.hidden, q.hidden)
            lnorm(lp + blk + ".output.LayerNorm")
        lin(lp + "intermediate_query.dense", q.inter, q.hidden)
        lin(lp + "output_query.dense", q.hidden, q.inter)
        lnorm(lp + "output_query.LayerNorm")
    lin(proj + ".0", d_model, q.hidden)
    lin(proj + ".2", d_model, d_model)
    return m


def decoder_manifest(d: DecoderConfig, lora: bool = True) -> Dict[str, Tuple[int, ...]]:
    """Decoder parameter names/shapes as in reference checkpoints (HF LLaMA/Qwen2 names + the hyper-LoRA tensors
    peft_hyper adds per wrapped linear: lora_A, lora_route, lora_B0..B{n-1}; peft_hyper/tuners/lora.py:283-290)."""
    m: Dict[str, Tuple[int, ...]] = {"model.embed_tokens.weight": (d.vocab, d.hidden), "model.norm.weight": (d.hidden,),
                                     "lm_head.weight": (d.vocab, d.hidden)}
    nq, nk = d.heads * d.head_dim, d.kv_heads * d.head_dim
    for i in range(d.layers):
        lp = f"model.layers.{i}."
        lins = {"self_attn.q_proj": (nq, d.hidden), "self_attn.k_proj": (nk, d.hidden), "self_attn.v_proj": (nk, d.hidden),
                "self_attn.o_proj": (d.hidden, nq), "mlp.gate_proj": (d.inter, d.hidden), "mlp.up_proj": (d.inter, d.hidden),
                "mlp.down_proj": (d.hidden, d.inter)}
        for n, (o, k) in lins.items():
            m[lp + n + ".weight"] = (o, k)
            if d.qkv_bias and n.split(".")[-1] in ("q_proj", "k_proj", "v_proj"):
                m[lp + n + ".bias"] = (o,)
            if lora:
                m[lp + n + ".lora_A.weight"] = (d.lora_r, k)
                m[lp + n + ".lora_route.weight"] = (d.lora_nums, k)
                for j in range(d.lora_nums):
                    m[lp + n + f".lora_B{j}.weight"] = (o, d.lora_r)
        m[lp + "input_layernorm.weight"] = (d.hidden,)
        m[lp + "post_attention_layernorm.weight"] = (d.hidden,)
    return m


def full_manifest(cfg: CrabConfig, d_model: Optional[int] = None, lora: bool = True) -> Dict[str, Tuple[int, ...]]:
    """Every tensor the engine reads, under the reference's state-dict names (`model.visual_encoder.…`, …)."""
    d_model = d_model or cfg.decoder.hidden
    m = decoder_manifest(cfg.decoder, lora)
    c = cfg.clip
    for k, v in clip_manifest(c.hidden, c.inter, c.layers, c.patch, c.image).items():
        m["model.visual_encoder." + k] = v
    for k, v in beats_manifest(cfg.beats).items():
        m["model.audio_encoder." + k] = v
    for k, v in projector_manifest("visual", cfg.qformer, c.hidden, d_model, cfg.n_query).items():
        m["model.vl_projector." + k] = v
    for k, v in projector_manifest("audio", cfg.qformer, cfg.beats.dim, d_model, cfg.n_query).items():
        m["model.al_projector." + k] = v
    return m


def seg_manifest(d_model: int, prompt_dim: int = 256, vit_dim: int = 1024, scales: int = 2, depth: int = 2, queries: int = 300,
                 query_layers: int = 2, mlp_dim: int = 2048) -> Dict[str, Tuple[int, ...]]:
    """Parameter / buffer names and shapes of the reference `SegModule` (models/multimodal_encoder.py:268-353; mask decoder
    :891-963, two-way transformer :1163-1299, query generator :1396-1439), i.e. the `seg_module.*` keys of a checkpoint."""
    E, m = prompt_dim, {}

    def lin(name, o, i, bias=True):
        m[name + ".weight"] = (o, i)
        if bias:
            m[name + ".bias"] = (o,)

    def norm(name, c=E):
        m[name + ".weight"], m[name + ".bias"] = (c,), (c,)

    lin("text_hidden_fcs.0.0", d_model, d_model)
    lin("text_hidden_fcs.0.2", E, d_model)
    m["no_mask_embed.weight"] = (1, E)
    m["image_feature_neck.0.weight"] = (E, vit_dim, 1, 1)
    norm("image_feature_neck.1")
    m["image_feature_neck.2.weight"] = (E, E, 3, 3)
    norm("image_feature_neck.3")
    m["pe_layer.positional_encoding_gaussian_matrix"] = (2, E // 2)
    md = "mask_decoder."
    for l in range(scales):
        tp = f"{md}transformer.{l}."

        def attn(name, internal):
            for pj in ("q_proj", "k_proj", "v_proj"):
                lin(f"{name}.{pj}", internal, E)
            lin(f"{name}.out_proj", E, internal)

        for i in range(depth):
            lp = f"{tp}layers.{i}."
            attn(lp + "self_attn", E)
            norm(lp + "norm1")
            attn(lp + "cross_attn_token_to_image", E // 2)
            norm(lp + "norm2")
            lin(lp + "mlp.lin1", mlp_dim, E)
            lin(lp + "mlp.lin2", E, mlp_dim)
            norm(lp + "norm3")
            norm(lp + "norm4")
            attn(lp + "cross_attn_image_to_token", E // 2)
        attn(tp + "final_attn_token_to_image", E // 2)
        norm(tp + "norm_final_attn")
    m[md + "avs_query_tokens.weight"] = (queries, E)
    for i in range(query_layers):
        qp = f"{md}query_generator.layers.{i}."
        for a in ("self_attn", "cross_attn"):
            m[qp + a + ".in_proj_weight"], m[qp + a + ".in_proj_bias"] = (3 * E, E), (3 * E,)
            lin(qp + a + ".out_proj", E, E)
        lin(qp + "ffn.0", mlp_dim, E)
        lin(qp + "ffn.2", E, mlp_dim)
        for n_ in ("norm1", "norm2", "norm3"):
            norm(qp + n_)
    dims = [queries, E, E, E // 8]
    for i in range(3):
        m[f"{md}hyper_mlp_out.layers.{i}.weight"], m[f"{md}hyper_mlp_out.layers.{i}.bias"] = (dims[i + 1], dims[i], 1, 1), (dims[i + 1],)
    dims = [E, E, E, E // 8]
    for i in range(3):
        lin(f"{md}hyper_mlp.layers.{i}", dims[i + 1], dims[i])
    m[md + "output_upscaling.0.weight"], m[md + "output_upscaling.0.bias"] = (E, E // 8, 2, 2), (E // 8,)
    norm(md + "output_upscaling.1", E // 8)
    m[md + "upsample_2x.0.weight"], m[md + "upsample_2x.0.bias"] = (E, E, 2, 2), (E,)
    norm(md + "upsample_2x.1")
    m[md + "pe1.positional_encoding_gaussian_matrix"] = (2, E // 2)
    m[md + "level_embed.weight"] = (scales, E)
    m[md + "ms3_s4_classfier.weight"] = (1, E // 8, 1, 1)
    m[md + "avss_classifier.weight"] = (71, E // 8, 1, 1)
    return m


def special_token_ids(base_vocab: int, mask_token_nums: int = 6) -> Dict[str, int]:
    names = _IMAGE_TOKENS + _VIDEO_TOKENS + _AUDIO_TOKENS + _MASK_TOKENS + [f"<mask_{i}>" for i in range(mask_token_nums)]
    return {t: base_vocab + i for i, t in enumerate(names)}


def _load_weights_into(tree: nn.Module, sd: Dict[str, torch.Tensor], prefix: str = "") -> List[str]:
    own = dict(tree.named_parameters())
    missing = []
    with torch.no_grad():
        for name, p in own.items():
            t = sd.get(prefix + name)
            if t is None:
                missing.append(name)
            else:
                p.copy_(t.to(p.dtype).reshape(p.shape))
    return missing


class VisualEncoder(ParamTree):
    """Container with the reference VisualEncoder's names (models/multimodal_encoder.py:33-84) + its image_processor."""

    def __init__(self, model_name_or_path: str, select_layer_list, select_feature: str = "patch"):
        cfg_path = os.path.join(model_name_or_path, "config.json")
        with open(cfg_path) as f:
            raw = json.load(f)
        vc = raw.get("vision_config", raw)
        self.clip_cfg = ClipConfig(hidden=vc["hidden_size"], inter=vc["intermediate_size"], heads=vc["num_attention_heads"],
                                   layers=vc["num_hidden_layers"], patch=vc["patch_size"], image=vc["image_size"],
                                   eps=vc.get("layer_norm_eps", 1e-5))
        c = self.clip_cfg
        super().__init__(clip_manifest(c.hidden, c.inter, c.layers, c.patch, c.image))
        self.select_layer_list = list(select_layer_list)
        self.select_feature = select_feature
        if select_feature != "patch":
            raise ValueError(f"Unexpected select feature: {select_feature}")
        try:
            from transformers import CLIPImageProcessor

            self.image_processor = CLIPImageProcessor.from_pretrained(model_name_or_path, local_files_only=True)
        except Exception:  # preprocessing is host-side and outside the accelerated path
            self.image_processor = None
        self._load_checkpoint(model_name_or_path)

    def _load_checkpoint(self, path: str):
        sd = None
        st = os.path.join(path, "model.safetensors")
        pt = os.path.join(path, "pytorch_model.bin")
        if os.path.exists(st):
            from safetensors.torch import load_file

            sd = load_file(st)
        elif os.path.exists(pt):
            sd = torch.load(pt, map_location="cpu")
        if sd is None:
            raise FileNotFoundError(f"no CLIP weights under {path}")
        missing = _load_weights_into(self.vision_tower, sd, "")
        if any("post_layernorm" not in m for m in missing):
            raise RuntimeError(f"CLIP checkpoint is missing {missing[:4]}…")


class AudioEncoder(ParamTree):
    """Container with the reference AudioEncoder's names (models/multimodal_encoder.py:150-186); reads the BEATs
    checkpoint's own `cfg` dict exactly as the reference does (:157-158)."""

    def __init__(self, ckpt_path: str):
        ck = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        c = ck["cfg"]
        assert c.get("deep_norm", False) and not c.get("layer_norm_first", False) and c.get("gru_rel_pos", False) and \
            c.get("relative_position_embedding", False), "only the BEATs iter3+ configuration (deep-norm, gated rel-pos) is supported"
        self.beats_cfg = BeatsConfig(patch=c["input_patch_size"], embed=c["embed_dim"], dim=c["encoder_embed_dim"],
                                     ffn=c["encoder_ffn_embed_dim"], heads=c["encoder_attention_heads"],
                                     layers=c["encoder_layers"], conv_pos=c["conv_pos"], conv_groups=c["conv_pos_groups"],
                                     num_buckets=c["num_buckets"], max_distance=c["max_distance"])
        super().__init__(beats_manifest(self.beats_cfg))
        sd = dict(ck["model"])
        pc = "encoder.pos_conv.0."
        if pc + "weight_g" not in sd and pc + "parametrizations.weight.original0" in sd:
            sd[pc + "weight_g"] = sd[pc + "parametrizations.weight.original0"]
            sd[pc + "weight_v"] = sd[pc + "parametrizations.weight.original1"]
        rel0 = "encoder.layers.0.self_attn.relative_attention_bias.weight"
        for i in range(1, self.beats_cfg.layers):  # shared parameter (backbone.py:78-81)
            sd.setdefault(f"encoder.layers.{i}.self_attn.relative_attention_bias.weight", sd[rel0])
        missing = _load_weights_into(self.audio_encoder, sd, "")
        if missing:
            raise RuntimeError(f"BEATs checkpoint is missing {missing[:4]}…")


class UnifiedMetaModel:
    """Mixin for the decoder-parameter container (see unified_llama.UnifiedModel)."""

    def init_multimodal_modules(
        self, d_model=4096, vit_ckpt_path="", select_layer_list=(14, 22, 23), select_feature="patch", image_size=224,
        patch_size=14, visual_query_token_nums=32, BEATs_ckpt_path="", audio_query_token_nums=32, image_scale_nums=2,
        token_nums_per_scale=3, avs_query_num=300, num_classes=1, query_generator_num_layers=2, prompt_embed_dim=256,
        mask_decoder_transformer_depth=2, low_res_mask_size=112, dice_loss_weight=0.5, bce_loss_weight=2.0,
        vit_image_embedding_dim=1024, visual_branch=False, audio_branch=False, segment_branch=False, use_vqgan=False,
        qformer_config: Optional[QformerConfig] = None,
    ):
        """Same keyword surface as the reference (models/unified_arch.py:31-61).  `qformer_config` is an extra,
        optional knob: the reference reads bert-base-uncased's config from a hard-coded path
        (models/multimodal_encoder.py:90,192); the defaults here are those values."""
        if use_vqgan:
            raise NotImplementedError("the VQGAN mask tokeniser is out of scope for the B200 path")
        if segment_branch:
            # parameter container with the reference SegModule's names; the arithmetic is crab_b200.seg.SegHead
            assert (image_scale_nums, token_nums_per_scale, prompt_embed_dim) == (2, 3, 256), \
                "the B200 segmentation head is built for the reference's shipped geometry (2 scales x 3 tokens, 256-d prompts)"
            self.low_res_mask_size = low_res_mask_size
            self.seg_module = ParamTree(seg_manifest(d_model, prompt_embed_dim, vit_image_embedding_dim, image_scale_nums,
                                                     mask_decoder_transformer_depth, avs_query_num, query_generator_num_layers))
            for n_, p_ in self.seg_module.named_parameters():   # the two PositionEmbeddingRandom buffers are random in the reference
                if n_.endswith("positional_encoding_gaussian_matrix"):
                    nn.init.normal_(p_)
        q = qformer_config or QformerConfig()
        self.qformer_cfg = q
        self.select_layer_list = list(select_layer_list)
        self.n_query = visual_query_token_nums
        assert visual_query_token_nums == audio_query_token_nums
        if visual_branch:
            self.visual_encoder = VisualEncoder(vit_ckpt_path, select_layer_list, select_feature)
            self.vl_projector = ParamTree(projector_manifest("visual", q, 1024, d_model, visual_query_token_nums))
            nn.init.normal_(self.vl_projector.visual_query_tokens, std=0.02)
        if audio_branch:
            self.audio_encoder = AudioEncoder(BEATs_ckpt_path)
            self.al_projector = ParamTree(projector_manifest("audio", q, 768, d_model, audio_query_token_nums))
            nn.init.normal_(self.al_projector.audio_query_tokens, std=0.02)
        self._engine_stale()


class UnifiedMetaForCausalLM:
    """Mixin for the causal-LM wrapper: tokenizer bookkeeping + the multimodal entry points."""

    KEYS: List[str] = ["<image>", "<video>", "<audio>"]

    def initialize_MM_tokenizer(self, tokenizer, mask_token_nums=6, output_embeddings_require_grad=False, use_vqgan=False):
        if use_vqgan:
            raise NotImplementedError("VQGAN mask tokens are out of scope")
        base = len(tokenizer)
        special = _IMAGE_TOKENS + _VIDEO_TOKENS + _AUDIO_TOKENS + _MASK_TOKENS
        tokenizer.add_tokens(special, special_tokens=True)
        seg = [f"<mask_{i}>" for i in range(mask_token_nums)]
        tokenizer.add_tokens(seg, special_tokens=False)
        names = special + seg
        self.KEYS = ["<image>", "<video>", "<audio>"]
        self.MASK = seg
        self.SPECIAL_TOKEN_2_IDS = {t: base + i for i, t in enumerate(names)}
        self.IDS_2_SPECIAL_TOKEN = {base + i: t for i, t in enumerate(names)}
        self.resize_token_embeddings(len(tokenizer))

    # -- encoders (same shapes as the reference helpers) -------------------------------------------------------
    def encode_video(self, video: torch.Tensor, batch_first: bool = True):
        """(b,t,3,H,W) [or (t,3,H,W) with batch_first=False] -> Q-Former features (b, t*32, d_model).  Returns
        ([], [feature]) to keep the reference's `(vit_feature_list, qformer_feature_list)` shape; only the last
        tap's projection exists because it is the only one the generation path reads (unified_arch.py:290)."""
        eng = self.engine()
        v = video if batch_first else video.unsqueeze(0)
        b, t = v.shape[:2]
        f = eng.encode_video(v.reshape(b * t, *v.shape[2:]).to(eng.dev, torch.float32).contiguous()).view(b, t * eng.cfg.n_query, -1)
        return [], [f if batch_first else f[0]]

    def encode_audio(self, audio: torch.Tensor, batch_first: bool = True):
        eng = self.engine()
        a = audio if batch_first else audio.unsqueeze(0)
        if a.dim() == 3:
            a = a.unsqueeze(1)
        b, t = a.shape[:2]
        f = eng.encode_audio(a.reshape(b * t, *a.shape[2:]).to(eng.dev, torch.float32).contiguous()).view(b, t * eng.cfg.n_query, -1)
        return f if batch_first else f[0]

    def encode_ids(self, ids: torch.Tensor) -> torch.Tensor:
        eng = self.engine()
        ids = ids.to(eng.dev).long().reshape(-1)
        out = torch.empty((ids.numel(), eng.cfg.decoder.hidden), device=eng.dev, dtype=torch.bfloat16)
        from .. import ops

        ops.gather_rows(eng.embed, out, ids.numel(), out.shape[1], src_rows=ids.contiguous())
        return out

    def prepare_multimodal_inputs(self, batch_input_ids, batch_labels, batch_X_modals, batch_task_names=None,
                                  return_multi_scale_features=False, return_gt_mask=False):
        eng = self.engine()
        taps = tuple(eng.cfg.select_layers[:2]) if return_multi_scale_features else ()
        embeds, mask, pos = eng.prepare_inputs(batch_input_ids, batch_X_modals, want_image_taps=taps)
        labels = None
        if batch_labels is not None and all(lab is not None for lab in batch_labels):
            S = embeds.shape[1]
            rows = []
            for ids, lab, m in zip(batch_input_ids, batch_labels, mask):
                # modality spans and left padding are ignored by the loss (-100), text labels pass through
                full = torch.full((S,), -100, dtype=torch.long)
                keep = [i for i, t in enumerate(ids.tolist()) if t not in self.IDS_2_SPECIAL_TOKEN or
                        self.IDS_2_SPECIAL_TOKEN[t] not in self.KEYS]
                valid = int(m.sum())
                pos_out, i_prev = S - valid, 0
                nq = eng.cfg.n_query
                for i, t in enumerate(ids.tolist()):
                    if i in keep:
                        full[pos_out] = lab[i]
                        pos_out += 1
                    else:
                        X = batch_X_modals[len(rows)][self.IDS_2_SPECIAL_TOKEN[t]]
                        pos_out += (X.shape[0] if X.dim() > 2 else 1) * nq
                rows.append(full)
            labels = torch.stack(rows).to(eng.dev)
        out = {"input_ids": None, "inputs_embeds": embeds, "attention_mask": mask.to(eng.dev), "labels": labels,
               "position_ids": pos.to(eng.dev)}
        if return_multi_scale_features:
            # [(bs, 256, 1024)] * scales: ViT taps of every sample's '<image>' (zeros for samples without one, unified_arch.py:233-238)
            tok, dim = (eng.cfg.clip.image // eng.cfg.clip.patch) ** 2, eng.cfg.clip.hidden
            feats = []
            for sidx in range(len(taps)):
                rows = [t[sidx][:tok] if t is not None else torch.zeros((tok, dim), device=eng.dev, dtype=torch.bfloat16)
                        for t in eng.image_taps]
                feats.append(torch.stack(rows, 0))
            out["multi_scale_image_features"] = feats
        if return_gt_mask:
            gts = [X.get("<mask>") for X in batch_X_modals]
            out["gt_mask"] = torch.stack([g.to(eng.dev) for g in gts], 0) if all(g is not None for g in gts) else None
        return out


def build_crab_config(decoder: DecoderConfig, model, max_ctx: int) -> CrabConfig:
    inner = model.get_model()
    clip = getattr(getattr(inner, "visual_encoder", None), "clip_cfg", ClipConfig())
    beats = getattr(getattr(inner, "audio_encoder", None), "beats_cfg", BeatsConfig())
    return CrabConfig(decoder=decoder, clip=clip, beats=beats, qformer=getattr(inner, "qformer_cfg", QformerConfig()),
                      select_layers=tuple(getattr(inner, "select_layer_list", (14, 22, 23))),
                      n_query=getattr(inner, "n_query", 32), pad_token_id=int(getattr(inner, "pad_token_id", 0) or 0),
                      max_ctx=max_ctx, special_ids=dict(getattr(model, "SPECIAL_TOKEN_2_IDS", {})),
                      low_res_mask_size=int(getattr(inner, "low_res_mask_size", 112)))
