"""TEST INFRASTRUCTURE — CPU restatement (plain torch, functional, no nn.Module) of Crab's AV-prompt hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker or the timed CPU baseline — never as part of the product path (`crab_b200/`).

Every function cites the reference code it restates (paths relative to the reference checkout; "HF" =
transformers, the third-party dependency the reference delegates CLIP / LLaMA / Qwen2 arithmetic to — pinned
`transformers==4.37.2` in requirements.txt:181, restated here from the in-tree 4.37-era copies
models/modeling_llama.py and models/qwen/modeling_qwen2.py, and from transformers/models/clip/modeling_clip.py).

Pinning: the reference ships no tests or golden vectors, so this restatement is pinned against
outputs of the reference's own modules run in the build container (`oracle/make_golden.py`, through
`oracle/ref_shims.py`); the resulting fixtures live in tests/golden/ and `tests/test_oracle_golden.py` checks them.

Weights are passed as a flat state dict with the reference's parameter names (the `base_model.model.` prefix PEFT
adds is stripped by `strip_peft_prefix`).  Arithmetic follows the dtype of the weights (fp32 oracle; bf16 gives the
"HF-bf16 noise yardstick") with the same fp32 islands the reference has (RMSNorm, softmax, router softmax, GELU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def strip_peft_prefix(sd: SD) -> SD:
    out = {}
    for k, v in sd.items():
        out[k[len("base_model.model."):] if k.startswith("base_model.model.") else k] = v
    return out


def _lin(x, sd: SD, name: str, bias: bool = True):
    b = sd.get(name + ".bias") if bias else None
    return F.linear(x, sd[name + ".weight"], b)


def _ln(x, sd: SD, name: str, eps: float):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], eps)


# ----------------------------------------------------------------------------------------------------------------
# CLIP ViT (HF transformers/models/clip/modeling_clip.py; invoked at models/multimodal_encoder.py:66-72)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class ClipCfg:
    hidden: int = 1024
    heads: int = 16
    layers: int = 24
    patch: int = 14
    eps: float = 1e-5


def clip_hidden_states(sd: SD, prefix: str, pixels: torch.Tensor, cfg: ClipCfg, upto: Optional[int] = None):
    """pixels (n,3,H,W) -> list of hidden states: [0] = post-pre_layrnorm embeddings, [i] = output of layer i.
    HF CLIPVisionEmbeddings (:138-232), CLIPEncoderLayer (:354-400), CLIPAttention (:282-336, scale hd^-1/2),
    CLIPMLP quick_gelu (:339-351), pre_layrnorm (:659,:677)."""
    p = prefix + "vision_model."
    w = sd[p + "embeddings.patch_embedding.weight"]
    x = F.conv2d(pixels.to(w.dtype), w, None, stride=cfg.patch)  # (n, D, g, g)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[p + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"][: x.shape[1] + 1]
    x = _ln(x, sd, p + "pre_layrnorm", cfg.eps)
    hs = [x]
    hd = cfg.hidden // cfg.heads
    n_layers = cfg.layers if upto is None else min(upto, cfg.layers)
    for i in range(n_layers):
        lp = f"{p}encoder.layers.{i}."
        h = _ln(x, sd, lp + "layer_norm1", cfg.eps)
        n, s, _ = h.shape
        q = _lin(h, sd, lp + "self_attn.q_proj").view(n, s, cfg.heads, hd).transpose(1, 2)
        k = _lin(h, sd, lp + "self_attn.k_proj").view(n, s, cfg.heads, hd).transpose(1, 2)
        v = _lin(h, sd, lp + "self_attn.v_proj").view(n, s, cfg.heads, hd).transpose(1, 2)
        a = torch.matmul(q, k.transpose(-1, -2)) * hd ** -0.5
        a = torch.softmax(a, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(a, v).transpose(1, 2).reshape(n, s, cfg.hidden)
        x = x + _lin(o, sd, lp + "self_attn.out_proj")
        h = _ln(x, sd, lp + "layer_norm2", cfg.eps)
        h = _lin(h, sd, lp + "mlp.fc1")
        h = h * torch.sigmoid(1.702 * h)
        x = x + _lin(h, sd, lp + "mlp.fc2")
        hs.append(x)
    return hs


def visual_encoder(sd: SD, video: torch.Tensor, cfg: ClipCfg, select_layers: Sequence[int]) -> List[torch.Tensor]:
    """VisualEncoder.forward / encode_video / feature_select (models/multimodal_encoder.py:52-84): video
    (b,t,3,H,W) -> one (b, t*n_patch, D) tensor per selected hidden state, CLS dropped ('patch')."""
    b, t = video.shape[:2]
    hs = clip_hidden_states(sd, "model.visual_encoder.vision_tower.", video.reshape(b * t, *video.shape[2:]), cfg,
                            upto=max(select_layers))
    out = []
    for lyr in select_layers:
        f = hs[lyr][:, 1:]
        out.append(f.reshape(b, t * f.shape[1], f.shape[2]))
    return out


# ----------------------------------------------------------------------------------------------------------------
# Q-Former (models/Qformer.py) and the two projectors (models/multimodal_encoder.py:87-144, 189-262)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class QformerCfg:
    hidden: int = 768
    heads: int = 12
    layers: int = 2
    eps: float = 1e-12


def _bert_attn(x, kv, sd: SD, p: str, heads: int):
    """BertSelfAttention.forward (models/Qformer.py:171-277) with all-ones masks (additive 0, :803)."""
    n, s, d = x.shape
    hd = d // heads
    q = _lin(x, sd, p + "query").view(n, s, heads, hd).transpose(1, 2)
    k = _lin(kv, sd, p + "key").view(n, kv.shape[1], heads, hd).transpose(1, 2)
    v = _lin(kv, sd, p + "value").view(n, kv.shape[1], heads, hd).transpose(1, 2)
    a = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    a = torch.softmax(a, dim=-1)
    return torch.matmul(a, v).permute(0, 2, 1, 3).reshape(n, s, d)


def qformer(sd: SD, prefix: str, query_tokens: torch.Tensor, enc: torch.Tensor, cfg: QformerCfg) -> torch.Tensor:
    """BertModel.forward with query_embeds only (models/Qformer.py:806-967): BertEmbeddings = LayerNorm(query)
    (:79-110); per layer (:404-476): self-attn + BertSelfOutput post-LN (:280-292), cross-attn every layer
    (cross_attention_freq=1) + post-LN, intermediate_query (GELU) / output_query + post-LN (:483-486)."""
    p = prefix + "bert."
    x = _ln(query_tokens.expand(enc.shape[0], -1, -1), sd, p + "embeddings.LayerNorm", cfg.eps)
    for i in range(cfg.layers):
        lp = f"{p}encoder.layer.{i}."
        a = _bert_attn(x, x, sd, lp + "attention.self.", cfg.heads)
        x = _ln(_lin(a, sd, lp + "attention.output.dense") + x, sd, lp + "attention.output.LayerNorm", cfg.eps)
        a = _bert_attn(x, enc, sd, lp + "crossattention.self.", cfg.heads)
        x = _ln(_lin(a, sd, lp + "crossattention.output.dense") + x, sd, lp + "crossattention.output.LayerNorm", cfg.eps)
        h = F.gelu(_lin(x, sd, lp + "intermediate_query.dense"))
        x = _ln(_lin(h, sd, lp + "output_query.dense") + x, sd, lp + "output_query.LayerNorm", cfg.eps)
    return x


def _mlp2(x, sd: SD, p: str):
    """build_mlp(depth=2): Linear -> GELU -> Linear (models/multimodal_encoder.py:25-30)."""
    return _lin(F.gelu(_lin(x, sd, p + "0")), sd, p + "2")


def vl_projector(sd: SD, feat: torch.Tensor, qcfg: QformerCfg, image_tokens: int, nq: int = 32) -> torch.Tensor:
    """VLProjector.forward (models/multimodal_encoder.py:119-144): (b, t*n, 1024) -> (b, t*32, d_model)."""
    p = "model.vl_projector."
    b, tn, dim = feat.shape
    t = tn // image_tokens
    f = _ln(feat.reshape(b * t, image_tokens, dim), sd, p + "visual_ln", 1e-5)
    q = qformer(sd, p + "visual_Qformer.", sd[p + "visual_query_tokens"], f, qcfg)
    out = _mlp2(q[:, :nq], sd, p + "visual_proj.")
    return out.reshape(b, t * nq, -1)


def al_projector(sd: SD, feat: torch.Tensor, qcfg: QformerCfg, nq: int = 32) -> torch.Tensor:
    """ALProjector.forward, 4-D branch (models/multimodal_encoder.py:226-244): (b,t,n,768) -> (b, t*32, d_model)."""
    p = "model.al_projector."
    b, t, n, d = feat.shape
    f = _ln(feat.reshape(b * t, n, d), sd, p + "audio_ln", 1e-5)
    q = qformer(sd, p + "audio_Qformer.", sd[p + "audio_query_tokens"], f, qcfg)
    return _mlp2(q[:, :nq].reshape(b, t * nq, -1), sd, p + "audio_proj.")


# ----------------------------------------------------------------------------------------------------------------
# BEATs (models/beats/BEATs.py, models/beats/backbone.py)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class BeatsCfg:
    patch: int = 16
    embed: int = 512
    dim: int = 768
    heads: int = 12
    layers: int = 12
    conv_pos: int = 128
    conv_groups: int = 16
    num_buckets: int = 320
    max_distance: int = 800
    eps: float = 1e-5

    @property
    def alpha(self) -> float:  # deep_norm_alpha (backbone.py:208-212)
        return math.pow(2 * self.layers, 0.25)


def beats_relative_buckets(qlen: int, klen: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """_relative_positions_bucket, bidirectional (models/beats/backbone.py:392-417) on memory - context."""
    ctx = torch.arange(qlen, dtype=torch.long)[:, None]
    mem = torch.arange(klen, dtype=torch.long)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (nb - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rel, large)


def beats_position_bias(sd: SD, prefix: str, T: int, cfg: BeatsCfg) -> torch.Tensor:
    """compute_bias (backbone.py:419-430): (H, T, T) table from layer 0's relative_attention_bias (shared :78-81)."""
    b = beats_relative_buckets(T, T, cfg.num_buckets, cfg.max_distance)
    return sd[prefix + "encoder.layers.0.self_attn.relative_attention_bias.weight"][b].permute(2, 0, 1)


def beats_pos_conv_weight(sd: SD, prefix: str) -> torch.Tensor:
    """weight_norm(dim=2) folded: w = v * g / ||v||_(0,1)  (backbone.py:45)."""
    p = prefix + "encoder.pos_conv.0."
    if p + "weight_g" in sd:
        g, v = sd[p + "weight_g"], sd[p + "weight_v"]
    else:
        g, v = sd[p + "parametrizations.weight.original0"], sd[p + "parametrizations.weight.original1"]
    return v * (g / v.float().norm(dim=(0, 1), keepdim=True).to(v.dtype))


def beats_attention(sd: SD, lp: str, x: torch.Tensor, pos_bias: torch.Tensor, cfg: BeatsCfg) -> torch.Tensor:
    """MultiheadAttention.forward (backbone.py:432-684) on (B,T,C): q scaled by hd^-1/2/32, (qk - rowmax)*32
    == hd^-1/2 * qk up to a per-row constant (:513-515, :623-624), plus the gated relative position bias computed
    from the *unscaled* q (:650-662)."""
    B, T, Cdim = x.shape
    H, hd = cfg.heads, Cdim // cfg.heads
    q = _lin(x, sd, lp + "q_proj")
    k = _lin(x, sd, lp + "k_proj")
    v = _lin(x, sd, lp + "v_proj")
    qh = q.view(B, T, H, hd).transpose(1, 2)  # unscaled
    kh = k.view(B, T, H, hd).transpose(1, 2)
    vh = v.view(B, T, H, hd).transpose(1, 2)
    alpha = 32.0
    qs = qh * (hd ** -0.5) * (1.0 / alpha)
    aw = torch.matmul(qs, kh.transpose(-1, -2))
    aw = (aw - aw.max(dim=-1, keepdim=True)[0]) * alpha
    g = _lin(qh, sd, lp + "grep_linear").view(B, H, T, 2, 4).sum(-1)
    gate_a, gate_b = torch.sigmoid(g).chunk(2, dim=-1)
    gate = gate_a * (gate_b * sd[lp + "grep_a"] - 1.0) + 2.0  # (B,H,T,1)
    aw = aw + gate * pos_bias.unsqueeze(0)
    pr = torch.softmax(aw, dim=-1)
    o = torch.matmul(pr, vh).transpose(1, 2).reshape(B, T, Cdim)
    return _lin(o, sd, lp + "out_proj")


def beats_extract_features(sd: SD, prefix: str, fbank: torch.Tensor, cfg: BeatsCfg) -> torch.Tensor:
    """BEATs.extract_features(feature_only=True) (BEATs.py:134-182) + TransformerEncoder.extract_features
    (backbone.py:109-150) + post-LN deep-norm layers (backbone.py:248-273).  fbank (B,T,128) -> (B, n, 768)."""
    w = sd[prefix + "patch_embedding.weight"]
    x = F.conv2d(fbank.to(w.dtype).unsqueeze(1), w, None, stride=cfg.patch)
    x = x.reshape(x.shape[0], x.shape[1], -1).transpose(1, 2)
    x = _ln(x, sd, prefix + "layer_norm", cfg.eps)
    x = _lin(x, sd, prefix + "post_extract_proj")
    wc = beats_pos_conv_weight(sd, prefix)
    xc = F.conv1d(x.transpose(1, 2), wc, sd[prefix + "encoder.pos_conv.0.bias"], padding=cfg.conv_pos // 2,
                  groups=cfg.conv_groups)
    if cfg.conv_pos % 2 == 0:
        xc = xc[:, :, :-1]
    x = x + F.gelu(xc).transpose(1, 2)
    x = _ln(x, sd, prefix + "encoder.layer_norm", cfg.eps)
    pos_bias = beats_position_bias(sd, prefix, x.shape[1], cfg)
    for i in range(cfg.layers):
        lp = f"{prefix}encoder.layers.{i}."
        a = beats_attention(sd, lp + "self_attn.", x, pos_bias, cfg)
        x = _ln(x * cfg.alpha + a, sd, lp + "self_attn_layer_norm", cfg.eps)
        h = F.gelu(_lin(x, sd, lp + "fc1").float()).to(x.dtype)
        h = _lin(h, sd, lp + "fc2")
        x = _ln(x * cfg.alpha + h, sd, lp + "final_layer_norm", cfg.eps)
    return x


def audio_encoder(sd: SD, audio: torch.Tensor, cfg: BeatsCfg) -> torch.Tensor:
    """AudioEncoder.forward 4-D branch (models/multimodal_encoder.py:174-186): (b,t,L,128) -> (b,t,n,768)."""
    b, t, L, d = audio.shape
    f = beats_extract_features(sd, "model.audio_encoder.audio_encoder.", audio.reshape(b * t, L, d), cfg)
    return f.reshape(b, t, f.shape[1], f.shape[2])


# ----------------------------------------------------------------------------------------------------------------
# Decoder: LLaMA / Qwen2 with hyper-LoRA linears
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class DecoderCfg:
    hidden: int = 4096
    inter: int = 11008
    layers: int = 32
    heads: int = 32
    kv_heads: int = 32
    head_dim: int = 128
    vocab: int = 32017
    rope_theta: float = 10000.0
    eps: float = 1e-6
    qkv_bias: bool = False  # Qwen2: True (models/qwen/modeling_qwen2.py:234-237)
    lora_r: int = 8
    lora_alpha: int = 16
    lora_nums: int = 3

    @property
    def scaling(self) -> float:
        return self.lora_alpha / self.lora_r


def hyper_lora_linear(x: torch.Tensor, sd: SD, name: str, cfg: DecoderCfg, bias: bool = False) -> torch.Tensor:
    """hyper-LoRA Linear.forward (peft_hyper/tuners/lora.py:338-369): y = xW^T (+b) + sum_i softmax_i(xR^T) *
    B_i(A x) * alpha/r, router softmax in fp32 (:347); dropout is identity in eval."""
    y = F.linear(x, sd[name + ".weight"], sd.get(name + ".bias") if bias else None)
    if name + ".lora_A.weight" not in sd:
        return y
    route = torch.softmax(F.linear(x, sd[name + ".lora_route.weight"]), dim=-1, dtype=torch.float32).to(y.dtype)
    u = F.linear(x, sd[name + ".lora_A.weight"])
    for i in range(cfg.lora_nums):
        y = y + route[..., i:i + 1] * F.linear(u, sd[f"{name}.lora_B{i}.weight"]) * cfg.scaling
    return y


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm (models/modeling_llama.py:103-117): fp32 statistics, weight * x.to(input dtype)."""
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(x.dtype)


def rope_cos_sin(positions: torch.Tensor, head_dim: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaRotaryEmbedding (models/modeling_llama.py:123-156): fp32 cos/sin of pos * theta^(-2i/d), duplicated."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    fr = positions.float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_rotary_pos_emb (models/modeling_llama.py:204-236); x (b,h,s,d), cos/sin (s,d) cast to x.dtype."""
    return x * cos.to(x.dtype) + _rotate_half(x) * sin.to(x.dtype)


@dataclass
class KVCache:
    k: List[torch.Tensor] = field(default_factory=list)  # per layer (b, kv_heads, ctx, hd)
    v: List[torch.Tensor] = field(default_factory=list)


def decoder_forward(sd: SD, x: torch.Tensor, cfg: DecoderCfg, cache: Optional[KVCache] = None,
                    collect_hidden: bool = False):
    """LlamaModel/Qwen2Model forward over inputs_embeds x (b, s, D) appended after `cache` (pre-norm residual
    layers models/modeling_llama.py:765-837, attention :286-464 with fp32 softmax :433 and 1/sqrt(hd) :417, causal
    mask, GQA repeat_kv models/qwen/modeling_qwen2.py:190-199, SwiGLU MLP :239-271, final norm :1119).
    All-ones attention mask and positions past..past+s-1 — what HF generate builds for inputs_embeds-only input
    (SURVEY §3.2 note).  Returns (final-normed hidden (b,s,D), cache[, per-layer hiddens])."""
    b, s, D = x.shape
    H, KV, hd = cfg.heads, cfg.kv_heads, cfg.head_dim
    if cache is None:
        cache = KVCache()
    past = cache.k[0].shape[2] if cache.k else 0
    cos, sin = rope_cos_sin(torch.arange(past, past + s), hd, cfg.rope_theta)
    mask = torch.full((s, past + s), float("-inf")).triu(diagonal=past + 1)
    hiddens = []
    for i in range(cfg.layers):
        lp = f"model.layers.{i}."
        if collect_hidden:
            hiddens.append(x)
        h = rms_norm(x, sd[lp + "input_layernorm.weight"], cfg.eps)
        q = hyper_lora_linear(h, sd, lp + "self_attn.q_proj", cfg, cfg.qkv_bias).view(b, s, H, hd).transpose(1, 2)
        k = hyper_lora_linear(h, sd, lp + "self_attn.k_proj", cfg, cfg.qkv_bias).view(b, s, KV, hd).transpose(1, 2)
        v = hyper_lora_linear(h, sd, lp + "self_attn.v_proj", cfg, cfg.qkv_bias).view(b, s, KV, hd).transpose(1, 2)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        if len(cache.k) > i:
            k = torch.cat([cache.k[i], k], dim=2)
            v = torch.cat([cache.v[i], v], dim=2)
            cache.k[i], cache.v[i] = k, v
        else:
            cache.k.append(k)
            cache.v.append(v)
        kr = k.repeat_interleave(H // KV, dim=1) if KV != H else k
        vr = v.repeat_interleave(H // KV, dim=1) if KV != H else v
        a = torch.matmul(q, kr.transpose(-1, -2)) / math.sqrt(hd) + mask.to(q.dtype)
        a = torch.softmax(a, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(a, vr).transpose(1, 2).reshape(b, s, H * hd)
        x = x + hyper_lora_linear(o, sd, lp + "self_attn.o_proj", cfg)
        h = rms_norm(x, sd[lp + "post_attention_layernorm.weight"], cfg.eps)
        g = hyper_lora_linear(h, sd, lp + "mlp.gate_proj", cfg)
        u = hyper_lora_linear(h, sd, lp + "mlp.up_proj", cfg)
        x = x + hyper_lora_linear(F.silu(g) * u, sd, lp + "mlp.down_proj", cfg)
    if collect_hidden:
        hiddens.append(x)
    x = rms_norm(x, sd["model.norm.weight"], cfg.eps)
    return (x, cache, hiddens) if collect_hidden else (x, cache)


def lm_head(sd: SD, h: torch.Tensor) -> torch.Tensor:
    """lm_head, logits in fp32 (models/unified_llama.py:35; models/modeling_llama.py:1254-1261)."""
    return F.linear(h, sd["lm_head.weight"]).float()


# ----------------------------------------------------------------------------------------------------------------
# The path: prepare_multimodal_inputs -> prefill -> greedy decode
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class CrabCfg:
    decoder: DecoderCfg = field(default_factory=DecoderCfg)
    clip: ClipCfg = field(default_factory=ClipCfg)
    beats: BeatsCfg = field(default_factory=BeatsCfg)
    qformer: QformerCfg = field(default_factory=QformerCfg)
    select_layers: Tuple[int, ...] = (14, 22, 23)
    image_tokens: int = 256
    n_query: int = 32
    base_vocab: int = 32000  # len(tokenizer) before initialize_MM_tokenizer
    pad_token_id: int = 0

    @property
    def special_ids(self) -> Dict[str, int]:
        """initialize_MM_tokenizer (models/unified_arch.py:409-459): 11 special + 6 <mask_i> tokens."""
        names = ["<image>", "<image_start>", "<image_end>", "<video>", "<video_start>", "<video_end>", "<audio>",
                 "<audio_start>", "<audio_end>", "<mask_start>", "<mask_end>"] + [f"<mask_{i}>" for i in range(6)]
        return {t: self.base_vocab + i for i, t in enumerate(names)}


def encode_video(sd: SD, video: torch.Tensor, cfg: CrabCfg) -> torch.Tensor:
    """UnifiedMetaModel.encode_video (models/unified_arch.py:144-149), last tap only — the one the path consumes
    (:290).  video (t,3,H,W) -> (t*32, d_model)."""
    feats = visual_encoder(sd, video.unsqueeze(0), cfg.clip, cfg.select_layers)
    return vl_projector(sd, feats[-1], cfg.qformer, cfg.image_tokens, cfg.n_query)[0]


def encode_audio(sd: SD, audio: torch.Tensor, cfg: CrabCfg) -> torch.Tensor:
    """UnifiedMetaModel.encode_audio (models/unified_arch.py:152-155). audio (t,L,128) -> (t*32, d_model)."""
    return al_projector(sd, audio_encoder(sd, audio.unsqueeze(0), cfg.beats), cfg.qformer, cfg.n_query)[0]


def prepare_multimodal_inputs(sd: SD, batch_input_ids: List[torch.Tensor], batch_X_modals: List[dict], cfg: CrabCfg):
    """UnifiedMetaForCausalLM.prepare_multimodal_inputs (models/unified_arch.py:217-406), generation branch:
    splice modality embeddings at placeholder ids, left-pad with pad-token embeddings, mask / position ids."""
    ids = cfg.special_ids
    keys = {ids["<image>"]: "<image>", ids["<video>"]: "<video>", ids["<audio>"]: "<audio>"}
    emb = sd["model.embed_tokens.weight"]
    seqs = []
    for input_ids, X in zip(batch_input_ids, batch_X_modals):
        segs, pre = [], 0
        for idx in [i for i, t in enumerate(input_ids.tolist()) if t in keys]:
            segs.append(emb[input_ids[pre:idx]])
            key = keys[int(input_ids[idx])]
            if key == "<audio>":
                segs.append(encode_audio(sd, X[key], cfg))
            else:
                segs.append(encode_video(sd, X[key], cfg))
            pre = idx + 1
        segs.append(emb[input_ids[pre:]])
        seqs.append(torch.cat(segs, dim=0))
    L = max(s.shape[0] for s in seqs)
    embeds, masks = [], []
    for s in seqs:
        pad = emb[torch.full((L - s.shape[0],), cfg.pad_token_id, dtype=torch.long)]
        embeds.append(torch.cat([pad, s], dim=0))
        masks.append(torch.cat([torch.zeros(L - s.shape[0], dtype=torch.int32), torch.ones(s.shape[0], dtype=torch.int32)]))
    embeds, masks = torch.stack(embeds), torch.stack(masks)
    pos = torch.cumsum(masks, dim=-1) - 1
    pos[pos == -1] = 0
    return {"inputs_embeds": embeds, "attention_mask": masks, "position_ids": pos}


def greedy_generate(sd: SD, inputs_embeds: torch.Tensor, cfg: DecoderCfg, max_new_tokens: int,
                    teacher_tokens: Optional[torch.Tensor] = None):
    """UnifiedForCausalLM.generate -> HF greedy loop with inputs_embeds only (models/unified_llama.py:244-267,
    47-161): prefill over the embeddings, then one token per step through embed_tokens (:125-127), KV cache on.
    EOS disabled (fixed-length runs).  Returns (ids (b, n), per-step last-position logits (n, b, V))."""
    inputs_embeds = inputs_embeds.to(sd["model.embed_tokens.weight"].dtype)  # models/unified_llama.py:149
    h, cache = decoder_forward(sd, inputs_embeds, cfg)
    logits = lm_head(sd, h[:, -1])
    ids, all_logits = [], []
    for step in range(max_new_tokens):
        all_logits.append(logits)
        nxt = logits.argmax(dim=-1)
        ids.append(nxt)
        if step + 1 == max_new_tokens:
            break
        feed = nxt if teacher_tokens is None else teacher_tokens[:, step]
        h, cache = decoder_forward(sd, sd["model.embed_tokens.weight"][feed].unsqueeze(1), cfg, cache)
        logits = lm_head(sd, h[:, -1])
    return torch.stack(ids, dim=1), torch.stack(all_logits, dim=0)
