"""Host-side orchestration of Crab's hot path on one B200: weight packing (reference state-dict names -> packed
bf16 device buffers) and the kernel sequences for the encoders, the Q-Former bridges, decoder prefill and the
CUDA-graph decode step.  Python here is tensor plumbing only: every arithmetic step is a call into
libcrab_b200.so through `crab_b200.ops` (there is no torch math on the data path, and no CPU fallback).

Reference call stack being replaced:
  UnifiedForCausalLM.generate (models/unified_llama.py:244-267)
    -> prepare_multimodal_inputs (models/unified_arch.py:217-406)
         -> encode_video / encode_audio (models/unified_arch.py:113-155; models/multimodal_encoder.py)
    -> HF greedy loop over LlamaForCausalLM / Qwen2ForCausalLM with hyper-LoRA linears
       (peft_hyper/tuners/lora.py:338-369)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

SD = Dict[str, torch.Tensor]
DEFAULT_DECODE_MODE = "skinny"
DEFAULT_PDL_PLAN = "19,17"  # streaming GEMM + light kernels launch programmatically; attention / RoPE stay plain launches


# --------------------------------------------------------------------------------------------------------------
# configuration (mirrors the fields the reference reads from LlamaConfig/Qwen2Config, the CLIP/BEATs/Bert configs
# and configs/unified_config.py)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class DecoderConfig:
    hidden: int = 4096
    inter: int = 11008
    layers: int = 32
    heads: int = 32
    kv_heads: int = 32
    head_dim: int = 128
    vocab: int = 32017
    rope_theta: float = 10000.0
    eps: float = 1e-6
    qkv_bias: bool = False
    lora_r: int = 8
    lora_alpha: int = 16
    lora_nums: int = 3


@dataclass
class ClipConfig:
    hidden: int = 1024
    inter: int = 4096
    heads: int = 16
    layers: int = 24
    patch: int = 14
    image: int = 224
    eps: float = 1e-5
    image_mean: Tuple[float, float, float] = (0.48145466, 0.4578275, 0.40821073)  # CLIPImageProcessor (OpenAI CLIP)
    image_std: Tuple[float, float, float] = (0.26862954, 0.26130258, 0.27577711)


@dataclass
class BeatsConfig:
    patch: int = 16
    embed: int = 512
    dim: int = 768
    ffn: int = 3072
    heads: int = 12
    layers: int = 12
    conv_pos: int = 128
    conv_groups: int = 16
    num_buckets: int = 320
    max_distance: int = 800
    eps: float = 1e-5


@dataclass
class QformerConfig:
    hidden: int = 768
    heads: int = 12
    inter: int = 3072
    layers: int = 2
    eps: float = 1e-12


@dataclass
class CrabConfig:
    decoder: DecoderConfig = field(default_factory=DecoderConfig)
    clip: ClipConfig = field(default_factory=ClipConfig)
    beats: BeatsConfig = field(default_factory=BeatsConfig)
    qformer: QformerConfig = field(default_factory=QformerConfig)
    select_layers: Tuple[int, ...] = (14, 22, 23)
    n_query: int = 32
    pad_token_id: int = 0
    max_ctx: int = 1280
    special_ids: Dict[str, int] = field(default_factory=dict)  # from initialize_MM_tokenizer
    low_res_mask_size: int = 112                               # init_multimodal_modules(low_res_mask_size=) -> SegModule


def _bf(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _pad_cols(w: torch.Tensor, cols: int) -> torch.Tensor:
    if w.shape[1] == cols:
        return w.contiguous()
    out = torch.zeros((w.shape[0], cols), dtype=w.dtype, device=w.device)
    out[:, : w.shape[1]] = w
    return out


class _Lin:
    """A packed dense layer: bf16 weight [N, K(+ext)], fp32 bias."""

    __slots__ = ("w", "b")

    def __init__(self, w, b=None):
        self.w, self.b = w, b


# --------------------------------------------------------------------------------------------------------------
# Encoders + bridges
# --------------------------------------------------------------------------------------------------------------
class _QformerWeights:
    def __init__(self, sd: SD, p: str, cfg: QformerConfig, dev):
        b = p + "bert."
        self.emb_ln = (_f32(sd[b + "embeddings.LayerNorm.weight"], dev), _f32(sd[b + "embeddings.LayerNorm.bias"], dev))
        self.layers = []
        for i in range(cfg.layers):
            lp = f"{b}encoder.layer.{i}."

            def lin(n):
                return _Lin(_bf(sd[lp + n + ".weight"], dev), _f32(sd[lp + n + ".bias"], dev))

            def ln(n):
                return (_f32(sd[lp + n + ".weight"], dev), _f32(sd[lp + n + ".bias"], dev))

            def cat(names):
                return _Lin(_bf(torch.cat([sd[lp + n + ".weight"] for n in names], 0), dev),
                            _f32(torch.cat([sd[lp + n + ".bias"] for n in names], 0), dev))

            self.layers.append(dict(
                self_qkv=cat(["attention.self.query", "attention.self.key", "attention.self.value"]),
                self_out=lin("attention.output.dense"), self_ln=ln("attention.output.LayerNorm"),
                cross_q=lin("crossattention.self.query"),
                cross_kv=cat(["crossattention.self.key", "crossattention.self.value"]),
                cross_out=lin("crossattention.output.dense"), cross_ln=ln("crossattention.output.LayerNorm"),
                ffn_in=lin("intermediate_query.dense"), ffn_out=lin("output_query.dense"),
                ffn_ln=ln("output_query.LayerNorm")))


class CrabEngine:
    def __init__(self, sd: SD, cfg: CrabConfig, device: Optional[torch.device] = None, load_encoders: bool = True,
                 decode_packed: bool = True):
        if not torch.cuda.is_available():
            raise ops._l.CrabError("crab_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        ops.init(self.dev.index or 0)
        self.cfg = cfg
        self.decode_packed = decode_packed
        # decode step for <= 32 rows (CRAB_DECODE_MODE): "rows" = round-1 organisation, a cluster row kernel (RMSNorm + hyper-LoRA
        # pre-pass) in front of every weight-streaming GEMM (8 launches per layer); "skinny" = one fused launch per linear, the norm
        # as an epilogue scale and the pre-pass by an in-launch statistics cluster (5 per layer); "chain" = the persistent
        # o -> gate/up -> down -> qkv kernel (2 per layer).  Measured on B200 at bs 32: see DEFAULT_DECODE_MODE.
        self.decode_mode = os.environ.get("CRAB_DECODE_MODE", DEFAULT_DECODE_MODE)
        sd = {(k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v for k, v in sd.items()}
        self._pack_decoder(sd)
        self.seg = None
        if "model.seg_module.no_mask_embed.weight" in sd:   # segmentation branch present: the mask head (crab_b200/seg.py)
            from .seg import SegHead

            self.seg = SegHead(sd, self.dev, prefix="model.seg_module", grid=cfg.clip.image // cfg.clip.patch,
                               low_res=cfg.low_res_mask_size)
        self.has_encoders = load_encoders and ("model.vl_projector.visual_ln.weight" in sd)
        if self.has_encoders:
            self._pack_clip(sd)
            self._pack_beats(sd)
            self._pack_bridges(sd)
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._tap_layers: Tuple[int, ...] = ()   # CLIP layers whose patch tokens prepare_inputs(want_image_taps=True) keeps
        self._last_taps: Dict[int, torch.Tensor] = {}
        # Programmatic dependent launch over the decode chain: (mask for the chain, mask for the kernel that follows the
        # decode attention).  Env CRAB_PDL_PLAN="chain,after_attn" overrides; see profiles/r02_pdl_plans.txt for the A/B.
        plan = os.environ.get("CRAB_PDL_PLAN", DEFAULT_PDL_PLAN).split(",")
        self.pdl_chain, self.pdl_after_attn = int(plan[0]), int(plan[-1])
        # K-split (cluster size) per decode linear, 0 = the library's choice; env CRAB_SKINNY_SPLITS="qkv:2,o:8,gu:1,d:8" overrides
        # o_proj: 4 rather than the library's 8 — measured inside the step (tools/sweep_splits.sh): 14-15 us instead of 20.6 us per
        # launch; its input arrives from the attention kernel all at once, so a shorter DSMEM reduce beats the extra CTAs
        # down_proj: 4 (library: 8) — with the statistics clusters and the st.async exchange the timeline shows 20.5 us against
        # 25.4 us (tools/sweep_splits_timeline.sh, profiles/r04_split_sweep_timeline.txt)
        self.skinny_splits = {"qkv": 0, "o": 4, "gu": 0, "d": 4}
        for kv in filter(None, os.environ.get("CRAB_SKINNY_SPLITS", "").split(",")):
            k_, v_ = kv.split(":")
            self.skinny_splits[k_] = int(v_)
        # flag slots in a ring (each launch zeroes its predecessor's) and several statistics clusters per launch (0 = library's choice)
        # CUDA graphs around launch-bound call sites outside the decode loop (encoders of <= 32 frames / segments, prefill of <= 2048 rows)
        self.call_graphs = os.environ.get("CRAB_CALL_GRAPHS", "1") != "0"
        self._call_graphs = {}
        self.flag_ring = os.environ.get("CRAB_FLAG_RING", "1") != "0"
        self.stats_clusters = int(os.environ.get("CRAB_STATS_CLUSTERS", "0"))
        # L2 prefetch of the next linear's weights from the tail of each decode linear (MB per launch, 0 = off)
        self.prefetch_mb = int(os.environ.get("CRAB_PREFETCH_MB", "0"))
        # K-split (= thread-block-cluster size) of the persistent decode GEMM chain
        self.chain_cluster = int(os.environ.get("CRAB_CHAIN_CLUSTER", "4"))
        # decode step: RoPE + KV append + o_proj LoRA pre-pass inside the attention kernel (8 launches per layer, not 10)
        self.fuse_decode_attn = os.environ.get("CRAB_DECODE_FUSE", "1") != "0"
        self.gqa_decode_tc = os.environ.get("CRAB_GQA_DECODE_TC", "1") != "0"
        self._graph = None
        self._graph_bs = None
        self._use_graph = False
        self.past_dev = None
        self.len_dev = None
        self.cur_len = 0
        self.k_cache: List[torch.Tensor] = []
        self.v_cache: List[torch.Tensor] = []

    # ---- buffers -------------------------------------------------------------------------------------------
    def _buf(self, name: str, shape, dtype=torch.bfloat16, zero=False) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), device=self.dev, dtype=dtype)
            self._bufs[key] = t
        return t

    # ---- decoder weights -------------------------------------------------------------------------------------
    def _graph_replay(self, key, static_inputs: Sequence[torch.Tensor], inputs: Sequence[torch.Tensor], fn):
        """Launch-bound call sites (a single sample's encoders / prefill are ~1.2 k launches of a few microseconds): run
        `fn(*static_inputs)` as ONE CUDA-graph replay, captured once per `key` (shapes); `inputs` are copied into the static
        buffers first.  Returns what fn returned at capture time (tensors in the graph's memory pool: valid until the next replay
        of the same key).  At most 8 graphs are kept; beyond that the call runs eagerly."""
        ent = self._call_graphs.get(key)
        if ent is None:
            if len(self._call_graphs) >= 8:
                for d_, s_ in zip(static_inputs, inputs):
                    d_.copy_(s_)
                return fn(*static_inputs)
            for d_, s_ in zip(static_inputs, inputs):
                d_.copy_(s_)
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(*static_inputs)                      # warm-up outside capture: kernel attributes, lazily built tables, allocator
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            for d_, s_ in zip(static_inputs, inputs):   # fn may consume its input in place
                d_.copy_(s_)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                out = fn(*static_inputs)
            kernels = ops.launch_count() - n0
            ops.count_launches(-kernels)                # capture launches nothing
            ent = self._call_graphs[key] = (g, out, kernels)
        g, out, kernels = ent
        for d_, s_ in zip(static_inputs, inputs):
            d_.copy_(s_)
        g.replay()
        ops.count_launches(kernels)
        return out

    def _small_call_graphs(self) -> bool:
        return self.call_graphs and self.dev.type == "cuda" and ops._timer is None and not torch.cuda.is_current_stream_capturing()

    def _pack_decoder(self, sd: SD):
        c, dev = self.cfg.decoder, self.dev
        D, F, H, KV, hd = c.hidden, c.inter, c.heads, c.kv_heads, c.head_dim
        assert F % 64 == 0 and D % 8 == 0 and hd in (64, 128)
        self.lora = "model.layers.0.self_attn.q_proj.lora_A.weight" in sd
        self.embed = _bf(sd["model.embed_tokens.weight"], dev)
        self.final_norm = _f32(sd["model.norm.weight"], dev)
        V = sd["lm_head.weight"].shape[0]
        self.vocab = V
        self.vocab_pad = (V + 7) // 8 * 8
        lm = torch.zeros((self.vocab_pad, D), dtype=torch.bfloat16, device=dev)
        lm[:V] = _bf(sd["lm_head.weight"], dev)
        self.lm_head = lm
        # decode / last-position head: final-norm gamma folded into the streaming copy (the chain applies rstd in its epilogue)
        self.lm_head_c = self.lm_head_p = None
        if self.decode_packed and self.decode_mode == "rows":
            self.lm_head_p = ops.pack_skinny_weight(lm)
        elif self.decode_packed:
            lmf = lm.clone()
            lmf[:] = (lm.float() * self.final_norm[None, :]).to(torch.bfloat16)
            self.lm_head_c = ops.pack_skinny_weight(lmf)
            del lmf
        self.scaling = c.lora_alpha / c.lora_r
        nl, r = c.lora_nums, c.lora_r
        zw = nl * r  # 24 z columns per linear
        self.EXT_QKV, self.EXT_O, self.EXT_GU, self.EXT_D = 96, 32, 64, 32
        self.layers = []

        def put(dst, src):  # packed buffers are assembled directly in bf16 on the device, slice by slice
            dst.copy_(src.detach().to(device=dev, dtype=torch.bfloat16, non_blocking=True))

        def bcat(dst, name):  # [out, 24] = [B0 | B1 | B2]
            for j in range(nl):
                put(dst[:, j * r:(j + 1) * r], sd[f"{name}.lora_B{j}.weight"])

        def ra(names):  # rows [R (3); A (8)] per linear
            out = torch.empty((len(names) * (nl + r), sd[names[0] + ".lora_A.weight"].shape[1]), device=dev, dtype=torch.bfloat16)
            for j, n in enumerate(names):
                put(out[j * (nl + r):j * (nl + r) + nl], sd[n + ".lora_route.weight"])
                put(out[j * (nl + r) + nl:(j + 1) * (nl + r)], sd[n + ".lora_A.weight"])
            return out

        def zeros(rows, cols):
            return torch.zeros((rows, cols), device=dev, dtype=torch.bfloat16)

        for i in range(c.layers):
            lp = f"model.layers.{i}."
            at, ml = lp + "self_attn.", lp + "mlp."
            L = {}
            nq, nk = H * hd, KV * hd
            wq = zeros(nq + 2 * nk, D + self.EXT_QKV)
            put(wq[:nq, :D], sd[at + "q_proj.weight"])
            put(wq[nq:nq + nk, :D], sd[at + "k_proj.weight"])
            put(wq[nq + nk:, :D], sd[at + "v_proj.weight"])
            if self.lora:
                bcat(wq[:nq, D:D + zw], at + "q_proj")
                bcat(wq[nq:nq + nk, D + zw:D + 2 * zw], at + "k_proj")
                bcat(wq[nq + nk:, D + 2 * zw:D + 3 * zw], at + "v_proj")
                L["ra_qkv"] = ra([at + "q_proj", at + "k_proj", at + "v_proj"])
            L["wqkv"] = wq
            L["bqkv"] = (torch.cat([_f32(sd[at + "q_proj.bias"], dev), _f32(sd[at + "k_proj.bias"], dev),
                                    _f32(sd[at + "v_proj.bias"], dev)]) if c.qkv_bias else None)
            wo = zeros(D, nq + self.EXT_O)
            put(wo[:, :nq], sd[at + "o_proj.weight"])
            if self.lora:
                bcat(wo[:, nq:nq + zw], at + "o_proj")
                L["ra_o"] = ra([at + "o_proj"])
            L["wo"] = wo
            # gate/up rows interleaved in [64 gate | 64 up] groups so one output tile holds both halves (SWIGLU)
            wgu = zeros(2 * F, D + self.EXT_GU).view(F // 64, 2, 64, D + self.EXT_GU)
            put(wgu[:, 0, :, :D], sd[ml + "gate_proj.weight"].view(F // 64, 64, D))
            put(wgu[:, 1, :, :D], sd[ml + "up_proj.weight"].view(F // 64, 64, D))
            if self.lora:
                for j in range(nl):
                    put(wgu[:, 0, :, D + j * r:D + (j + 1) * r], sd[f"{ml}gate_proj.lora_B{j}.weight"].view(F // 64, 64, r))
                    put(wgu[:, 1, :, D + zw + j * r:D + zw + (j + 1) * r], sd[f"{ml}up_proj.lora_B{j}.weight"].view(F // 64, 64, r))
                L["ra_gu"] = ra([ml + "gate_proj", ml + "up_proj"])
            L["wgu"] = wgu.view(2 * F, D + self.EXT_GU)
            wd = zeros(D, F + self.EXT_D)
            put(wd[:, :F], sd[ml + "down_proj.weight"])
            if self.lora:
                bcat(wd[:, F:F + zw], ml + "down_proj")
                L["ra_d"] = ra([ml + "down_proj"])
            L["wd"] = wd
            L["ln1"] = _f32(sd[lp + "input_layernorm.weight"], dev)
            L["ln2"] = _f32(sd[lp + "post_attention_layernorm.weight"], dev)
            if self.decode_packed and self.decode_mode == "rows":
                lo = self.lora
                L["wqkv_p"] = ops.pack_skinny_weight(wq, k=D + (self.EXT_QKV if lo else 0))
                L["wo_p"] = ops.pack_skinny_weight(wo, k=nq + (self.EXT_O if lo else 0))
                L["wgu_p"] = ops.pack_skinny_weight(L["wgu"], k=D + (self.EXT_GU if lo else 0), swiglu=True)
                L["wd_p"] = ops.pack_skinny_weight(wd, k=F + (self.EXT_D if lo else 0))
            elif self.decode_packed:
                # second copy of the decode-step weights in the streaming layout (contiguous pre-swizzled 16 KB tile
                # blocks): the prefill GEMM wants row-major K-extended rows, the M<=32 chain wants sequential HBM.
                # The chain applies RMSNorm as rstd[b] in its epilogue, so gamma is folded into the columns here; the
                # hyper-LoRA router/A rows go into the chain's statistics stream (gamma folded likewise).
                lo = self.lora

                def fold(w, K, gamma):
                    t = w.clone()
                    t[:, :K] = (w[:, :K].float() * gamma[None, :]).to(torch.bfloat16)
                    return t

                L["wqkv_c"] = ops.pack_skinny_weight(fold(wq, D, L["ln1"]), k=D + (self.EXT_QKV if lo else 0))
                L["wo_c"] = ops.pack_skinny_weight(wo, k=nq + (self.EXT_O if lo else 0))
                L["wgu_c"] = ops.pack_skinny_weight(fold(L["wgu"], D, L["ln2"]), k=D + (self.EXT_GU if lo else 0), swiglu=True)
                L["wd_c"] = ops.pack_skinny_weight(wd, k=F + (self.EXT_D if lo else 0))
                if lo:
                    L["st_qkv"] = ops.pack_chain_stats(L["ra_qkv"], L["ln1"])
                    L["st_o"] = ops.pack_chain_stats(L["ra_o"])
                    L["st_gu"] = ops.pack_chain_stats(L["ra_gu"], L["ln2"])
                    L["st_d"] = ops.pack_chain_stats(L["ra_d"])
            self.layers.append(L)
        self.rope = ops.rope_table(self.cfg.max_ctx, hd, c.rope_theta, dev)

    # ---- CLIP --------------------------------------------------------------------------------------------------
    def _pack_clip(self, sd: SD):
        c, dev = self.cfg.clip, self.dev
        p = "model.visual_encoder.vision_tower.vision_model."
        # hidden_states[k] = output of layer k (index 0 = embeddings); the reference also ships negative indices
        # ([-11, -2, -1], configs/unified_config.py): normalise against the 25-entry tuple like Python indexing does
        nh = c.layers + 1
        sel = tuple(k if k >= 0 else k + nh for k in self.cfg.select_layers)
        if any(k < 1 or k >= nh for k in sel):
            raise ValueError(f"select_layers {self.cfg.select_layers} out of range for a {c.layers}-layer ViT")
        self.cfg.select_layers = sel
        self.clip_layers_run = max(sel)
        kp = 3 * c.patch * c.patch
        self.clip_kpad = (kp + 7) // 8 * 8
        self.clip_patch_w = _bf(_pad_cols(sd[p + "embeddings.patch_embedding.weight"].reshape(c.hidden, kp).float(), self.clip_kpad), dev)
        self.clip_cls = _f32(sd[p + "embeddings.class_embedding"], dev)
        self.clip_pos = _f32(sd[p + "embeddings.position_embedding.weight"], dev)
        self.clip_pre_ln = (_f32(sd[p + "pre_layrnorm.weight"], dev), _f32(sd[p + "pre_layrnorm.bias"], dev))
        self.clip_layers = []
        for i in range(self.clip_layers_run):
            lp = f"{p}encoder.layers.{i}."
            a = lp + "self_attn."
            self.clip_layers.append(dict(
                ln1=(_f32(sd[lp + "layer_norm1.weight"], dev), _f32(sd[lp + "layer_norm1.bias"], dev)),
                ln2=(_f32(sd[lp + "layer_norm2.weight"], dev), _f32(sd[lp + "layer_norm2.bias"], dev)),
                qkv=_Lin(_bf(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0), dev),
                         _f32(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0), dev)),
                out=_Lin(_bf(sd[a + "out_proj.weight"], dev), _f32(sd[a + "out_proj.bias"], dev)),
                fc1=_Lin(_bf(sd[lp + "mlp.fc1.weight"], dev), _f32(sd[lp + "mlp.fc1.bias"], dev)),
                fc2=_Lin(_bf(sd[lp + "mlp.fc2.weight"], dev), _f32(sd[lp + "mlp.fc2.bias"], dev))))

    # ---- BEATs -------------------------------------------------------------------------------------------------
    def _pack_beats(self, sd: SD):
        c, dev = self.cfg.beats, self.dev
        p = "model.audio_encoder.audio_encoder."
        self.beats_patch_w = _bf(sd[p + "patch_embedding.weight"].reshape(c.embed, c.patch * c.patch), dev)
        self.beats_ln0 = (_f32(sd[p + "layer_norm.weight"], dev), _f32(sd[p + "layer_norm.bias"], dev))
        self.beats_proj = _Lin(_bf(sd[p + "post_extract_proj.weight"], dev), _f32(sd[p + "post_extract_proj.bias"], dev))
        pc = p + "encoder.pos_conv.0."
        if pc + "weight_g" in sd:
            g, v = sd[pc + "weight_g"].float(), sd[pc + "weight_v"].float()
        else:
            g, v = sd[pc + "parametrizations.weight.original0"].float(), sd[pc + "parametrizations.weight.original1"].float()
        # weight_norm(dim=2) folded once at load time (models/beats/backbone.py:45)
        self.beats_conv_w = (v * (g / v.norm(dim=(0, 1), keepdim=True))).to(dev)  # [C, C/G, K] fp32
        self.beats_conv_b = _f32(sd[pc + "bias"], dev)
        self.beats_enc_ln = (_f32(sd[p + "encoder.layer_norm.weight"], dev), _f32(sd[p + "encoder.layer_norm.bias"], dev))
        self.beats_rel = sd[p + "encoder.layers.0.self_attn.relative_attention_bias.weight"].float().to(dev)  # [buckets, H]
        self.beats_alpha = math.pow(2 * c.layers, 0.25)
        self.beats_layers = []
        for i in range(c.layers):
            lp = f"{p}encoder.layers.{i}."
            a = lp + "self_attn."
            self.beats_layers.append(dict(
                qkv=_Lin(_bf(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0), dev),
                         _f32(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0), dev)),
                out=_Lin(_bf(sd[a + "out_proj.weight"], dev), _f32(sd[a + "out_proj.bias"], dev)),
                grep_w=_f32(sd[a + "grep_linear.weight"], dev), grep_b=_f32(sd[a + "grep_linear.bias"], dev),
                grep_a=_f32(sd[a + "grep_a"].reshape(-1), dev),
                ln1=(_f32(sd[lp + "self_attn_layer_norm.weight"], dev), _f32(sd[lp + "self_attn_layer_norm.bias"], dev)),
                fc1=_Lin(_bf(sd[lp + "fc1.weight"], dev), _f32(sd[lp + "fc1.bias"], dev)),
                fc2=_Lin(_bf(sd[lp + "fc2.weight"], dev), _f32(sd[lp + "fc2.bias"], dev)),
                ln2=(_f32(sd[lp + "final_layer_norm.weight"], dev), _f32(sd[lp + "final_layer_norm.bias"], dev))))
        self._beats_T_cache: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def _beats_tables(self, T: int):
        """Per-sequence-length constants, built once: the (H,T,T) relative-position bias table
        (backbone.py:392-430) and the Toeplitz-expanded pos-conv weights, one [T*cg, T*cg] matrix per conv group:
        W[g][(t,co),(t',ci)] = w[g*cg+co, ci, t'-t+pad]  (zero outside the K taps)."""
        if T in self._beats_T_cache:
            return self._beats_T_cache[T]
        c, dev = self.cfg.beats, self.dev
        ctx = torch.arange(T, device=dev)[:, None]
        mem = torch.arange(T, device=dev)[None, :]
        rel = mem - ctx
        nb = c.num_buckets // 2
        buckets = (rel > 0).long() * nb
        rel = rel.abs()
        max_exact = nb // 2
        large = max_exact + (torch.log(rel.float() / max_exact) / math.log(c.max_distance / max_exact)
                             * (nb - max_exact)).long()
        large = torch.minimum(large, torch.full_like(large, nb - 1))
        buckets = buckets + torch.where(rel < max_exact, rel, large)
        table = self.beats_rel[buckets].permute(2, 0, 1).contiguous()  # [H, T, T] fp32
        G, K, pad = c.conv_groups, c.conv_pos, c.conv_pos // 2
        cg = c.dim // G
        w = self.beats_conv_w.view(G, cg, cg, K)  # [g, co, ci, k]
        kidx = (torch.arange(T, device=dev)[None, :] - torch.arange(T, device=dev)[:, None]) + pad  # [t, t']
        valid = (kidx >= 0) & (kidx < K)
        wk = w[:, :, :, kidx.clamp(0, K - 1)] * valid  # [g, co, ci, t, t']
        wexp = wk.permute(0, 3, 1, 4, 2).reshape(G, T * cg, T * cg).to(torch.bfloat16).contiguous()
        self._beats_T_cache[T] = (table, wexp)
        return table, wexp

    # ---- bridges -----------------------------------------------------------------------------------------------
    def _pack_bridges(self, sd: SD):
        dev, q = self.dev, self.cfg.qformer
        v, a = "model.vl_projector.", "model.al_projector."
        self.visual_ln = (_f32(sd[v + "visual_ln.weight"], dev), _f32(sd[v + "visual_ln.bias"], dev))
        self.audio_ln = (_f32(sd[a + "audio_ln.weight"], dev), _f32(sd[a + "audio_ln.bias"], dev))
        self.vq = _QformerWeights(sd, v + "visual_Qformer.", q, dev)
        self.aq = _QformerWeights(sd, a + "audio_Qformer.", q, dev)
        self.v_query = _bf(sd[v + "visual_query_tokens"].reshape(-1, q.hidden), dev)
        self.a_query = _bf(sd[a + "audio_query_tokens"].reshape(-1, q.hidden), dev)
        self.v_proj = [_Lin(_bf(sd[v + f"visual_proj.{i}.weight"], dev), _f32(sd[v + f"visual_proj.{i}.bias"], dev)) for i in (0, 2)]
        self.a_proj = [_Lin(_bf(sd[a + f"audio_proj.{i}.weight"], dev), _f32(sd[a + f"audio_proj.{i}.bias"], dev)) for i in (0, 2)]

    # ==========================================================================================================
    # forward pieces
    # ==========================================================================================================
    def clip_forward(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels fp32 (n,3,H,W) — or uint8 frames (n,H,W,3) straight from the video decoder — on device ->
        hidden_states[select_layers[-1]] as bf16 [n*tokens, D] (CLS kept)."""
        c = self.cfg.clip
        n = pixels.shape[0]
        g = (pixels.shape[1] if pixels.dtype == torch.uint8 else pixels.shape[2]) // c.patch
        tokens = g * g + 1
        D = c.hidden
        if pixels.dtype == torch.uint8:
            # raw decoded frames (n, H, W, 3): CLIPImageProcessor's rescale + normalise fused into the im2col
            patches = ops.patchify_u8(pixels, c.patch, self.clip_kpad, c.image_mean, c.image_std)
        else:
            patches = ops.patchify(pixels, c.patch, self.clip_kpad)
        pe = ops.gemm(patches, self.clip_patch_w)
        x = ops.clip_embed_ln(pe, self.clip_cls, self.clip_pos, *self.clip_pre_ln, n, tokens, D, c.eps)
        M = n * tokens
        h = self._buf("clip_h", (M, D))
        qkv = self._buf("clip_qkv", (M, 3 * D))
        o = self._buf("clip_o", (M, D))
        m = self._buf("clip_m", (M, c.inter))
        hd = D // c.heads
        self._last_taps = {}
        for li, L in enumerate(self.clip_layers):
            ops.layernorm(x, *L["ln1"], c.eps, out=h)
            ops.gemm(h, L["qkv"].w, bias=L["qkv"].b, out=qkv)
            ops.flash_attn(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B=n, H=c.heads, KVH=c.heads, Sq=tokens, Sk=tokens,
                           head_dim=hd, q_strides=(tokens * 3 * D, 3 * D, hd), k_strides=(tokens * 3 * D, 3 * D, hd),
                           v_strides=(tokens * 3 * D, 3 * D, hd), o_strides=(tokens * D, D, hd), scale=hd ** -0.5)
            ops.gemm(o, L["out"].w, bias=L["out"].b, residual=x, out=x)
            ops.layernorm(x, *L["ln2"], c.eps, out=h)
            ops.gemm(h, L["fc1"].w, bias=L["fc1"].b, act=ops.ACT_QUICK_GELU, out=m)
            ops.gemm(m, L["fc2"].w, bias=L["fc2"].b, residual=x, out=x)
            if (li + 1) in self._tap_layers:
                # hidden_states[li + 1] without the CLS row: the multi-scale image features the segmentation head reads
                # (VisualEncoder.feature_select, models/multimodal_encoder.py:52-65; unified_arch.py:245-247)
                self._last_taps[li + 1] = x.view(n, tokens, D)[:, 1:].reshape(n * (tokens - 1), D).clone()
        return x

    def _qformer(self, W: _QformerWeights, query: torch.Tensor, enc: torch.Tensor, n: int, enc_tokens: int,
                 enc_row0: int, enc_rows_per_item: int) -> torch.Tensor:
        """Q-Former over n items.  `enc` is a bf16 [n*enc_rows_per_item, width] matrix (already layer-normed);
        each item's keys are rows [enc_row0, enc_row0+enc_tokens) of its block."""
        q = self.cfg.qformer
        Dq, nq, hd = q.hidden, query.shape[0], q.hidden // q.heads
        x0 = ops.layernorm(query, *W.emb_ln, q.eps)  # identical for every item: computed once, then broadcast
        rows = torch.arange(nq, device=self.dev).repeat(n)
        x = torch.empty((n * nq, Dq), device=self.dev, dtype=torch.bfloat16)
        ops.gather_rows(x0, x, n * nq, Dq, src_rows=rows)
        width = enc.shape[1]
        for L in W.layers:
            qkv = ops.gemm(x, L["self_qkv"].w, bias=L["self_qkv"].b)
            o = torch.empty((n * nq, Dq), device=self.dev, dtype=torch.bfloat16)
            ops.flash_attn(qkv, qkv[:, Dq:], qkv[:, 2 * Dq:], o, B=n, H=q.heads, KVH=q.heads, Sq=nq, Sk=nq, head_dim=hd,
                           q_strides=(nq * 3 * Dq, 3 * Dq, hd), k_strides=(nq * 3 * Dq, 3 * Dq, hd),
                           v_strides=(nq * 3 * Dq, 3 * Dq, hd), o_strides=(nq * Dq, Dq, hd), scale=1 / math.sqrt(hd))
            y = ops.gemm(o, L["self_out"].w, bias=L["self_out"].b, residual=x)
            x = ops.layernorm(y, *L["self_ln"], q.eps)
            qc = ops.gemm(x, L["cross_q"].w, bias=L["cross_q"].b)
            kv = ops.gemm(enc, L["cross_kv"].w, bias=L["cross_kv"].b)  # [n*rows_per_item, 2*Dq]
            kv0 = kv[enc_row0:]
            ops.flash_attn(qc, kv0, kv0[:, Dq:], o, B=n, H=q.heads, KVH=q.heads, Sq=nq, Sk=enc_tokens, head_dim=hd,
                           q_strides=(nq * Dq, Dq, hd), k_strides=(enc_rows_per_item * 2 * Dq, 2 * Dq, hd),
                           v_strides=(enc_rows_per_item * 2 * Dq, 2 * Dq, hd), o_strides=(nq * Dq, Dq, hd),
                           scale=1 / math.sqrt(hd))
            y = ops.gemm(o, L["cross_out"].w, bias=L["cross_out"].b, residual=x)
            x = ops.layernorm(y, *L["cross_ln"], q.eps)
            hmid = ops.gemm(x, L["ffn_in"].w, bias=L["ffn_in"].b, act=ops.ACT_GELU)
            y = ops.gemm(hmid, L["ffn_out"].w, bias=L["ffn_out"].b, residual=x)
            x = ops.layernorm(y, *L["ffn_ln"], q.eps)
        return x

    def encode_video(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels fp32 (n_frames_total,3,H,W) -> bf16 [n_frames_total*32, d_model]  (VisualEncoder last tap ->
        VLProjector; models/unified_arch.py:144-149, models/multimodal_encoder.py:119-144)."""
        c = self.cfg.clip
        n = pixels.shape[0]
        tokens = ((pixels.shape[1] if pixels.dtype == torch.uint8 else pixels.shape[2]) // c.patch) ** 2 + 1
        x = self.clip_forward(pixels)
        f = ops.layernorm(x, *self.visual_ln, 1e-5)  # CLS rows are normalised too but never read
        qo = self._qformer(self.vq, self.v_query, f, n, tokens - 1, 1, tokens)
        h = ops.gemm(qo, self.v_proj[0].w, bias=self.v_proj[0].b, act=ops.ACT_GELU)
        return ops.gemm(h, self.v_proj[1].w, bias=self.v_proj[1].b)

    def beats_forward(self, fbank: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """fbank fp32 (n, L, 128) -> (bf16 [n*T, 768], T)   (models/beats/BEATs.py:134-182, backbone.py:109-273)."""
        c = self.cfg.beats
        n, Lf, nm = fbank.shape
        T = (Lf // c.patch) * (nm // c.patch)
        Dm, Hh = c.dim, c.heads
        hd = Dm // Hh
        patches = ops.patchify(fbank.reshape(n, 1, Lf, nm).contiguous(), c.patch, c.patch * c.patch)
        pe = ops.gemm(patches, self.beats_patch_w)
        pe = ops.layernorm(pe, *self.beats_ln0, c.eps)
        x = ops.gemm(pe, self.beats_proj.w, bias=self.beats_proj.b)  # [n*T, 768]
        table, wexp = self._beats_tables(T)
        G = c.conv_groups
        xg = ops.beats_group_pack(x, n, T, Dm, G)
        yg = torch.empty_like(xg)
        for g in range(G):
            ops.gemm(xg[g], wexp[g], out=yg[g])
        y = ops.beats_posconv_finish(x, yg, self.beats_conv_b, n, T, Dm, G)
        x = ops.layernorm(y, *self.beats_enc_ln, c.eps)
        M = n * T
        o = torch.empty((M, Dm), device=self.dev, dtype=torch.bfloat16)
        for L in self.beats_layers:
            qkv = ops.gemm(x, L["qkv"].w, bias=L["qkv"].b)
            gate = ops.beats_gate(qkv, L["grep_w"], L["grep_b"], L["grep_a"], n, T, Hh)
            ops.flash_attn(qkv, qkv[:, Dm:], qkv[:, 2 * Dm:], o, B=n, H=Hh, KVH=Hh, Sq=T, Sk=T, head_dim=hd,
                           q_strides=(T * 3 * Dm, 3 * Dm, hd), k_strides=(T * 3 * Dm, 3 * Dm, hd),
                           v_strides=(T * 3 * Dm, 3 * Dm, hd), o_strides=(T * Dm, Dm, hd), scale=hd ** -0.5, gate=gate,
                           bias_table=table)
            y = ops.gemm(o, L["out"].w, bias=L["out"].b, residual=x, res_scale=self.beats_alpha)
            x = ops.layernorm(y, *L["ln1"], c.eps)
            hmid = ops.gemm(x, L["fc1"].w, bias=L["fc1"].b, act=ops.ACT_GELU)
            y = ops.gemm(hmid, L["fc2"].w, bias=L["fc2"].b, residual=x, res_scale=self.beats_alpha)
            x = ops.layernorm(y, *L["ln2"], c.eps)
        return x, T

    def encode_audio(self, fbank: torch.Tensor) -> torch.Tensor:
        """fbank fp32 (n_segments_total, L, 128) -> bf16 [n_segments_total*32, d_model]
        (AudioEncoder -> ALProjector; models/unified_arch.py:152-155, models/multimodal_encoder.py:226-244)."""
        x, T = self.beats_forward(fbank)
        n = fbank.shape[0]
        f = ops.layernorm(x, *self.audio_ln, 1e-5)
        qo = self._qformer(self.aq, self.a_query, f, n, T, 0, T)
        h = ops.gemm(qo, self.a_proj[0].w, bias=self.a_proj[0].b, act=ops.ACT_GELU)
        return ops.gemm(h, self.a_proj[1].w, bias=self.a_proj[1].b)

    # ---- prepare_multimodal_inputs -----------------------------------------------------------------------------
    def prepare_inputs(self, batch_input_ids: Sequence[torch.Tensor], batch_X_modals: Sequence[dict],
                       want_image_taps: Sequence[int] = ()):
        """Splice modality embeddings at the placeholder ids, left-pad with pad-token embeddings
        (models/unified_arch.py:262-373).  The index bookkeeping runs on the host over the (host) token ids; the data
        movement is two gathers on the device.  Encoders run batched across samples (the reference runs them one
        sample at a time).  Returns (inputs_embeds bf16 [B,S,D], attention_mask int32 [B,S], position_ids [B,S])."""
        ids = self.cfg.special_ids
        key_of = {ids["<image>"]: "<image>", ids["<video>"]: "<video>", ids["<audio>"]: "<audio>"}
        nq = self.cfg.n_query
        D = self.cfg.decoder.hidden
        B = len(batch_input_ids)
        vis_items, aud_items = [], []  # tensors to encode, in order
        plans = []  # per sample: list of ("text", ids) | ("vis", item_idx, n_rows) | ("aud", item_idx, n_rows)
        for b in range(B):
            t = batch_input_ids[b].detach().cpu()
            tl = t.tolist()
            plan, pre = [], 0
            for i, tok in enumerate(tl):
                if tok in key_of:
                    plan.append(("text", t[pre:i]))
                    key = key_of[tok]
                    X = batch_X_modals[b][key]
                    if key == "<audio>":
                        X = X if X.dim() == 3 else X.unsqueeze(0)
                        plan.append(("aud", len(aud_items), X.shape[0] * nq))
                        aud_items.append(X)
                    else:
                        plan.append(("vis", len(vis_items), X.shape[0] * nq, key))
                        vis_items.append(X)
                    pre = i + 1
            plan.append(("text", t[pre:]))
            plans.append(plan)
        lens = [sum((len(p[1]) if p[0] == "text" else p[2]) for p in plan) for plan in plans]
        S = max(lens)
        # encoders, batched by input shape
        self._tap_layers, self._item_taps = tuple(want_image_taps), {}
        try:
            vis_out = self._encode_grouped(vis_items, self.encode_video)
        finally:
            self._tap_layers = ()
        aud_out = self._encode_grouped(aud_items, self.encode_audio)
        # per sample: the taps of its first '<image>' item (generate_avs reads one image per sample, unified_arch.py:243-247)
        self.image_taps = []
        if want_image_taps:
            for plan in plans:
                first = next((p[1] for p in plan if p[0] == "vis" and p[3] == "<image>"), None)
                self.image_taps.append(None if first is None else self._item_taps[first])
        embeds = torch.empty((B * S, D), device=self.dev, dtype=torch.bfloat16)
        txt_src, txt_dst = [], []
        mask = torch.zeros((B, S), dtype=torch.int32)
        for b, plan in enumerate(plans):
            padn = S - lens[b]
            txt_src.append(torch.full((padn,), self.cfg.pad_token_id, dtype=torch.long))
            txt_dst.append(torch.arange(b * S, b * S + padn))
            mask[b, padn:] = 1
            pos = b * S + padn
            for p in plan:
                if p[0] == "text":
                    txt_src.append(p[1].long())
                    txt_dst.append(torch.arange(pos, pos + len(p[1])))
                    pos += len(p[1])
                else:
                    src = (vis_out if p[0] == "vis" else aud_out)[p[1]]
                    dst_rows = torch.arange(pos, pos + p[2], device=self.dev)
                    ops.gather_rows(src, embeds, p[2], D, dst_rows=dst_rows)
                    pos += p[2]
        src = torch.cat(txt_src).to(self.dev, non_blocking=True)
        dst = torch.cat(txt_dst).to(self.dev, non_blocking=True)
        if src.numel():
            ops.gather_rows(self.embed, embeds, src.numel(), D, src_rows=src, dst_rows=dst)
        position_ids = torch.cumsum(mask, dim=-1) - 1
        position_ids[position_ids == -1] = 0
        return embeds.view(B, S, D), mask, position_ids

    def _encode_grouped(self, items: List[torch.Tensor], fn) -> List[torch.Tensor]:
        out: List[Optional[torch.Tensor]] = [None] * len(items)
        groups: Dict[Tuple, List[int]] = {}
        for i, x in enumerate(items):
            groups.setdefault((x.dtype == torch.uint8,) + tuple(x.shape[1:]), []).append(i)
        nq = self.cfg.n_query
        for shape, idxs in groups.items():
            u8 = all(items[i].dtype == torch.uint8 for i in idxs)  # raw frames stay uint8 (fused normalise + im2col)
            xs = torch.cat([items[i].to(self.dev, dtype=torch.uint8 if u8 else torch.float32, non_blocking=True)
                            for i in idxs], 0).contiguous()
            if self._small_call_graphs() and xs.shape[0] <= 32:
                name = "video" if fn == self.encode_video else "audio"
                static = self._buf("graph_in_" + name + "_" + "x".join(map(str, xs.shape)), tuple(xs.shape), xs.dtype)

                def run(xs_, fn=fn):   # the ViT taps are a side output of encode_video: they belong to the graph's outputs
                    y_ = fn(xs_)
                    return y_, dict(getattr(self, "_last_taps", {}))
                y, taps = self._graph_replay((name, tuple(xs.shape), xs.dtype), [static], [xs], run)
                y = y.clone()
                if name == "video":
                    self._last_taps = taps
            else:
                y = fn(xs)
            r = 0
            for i in idxs:
                nrows = items[i].shape[0] * nq
                out[i] = y[r:r + nrows]
                r += nrows
            if self._tap_layers and fn == self.encode_video:
                per = self._last_taps[self._tap_layers[0]].shape[0] // xs.shape[0]   # patch tokens per frame
                f0 = 0
                for i in idxs:
                    nf = items[i].shape[0]
                    self._item_taps[i] = [self._last_taps[k][f0 * per:(f0 + nf) * per] for k in self._tap_layers]
                    f0 += nf
        return out  # type: ignore

    # ---- decoder -----------------------------------------------------------------------------------------------
    def _alloc_cache(self, B: int):
        c = self.cfg.decoder
        if self.k_cache and self.k_cache[0].shape[0] == B:
            return
        self.k_cache = [torch.zeros((B, c.kv_heads, self.cfg.max_ctx, c.head_dim), device=self.dev, dtype=torch.bfloat16)
                        for _ in range(c.layers)]
        self.v_cache = [torch.zeros_like(k) for k in self.k_cache]
        self._graph = None
        self._call_graphs = {k: v for k, v in self._call_graphs.items() if k[0] != "prefill"}   # they hold the old cache's addresses

    def _decoder_layers(self, x: torch.Tensor, B: int, S: int, past: int, past_dev=None, len_dev=None, nsplit=1,
                        ws=None, tag="pf"):
        """x bf16 [B*S, D], updated in place through all layers: the dense-GEMM route.  S > 1: prefill (flash attention over
        the cache); S == 1 with past_dev/len_dev: a decode step for batches beyond the 32-row streaming chain."""
        c = self.cfg.decoder
        D, F, H, KV, hd = c.hidden, c.inter, c.heads, c.kv_heads, c.head_dim
        M = B * S
        nq, nk = H * hd, KV * hd
        xn = self._buf(tag + "_xn", (M, D + self.EXT_QKV), zero=True)
        qkv = self._buf(tag + "_qkv", (M, nq + 2 * nk))
        at = self._buf(tag + "_attn", (M, nq + self.EXT_O), zero=True)
        hh = self._buf(tag + "_h", (M, F + self.EXT_D), zero=True)
        ctx = self.cfg.max_ctx
        sc = self.scaling
        for li, L in enumerate(self.layers):
            ops.rmsnorm(x, L["ln1"], c.eps, out=xn[:, :D])
            if self.lora:
                ops.gemm(xn[:, :D], L["ra_qkv"], act=ops.ACT_LORA_Z, out_scale=sc, out=xn[:, D:D + 72])
            ops.gemm(xn, L["wqkv"], bias=L["bqkv"], out=qkv, k=D + (self.EXT_QKV if self.lora else 0))
            ops.rope_kv_append(qkv, self.rope, self.k_cache[li], self.v_cache[li], B, S, H, KV, hd, past=past, past_dev=past_dev)
            if S == 1 and len_dev is not None:
                ops.attn_decode(qkv, self.k_cache[li], self.v_cache[li], at[:, :nq], B=B, H=H, KVH=KV, head_dim=hd,
                                scale=1 / math.sqrt(hd), len_dev=len_dev, nsplit=nsplit, workspace=ws)
            else:
                ops.flash_attn(qkv, self.k_cache[li], self.v_cache[li], at, B=B, H=H, KVH=KV, Sq=S, Sk=past + S, head_dim=hd,
                               q_strides=(S * (nq + 2 * nk), nq + 2 * nk, hd), k_strides=(KV * ctx * hd, hd, ctx * hd),
                               v_strides=(KV * ctx * hd, hd, ctx * hd), o_strides=(S * (nq + self.EXT_O), nq + self.EXT_O, hd),
                               scale=1 / math.sqrt(hd), causal=True)
            if self.lora:
                ops.gemm(at[:, :nq], L["ra_o"], act=ops.ACT_LORA_Z, out_scale=sc, out=at[:, nq:nq + 24])
            ops.gemm(at, L["wo"], residual=x, out=x, k=nq + (self.EXT_O if self.lora else 0))
            ops.rmsnorm(x, L["ln2"], c.eps, out=xn[:, :D])
            if self.lora:
                ops.gemm(xn[:, :D], L["ra_gu"], act=ops.ACT_LORA_Z, out_scale=sc, out=xn[:, D:D + 48])
            ops.gemm(xn, L["wgu"], act=ops.ACT_SWIGLU, out=hh[:, :F], k=D + (self.EXT_GU if self.lora else 0))
            if self.lora:
                ops.gemm(hh[:, :F], L["ra_d"], act=ops.ACT_LORA_Z, out_scale=sc, out=hh[:, F:F + 24])
            ops.gemm(hh, L["wd"], residual=x, out=x, k=F + (self.EXT_D if self.lora else 0))
        return x

    def _head_phase(self, x_rows: torch.Tensor, logits: torch.Tensor):
        """final RMSNorm (rstd in the epilogue, gamma folded into lm_head_c) + lm_head as one chain phase."""
        c = self.cfg.decoder
        return ops.ChainPhase(x_rows, self.lm_head_c, logits, k=c.hidden, norm=True, eps=c.eps,
                              rstd=self._buf("dec_rstd_head", (32,), torch.float32), n=self.vocab)

    def _head(self, x_last: torch.Tensor, logits: torch.Tensor, next_ids: torch.Tensor, flag_kw: Optional[dict] = None):
        """final RMSNorm -> lm_head (fp32 logits) -> greedy arg-max."""
        c = self.cfg.decoder
        if x_last.shape[0] <= 32 and self.lm_head_p is not None:
            hn = ops.rmsnorm(x_last, self.final_norm, c.eps, out=self._buf("head_hn", tuple(x_last.shape)))
            ops.gemm_skinny(hn, self.lm_head_p, out=logits, n=self.vocab)
        elif x_last.shape[0] <= 32 and self.lm_head_c is not None:
            # one launch: the statistics cluster computes rstd, the lm_head tiles apply it in their epilogue (gamma is in lm_head_c)
            ops.gemm_skinny(x_last, self.lm_head_c, out=logits, n=self.vocab, norm=True, eps=c.eps,
                            rstd=self._buf("dec_rstd_head", (32,), torch.float32), tag="lm_head_skinny",
                            **(flag_kw if flag_kw is not None else dict(flags=self._flags("head"))))
        else:
            hn = ops.rmsnorm(x_last, self.final_norm, c.eps, out=self._buf("head_hn", tuple(x_last.shape)))
            ops.gemm(hn, self.lm_head, out=logits)
        ops.argmax(logits, self.vocab, out=next_ids)

    def _chain_counters(self) -> torch.Tensor:
        return self._buf("dec_chain_cnt", (288,), torch.int32, zero=True)   # the kernel leaves them zero

    def _flags(self, name: str) -> torch.Tensor:
        """Per-GEMM flag words of the fused skinny launches (zero on entry, left zero).  One buffer per linear: with programmatic
        dependent launch the next kernel's CTAs may poll before the previous launch has reset ITS flags."""
        return self._buf("dec_flags_" + name, (64,), torch.int32, zero=True)

    def _flag_ring(self, n: int):
        """Flag slots for the n statistics-carrying launches of one decode step: launch k raises slot k and zeroes slot k - 1 (the
        first one zeroes the last slot of the previous step), so no launch spends its tail resetting its own flags; plus the
        scratch through which several statistics clusters of one launch combine.  Returns a function handing out the kwargs."""
        if not self.flag_ring:
            names = iter(range(n))
            return lambda: dict(flags=self._flags("l%d" % (next(names) % 4)))
        ring = self._buf("dec_flag_ring_%d" % n, (n, 32), torch.int32, zero=True)
        scratch = self._buf("dec_stats_scratch", (8 * 36 * 32,), torch.float32)
        it = iter(range(n))

        def slot():
            k = next(it)
            return dict(flags=ring[k], flags_clear=ring[(k - 1) % n], stats_scratch=scratch, stats_clusters=self.stats_clusters)
        return slot

    def prefill(self, inputs_embeds: torch.Tensor):
        """inputs_embeds bf16 [B,S,D] (consumed in place) -> (last-position logits fp32 [B, vocab], next ids [B])."""
        B, S, D = inputs_embeds.shape
        assert S < self.cfg.max_ctx
        self._alloc_cache(B)
        if inputs_embeds.dtype != torch.bfloat16 or not inputs_embeds.is_cuda or not inputs_embeds.is_contiguous():
            inputs_embeds = inputs_embeds.to(device=self.dev, dtype=torch.bfloat16).contiguous()
        x = inputs_embeds.view(B * S, D)
        t = getattr(self, "_tail_rows", 0)
        if self._small_call_graphs() and B * S <= 2048 and not t:
            self.logits = self._buf("logits", (B, self.vocab_pad), torch.float32)
            self.next_ids = self._buf("next_ids", (B,), torch.int64)

            def body(xs):
                self._decoder_layers(xs, B, S, past=0)
                last_ = self._buf("last_x", (B, D))
                ops.gather_rows(xs, last_, B, D, src_rows=(torch.arange(B, device=self.dev) * S + (S - 1)))
                self._head(last_, self.logits, self.next_ids)
            self._graph_replay(("prefill", B, S, id(self.k_cache[0])), [self._buf("graph_in_prefill_%dx%d" % (B, S), (B * S, D))], [x], body)
            self.cur_len = S
            return self.logits[:, : self.vocab], self.next_ids
        self._decoder_layers(x, B, S, past=0)
        self.cur_len = S
        last = self._buf("last_x", (B, D))
        rows = (torch.arange(B, device=self.dev) * S + (S - 1))
        ops.gather_rows(x, last, B, D, src_rows=rows)
        if t:  # final-normed hidden states of the last t prompt positions (HF hidden_states[0][-1][:, -t:])
            trows = (torch.arange(B, device=self.dev).view(B, 1) * S + torch.arange(S - t, S, device=self.dev).view(1, t)).reshape(-1)
            tail = torch.empty((B * t, D), device=self.dev, dtype=torch.bfloat16)
            ops.gather_rows(x, tail, B * t, D, src_rows=trows.contiguous())
            self.hidden_prefill_tail = ops.rmsnorm(tail, self.final_norm, self.cfg.decoder.eps).view(B, t, D)
        self.logits = self._buf("logits", (B, self.vocab_pad), torch.float32)
        self.next_ids = self._buf("next_ids", (B,), torch.int64)
        self._head(last, self.logits, self.next_ids)
        return self.logits[:, : self.vocab], self.next_ids

    # ---- decode ------------------------------------------------------------------------------------------------
    def _chain_plan(self, B: int, nsplit: int):
        """The decode step as a list of launches: ('chain', [phases]) / ('attn', layer).  Per layer: ONE attention launch
        (RoPE + KV append + attention [+ o_proj LoRA pre-pass]) and ONE persistent GEMM-chain launch
        o_proj -> gate/up+SwiGLU -> down_proj -> next layer's qkv (or final norm + lm_head): 2 launches per layer instead of 8."""
        key = (B, nsplit)
        if getattr(self, "_plan_key", None) == key:
            return self._plan
        c = self.cfg.decoder
        D, F, H, KV, hd = c.hidden, c.inter, c.heads, c.kv_heads, c.head_dim
        nq, nk = H * hd, KV * hd
        lo, sc = self.lora, self.scaling
        x = self._buf("dec_x", (B, D))
        qkv = self._buf("dec_qkv", (B, nq + 2 * nk))
        at = self._buf("dec_attn", (B, nq + self.EXT_O), zero=True)
        hh = self._buf("dec_h", (B, F))
        z = {n: self._buf("dec_z_" + n, (32, 128), zero=True) for n in ("qkv", "o", "gu", "d")}
        rs = {n: self._buf("dec_rstd_" + n, (32,), torch.float32) for n in ("qkv", "gu")}
        G = H // KV
        gqa_tc = G > 1 and hd == 128 and B * KV >= 64 and self.gqa_decode_tc
        fused = self.fuse_decode_attn and not gqa_tc
        o_fused_lora = lo and (fused or (gqa_tc and nsplit > 1))     # the attention (or its split-KV combine) kernel writes o_proj's z columns into at[:, nq:]
        self._dec_mode = (fused, gqa_tc, o_fused_lora)

        def qkv_phase(L):
            return ops.ChainPhase(x, L["wqkv_c"], qkv, k=D, z=z["qkv"] if lo else None, kext=self.EXT_QKV if lo else 0,
                                  stats=L.get("st_qkv"), stats_linears=3 if lo else 0, norm=True, eps=c.eps, lora_scale=sc,
                                  rstd=rs["qkv"], bias=L["bqkv"])

        plan = [("chain", [qkv_phase(self.layers[0])])]
        for li, L in enumerate(self.layers):
            plan.append(("attn", li))
            if o_fused_lora:
                o_ph = ops.ChainPhase(at, L["wo_c"], x, k=nq, z=at[:, nq:], kext=self.EXT_O, residual=x)
            else:
                o_ph = ops.ChainPhase(at, L["wo_c"], x, k=nq, z=z["o"] if lo else None, kext=self.EXT_O if lo else 0,
                                      stats=L.get("st_o"), stats_linears=1 if lo else 0, lora_scale=sc, residual=x)
            gu_ph = ops.ChainPhase(x, L["wgu_c"], hh, k=D, z=z["gu"] if lo else None, kext=self.EXT_GU if lo else 0,
                                   stats=L.get("st_gu"), stats_linears=2 if lo else 0, norm=True, eps=c.eps, lora_scale=sc,
                                   rstd=rs["gu"], act=ops.ACT_SWIGLU)
            d_ph = ops.ChainPhase(hh, L["wd_c"], x, k=F, z=z["d"] if lo else None, kext=self.EXT_D if lo else 0,
                                  stats=L.get("st_d"), stats_linears=1 if lo else 0, lora_scale=sc, residual=x)
            last = li + 1 == len(self.layers)
            tail = self._head_phase(x, self.logits) if last else qkv_phase(self.layers[li + 1])
            plan.append(("chain", [o_ph, gu_ph, d_ph, tail]))
        self._plan_key, self._plan = key, plan
        return plan

    def _decode_attention(self, li: int, B: int, nsplit: int, ws, qkv, at, fused: bool, gqa_tc: bool, o_fused_lora: bool):
        """The decode step's attention launch(es) of layer li: RoPE on q / new k, KV append, attention over past + 1 keys."""
        c = self.cfg.decoder
        H, KV, hd = c.heads, c.kv_heads, c.head_dim
        nq, nk = H * hd, KV * hd
        ctx = self.cfg.max_ctx
        L = self.layers[li]
        if gqa_tc:
            # grouped-query decode: the G query heads of a kv group are the Sq = G "rows" of one flash-attention problem, so QK^T / PV run
            # on tensor cores; B x KVH problems alone keep only 32 KB each in flight, so the keys are split over nsplit blocks and the
            # combine launch finishes (and does the o_proj LoRA pre-pass when o_fused_lora)
            ops.attn_decode_fused(qkv, self.rope, self.k_cache[li], self.v_cache[li], at[:, :nq], B=B, H=H, KVH=KV, head_dim=hd,
                                  scale=1 / math.sqrt(hd), past_dev=self.past_dev, nsplit=nsplit, workspace=ws, gqa_tc=True,
                                  ra=L["ra_o"] if o_fused_lora else None, z=at[:, nq:] if o_fused_lora else None, lora_scale=self.scaling,
                                  lora_ws=self._buf("dec_lora_ws", (B * H * 11,), torch.float32) if o_fused_lora else None,
                                  lora_counters=self._buf("dec_lora_cnt", (B,), torch.int32, zero=True) if o_fused_lora else None)
        elif fused:
            # one launch: RoPE on q / new k, cache append, attention over past + 1 keys, and (nsplit == 1) the o_proj
            # LoRA pre-pass whose z columns land in at[:, nq:]
            ops.attn_decode_fused(qkv, self.rope, self.k_cache[li], self.v_cache[li], at[:, :nq], B=B, H=H, KVH=KV,
                                  head_dim=hd, scale=1 / math.sqrt(hd), past_dev=self.past_dev, nsplit=nsplit, workspace=ws,
                                  ra=L["ra_o"] if o_fused_lora else None, z=at[:, nq:] if o_fused_lora else None, lora_scale=self.scaling,
                                  lora_ws=self._buf("dec_lora_ws", (B * H * 11,), torch.float32) if o_fused_lora else None,
                                  lora_counters=self._buf("dec_lora_cnt", (B,), torch.int32, zero=True) if o_fused_lora else None)
        else:
            ops.rope_kv_append(qkv, self.rope, self.k_cache[li], self.v_cache[li], B, 1, H, KV, hd, past=0, past_dev=self.past_dev)
            ops.attn_decode(qkv, self.k_cache[li], self.v_cache[li], at[:, :nq], B=B, H=H, KVH=KV, head_dim=hd,
                            scale=1 / math.sqrt(hd), len_dev=self.len_dev, nsplit=nsplit, workspace=ws)

    def _decode_body_skinny(self, B: int, nsplit: int, ws):
        """Decode step, 5 launches per layer: qkv | attention | o | gate/up | down, each linear ONE weight-streaming launch that
        carries its own RMSNorm (gamma folded into the packed weight, rstd applied in the epilogue) and hyper-LoRA pre-pass (a
        statistics cluster of the same launch) — the three row kernels per layer of round 1 are gone."""
        c = self.cfg.decoder
        D, F, H, KV, hd = c.hidden, c.inter, c.heads, c.kv_heads, c.head_dim
        nq, nk = H * hd, KV * hd
        lo, sc = self.lora, self.scaling
        x = self._buf("dec_x", (B, D))
        qkv = self._buf("dec_qkv", (B, nq + 2 * nk))
        at = self._buf("dec_attn", (B, nq + self.EXT_O), zero=True)
        hh = self._buf("dec_h", (B, F))
        z = {n: self._buf("dec_z_" + n, (32, 128), zero=True) for n in ("qkv", "o", "gu", "d")}
        rs = {n: self._buf("dec_rstd_" + n, (32,), torch.float32) for n in ("qkv", "gu")}
        G = H // KV
        gqa_tc = G > 1 and hd == 128 and B * KV >= 64 and self.gqa_decode_tc
        fused = self.fuse_decode_attn and not gqa_tc
        o_fused_lora = lo and (fused or (gqa_tc and nsplit > 1))     # the attention (or its split-KV combine) kernel writes o_proj's z columns into at[:, nq:]
        o_stats = lo and not o_fused_lora
        head_fused = B <= 32 and self.lm_head_p is None and self.lm_head_c is not None
        fslot = self._flag_ring(len(self.layers) * (2 + int(o_stats) + int(lo)) + int(head_fused))
        ops.set_pdl(self.pdl_chain)
        try:
            ops.gather_rows(self.embed, x, B, D, src_rows=self.next_ids)  # embed_tokens of the previous arg-max
            pfb = self.prefetch_mb << 20

            def pf(w):
                return dict(prefetch=w.data, prefetch_bytes=pfb) if pfb else {}
            for li, L in enumerate(self.layers):
                nxt_qkv = self.layers[li + 1]["wqkv_c"] if li + 1 < len(self.layers) else self.lm_head_c
                ops.gemm_skinny(x, L["wqkv_c"], bias=L["bqkv"], out=qkv, z=z["qkv"] if lo else None, kext=self.EXT_QKV if lo else 0,
                                stats=L.get("st_qkv"), stats_linears=3 if lo else 0, norm=True, eps=c.eps, lora_scale=sc, rstd=rs["qkv"],
                                splits=self.skinny_splits["qkv"], **fslot())
                self._decode_attention(li, B, nsplit, ws, qkv, at, fused, gqa_tc, o_fused_lora)
                # the kernel right after the decode attention is launched under its own PDL mask: early-resident streaming-GEMM
                # CTAs must not squat on the SMs while the 1024-CTA attention kernel still runs
                if self.pdl_after_attn != self.pdl_chain:
                    ops.set_pdl(self.pdl_after_attn)
                if o_fused_lora:
                    ops.gemm_skinny(at, L["wo_c"], residual=x, out=x, z=at[:, nq:], kext=self.EXT_O, splits=self.skinny_splits["o"], **pf(L["wgu_c"]))
                else:
                    ops.gemm_skinny(at, L["wo_c"], residual=x, out=x, z=z["o"] if lo else None, kext=self.EXT_O if lo else 0,
                                    stats=L.get("st_o"), stats_linears=1 if lo else 0, lora_scale=sc,
                                    splits=self.skinny_splits["o"], **(fslot() if lo else {}), **pf(L["wgu_c"]))
                if self.pdl_after_attn != self.pdl_chain:
                    ops.set_pdl(self.pdl_chain)
                ops.gemm_skinny(x, L["wgu_c"], act=ops.ACT_SWIGLU, out=hh, z=z["gu"] if lo else None, kext=self.EXT_GU if lo else 0,
                                stats=L.get("st_gu"), stats_linears=2 if lo else 0, norm=True, eps=c.eps, lora_scale=sc, rstd=rs["gu"],
                                splits=self.skinny_splits["gu"], **fslot(), **pf(L["wd_c"]))
                ops.gemm_skinny(hh, L["wd_c"], residual=x, out=x, z=z["d"] if lo else None, kext=self.EXT_D if lo else 0,
                                stats=L.get("st_d"), stats_linears=1 if lo else 0, lora_scale=sc,
                                splits=self.skinny_splits["d"], **(fslot() if lo else {}), **pf(nxt_qkv))
            self._head(x, self.logits, self.next_ids, flag_kw=fslot() if head_fused else None)
            ops.add_scalar_i32(self.past_dev, 1)
            ops.add_scalar_i32(self.len_dev, 1)
        finally:
            ops.set_pdl(0)  # prefill / encoder launches are never PDL launches

    def _decode_body_rows(self, B: int, nsplit: int, ws):
        """Decode step, round-1 organisation (8 launches per layer): a cluster row kernel (RMSNorm + hyper-LoRA router / A pre-pass
        -> normalised row + z columns) in front of each weight-streaming GEMM, whose K-extension columns sit behind the row."""
        c = self.cfg.decoder
        D, F, H, KV, hd = c.hidden, c.inter, c.heads, c.kv_heads, c.head_dim
        nq, nk = H * hd, KV * hd
        lo, sc = self.lora, self.scaling
        x = self._buf("dec_x", (B, D))
        xn = self._buf("dec_xn", (B, D + self.EXT_QKV), zero=True)
        qkv = self._buf("dec_qkv", (B, nq + 2 * nk))
        at = self._buf("dec_attn", (B, nq + self.EXT_O), zero=True)
        hh = self._buf("dec_hx", (B, F + self.EXT_D), zero=True)
        G = H // KV
        gqa_tc = G > 1 and hd == 128 and B * KV >= 64 and self.gqa_decode_tc
        fused = self.fuse_decode_attn and not gqa_tc
        o_fused_lora = lo and (fused or (gqa_tc and nsplit > 1))
        ops.set_pdl(self.pdl_chain)
        try:
            ops.gather_rows(self.embed, x, B, D, src_rows=self.next_ids)
            for li, L in enumerate(self.layers):
                ops.row_norm_loraz(x, gamma=L["ln1"], eps=c.eps, y=xn[:, :D], ra=L.get("ra_qkv"), groups=3 if lo else 0,
                                   z=xn[:, D:] if lo else None, scale=sc)
                ops.gemm_skinny(xn, L["wqkv_p"], bias=L["bqkv"], out=qkv)
                self._decode_attention(li, B, nsplit, ws, qkv, at, fused, gqa_tc, o_fused_lora)
                if self.pdl_after_attn != self.pdl_chain:
                    ops.set_pdl(self.pdl_after_attn)
                if lo and not o_fused_lora:
                    ops.row_norm_loraz(at[:, :nq], ra=L["ra_o"], groups=1, z=at[:, nq:], scale=sc)
                    if self.pdl_after_attn != self.pdl_chain:
                        ops.set_pdl(self.pdl_chain)
                ops.gemm_skinny(at, L["wo_p"], residual=x, out=x)
                if self.pdl_after_attn != self.pdl_chain:
                    ops.set_pdl(self.pdl_chain)
                ops.row_norm_loraz(x, gamma=L["ln2"], eps=c.eps, y=xn[:, :D], ra=L.get("ra_gu"), groups=2 if lo else 0,
                                   z=xn[:, D:] if lo else None, scale=sc)
                ops.gemm_skinny(xn, L["wgu_p"], act=ops.ACT_SWIGLU, out=hh[:, :F])
                if lo:
                    ops.row_norm_loraz(hh[:, :F], ra=L["ra_d"], groups=1, z=hh[:, F:], scale=sc)
                ops.gemm_skinny(hh, L["wd_p"], residual=x, out=x)
            self._head(x, self.logits, self.next_ids)
            ops.add_scalar_i32(self.past_dev, 1)
            ops.add_scalar_i32(self.len_dev, 1)
        finally:
            ops.set_pdl(0)

    def _decode_body(self, B: int, nsplit: int, ws):
        c = self.cfg.decoder
        D, H, KV, hd = c.hidden, c.heads, c.kv_heads, c.head_dim
        nq, nk = H * hd, KV * hd
        ctx = self.cfg.max_ctx
        if not (B <= 32 and self.decode_packed):
            # more than 32 rows: the dense-GEMM route
            x = self._buf("dec_x", (B, D))
            ops.gather_rows(self.embed, x, B, D, src_rows=self.next_ids)
            self._decoder_layers(x, B, 1, past=0, past_dev=self.past_dev, len_dev=self.len_dev, nsplit=nsplit, ws=ws, tag="dec")
            self._head(x, self.logits, self.next_ids)
            ops.add_scalar_i32(self.past_dev, 1)
            ops.add_scalar_i32(self.len_dev, 1)
            return
        if self.decode_mode == "rows":
            return self._decode_body_rows(B, nsplit, ws)
        if self.decode_mode != "chain":
            return self._decode_body_skinny(B, nsplit, ws)
        plan = self._chain_plan(B, nsplit)
        fused, gqa_tc, o_fused_lora = self._dec_mode
        x = self._buf("dec_x", (B, D))
        qkv = self._buf("dec_qkv", (B, nq + 2 * nk))
        at = self._buf("dec_attn", (B, nq + self.EXT_O), zero=True)
        cnt = self._chain_counters()
        sc = self.scaling
        ops.set_pdl(0)   # the persistent chain owns every SM: its neighbours are plain (fully ordered) launches
        try:
            ops.gather_rows(self.embed, x, B, D, src_rows=self.next_ids)  # embed_tokens of the previous arg-max
            for kind, arg in plan:
                if kind == "chain":
                    ops.decode_chain(arg, B, cnt, self.chain_cluster)
                    continue
                self._decode_attention(arg, B, nsplit, ws, qkv, at, fused, gqa_tc, o_fused_lora)
            ops.argmax(self.logits, self.vocab, out=self.next_ids)
            ops.add_scalar_i32(self.past_dev, 1)
            ops.add_scalar_i32(self.len_dev, 1)
        finally:
            ops.set_pdl(0)  # prefill / encoder launches are never PDL launches

    def begin_decode(self, B: int, use_graph: bool = True, max_len: Optional[int] = None):
        """Prepare the decode loop after a prefill.  With `use_graph`, one decode step (all layers + head + arg-max +
        position bump) is captured in a CUDA graph ONCE per (batch, buffers) and replayed every step of every later
        request: the context length lives in device memory (`past_dev`, `len_dev`), so nothing in the graph changes."""
        c = self.cfg.decoder
        if getattr(self, "past_dev", None) is None:
            self.past_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.len_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.past_dev.fill_(self.cur_len)
        self.len_dev.fill_(self.cur_len + 1)
        blocks = B * c.kv_heads
        nsplit = 1 if blocks >= 2 * 148 else max(1, min(16, (2 * 148 + blocks - 1) // blocks))
        G = c.heads // c.kv_heads
        if G > 1 and c.head_dim == 128 and blocks >= 64 and self.gqa_decode_tc:
            # tensor-core GQA decode: two 80 KB blocks per SM — keep all splits in one wave (3 x 128 blocks = a second wave: 30 us vs 18)
            nsplit = max(1, min(8, (2 * 148) // blocks))
            nsplit = int(os.environ.get("CRAB_GQA_NSPLIT", nsplit))
        if max_len is not None and max_len <= 256:
            # a short context is not worth the combine launch (~8 us per layer): one block per (b, kv head).  Longer ones keep the
            # split that fills the SMs — measured at bs 1, S = 638: 3 splits 3.38 ms / step, 10 splits 3.22 ms
            nsplit = 1
        ws = self._buf("dec_ws", (B * c.heads * nsplit * (c.head_dim + 2),), torch.float32) if nsplit > 1 else None
        self._dec_args = (B, nsplit, ws)
        self._use_graph = use_graph
        if use_graph and (self._graph is None or self._graph_bs != (B, nsplit)):
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream())
            saved = (self.next_ids.clone(), self.past_dev.clone(), self.len_dev.clone())
            with torch.cuda.stream(s):
                self._decode_body(*self._dec_args)  # warm-up outside capture (buffers, kernel attributes)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.next_ids.copy_(saved[0]); self.past_dev.copy_(saved[1]); self.len_dev.copy_(saved[2])
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                self._decode_body(*self._dec_args)
            self._graph_kernels = ops.launch_count() - n0
            ops.count_launches(-self._graph_kernels)  # capture launches nothing
            self.next_ids.copy_(saved[0]); self.past_dev.copy_(saved[1]); self.len_dev.copy_(saved[2])
            self._graph, self._graph_bs = g, (B, nsplit)

    begin_decode_cached = begin_decode

    def decode_step(self):
        """One greedy step: consumes self.next_ids, leaves the new arg-max there and fp32 logits in self.logits."""
        if self.cur_len >= self.cfg.max_ctx:
            # the step would write K/V row `cur_len` and read RoPE row `cur_len`: both are outside the allocations
            raise ops._l.CrabError(f"decode_step at position {self.cur_len}: the KV cache holds max_ctx = {self.cfg.max_ctx} positions")
        if self._use_graph:
            self._graph.replay()
            ops.count_launches(self._graph_kernels)
        else:
            self._decode_body(*self._dec_args)
        self.cur_len += 1
        return self.logits[:, : self.vocab], self.next_ids

    @torch.no_grad()
    def generate_from_embeds(self, inputs_embeds: torch.Tensor, max_new_tokens: int, use_graph: bool = True,
                             return_logits: bool = False, teacher_tokens: Optional[torch.Tensor] = None,
                             capture_hidden: int = 0, eos_token_id=None, pad_token_id: Optional[int] = None,
                             sampling: Optional[dict] = None):
        """Greedy loop.  Returns ids [B, n] (int64, device), n <= max_new_tokens.
        eos_token_id (int / list / None): HF `generate` semantics — a row that emitted EOS produces `pad_token_id` afterwards and
        decoding stops right after the step at which the last row finished (the flag is read back once every 8 tokens, so up to 7
        surplus steps are decoded and dropped).  None = fixed length.
        capture_hidden = t > 0 additionally records what HF's `output_hidden_states` exposes of the LAST layer (after the final
        norm): `self.hidden_prefill_tail` [B, min(t, S), D] (the last positions of the prompt pass) and `self.hidden_steps`
        [n - 1, B, D] (one row per decode step) — the inputs of generate_avs' mask-token pairing (unified_llama.py:335-345)."""
        B, S = inputs_embeds.shape[0], inputs_embeds.shape[1]
        if S + max_new_tokens > self.cfg.max_ctx:
            raise ops._l.CrabError(f"prompt of {S} positions + {max_new_tokens} new tokens exceeds the KV cache (max_ctx = "
                                   f"{self.cfg.max_ctx}): construct the model / CrabConfig with a larger max_ctx")
        self._tail_rows = min(int(capture_hidden), S) if capture_hidden else 0
        logits, nxt = self.prefill(inputs_embeds)
        self._tail_rows = 0
        if sampling is not None:
            nxt = self._sample(sampling)   # replaces the arg-max in self.next_ids, which the next decode step embeds
        pad = int(self.cfg.pad_token_id if pad_token_id is None else pad_token_id)
        eos = None
        if eos_token_id is not None:
            eos = torch.as_tensor(list(eos_token_id) if isinstance(eos_token_id, (list, tuple)) else [int(eos_token_id)],
                                  device=self.dev, dtype=torch.int64)
            done = torch.zeros(B, dtype=torch.bool, device=self.dev)
            alive = []  # per step: does any row still need tokens after it?  (device flags; read back in batches)
        out = torch.full((B, max_new_tokens), pad, device=self.dev, dtype=torch.int64)
        all_logits = [logits.clone()] if return_logits else None
        steps_h = []
        if max_new_tokens > 1:
            self.begin_decode(B, use_graph and self.dev.type == "cuda", max_len=S + max_new_tokens)
        steps = max_new_tokens
        for step in range(max_new_tokens):
            if eos is None:
                out[:, step].copy_(nxt)
            else:
                tok = torch.where(done, torch.full_like(nxt, pad), nxt)
                out[:, step] = tok
                done |= torch.isin(tok, eos)
                alive.append((~done).any())
                if step % 8 == 7 and bool(done.all()):  # one host sync every 8 tokens instead of every token
                    break
            if step + 1 == max_new_tokens:
                break
            if teacher_tokens is not None:
                self.next_ids.copy_(teacher_tokens[:, step])
            logits, nxt = self.decode_step()
            if sampling is not None:
                nxt = self._sample(sampling)
            if return_logits:
                all_logits.append(logits.clone())
            if capture_hidden:
                # HF's last hidden state of the step: the final norm of the residual stream the decode step left in dec_x
                # (the chain applies that norm inside the lm_head epilogue and never materialises it)
                steps_h.append(ops.rmsnorm(self._buf("dec_x", (B, self.cfg.decoder.hidden)), self.final_norm, self.cfg.decoder.eps))
        if eos is not None:
            hist = torch.stack(alive).cpu().tolist()
            steps = (hist.index(False) + 1) if False in hist else len(hist)
            out = out[:, :steps]
        if capture_hidden:
            D = self.cfg.decoder.hidden
            steps_h = steps_h[: max(steps - 1, 0)]
            self.hidden_steps = torch.stack(steps_h, 0) if steps_h else torch.empty((0, B, D), device=self.dev, dtype=torch.bfloat16)
        if return_logits:
            return out, torch.stack(all_logits[: out.shape[1]], 0)
        return out

    def _sample(self, sampling: dict) -> torch.Tensor:
        """Temperature / top-k / top-p draw from self.logits into self.next_ids (HF generate(do_sample=True) semantics; the uniforms
        come from torch's generator on this device, so runs are reproducible with a seeded `generator`)."""
        B = self.logits.shape[0]
        u = torch.rand(B, device=self.dev, dtype=torch.float32, generator=sampling.get("generator"))
        ops.sample_top_k_top_p(self.logits, self.vocab, u, temperature=float(sampling.get("temperature", 1.0)),
                               top_k=int(sampling.get("top_k", 0)), top_p=float(sampling.get("top_p", 1.0)), out=self.next_ids)
        return self.next_ids

    @torch.no_grad()
    def forward_logits(self, inputs_embeds: torch.Tensor, labels: Optional[torch.Tensor] = None):
        """UnifiedForCausalLM.forward over a whole sequence (models/unified_llama.py:129-160 -> HF LlamaForCausalLM.forward): fp32
        logits for ALL positions [B, S, vocab] and, with `labels` [B, S] (ignore_index -100), the mean shifted cross-entropy.
        The KV cache is left filled as after `prefill` (a decode step may follow)."""
        B, S, D = inputs_embeds.shape
        assert S < self.cfg.max_ctx
        self._alloc_cache(B)
        x = inputs_embeds.to(device=self.dev, dtype=torch.bfloat16).contiguous().clone().view(B * S, D)
        self._decoder_layers(x, B, S, past=0)
        self.cur_len = S
        hn = ops.rmsnorm(x, self.final_norm, self.cfg.decoder.eps)
        logits = torch.empty((B * S, self.vocab_pad), device=self.dev, dtype=torch.float32)
        ops.gemm(hn, self.lm_head, out=logits)
        loss = None
        if labels is not None:
            lab = torch.full((B, S), -100, device=self.dev, dtype=torch.int64)
            lab[:, :-1] = labels.to(self.dev)[:, 1:]          # position t predicts token t + 1
            per_row = ops.cross_entropy(logits, self.vocab, lab.view(-1).contiguous())
            loss = per_row.sum() / (lab >= 0).sum().clamp(min=1)
        # keep the greedy state consistent with prefill(): last-position logits / arg-max
        self.logits = self._buf("logits", (B, self.vocab_pad), torch.float32)
        self.next_ids = self._buf("next_ids", (B,), torch.int64)
        self.logits.copy_(logits.view(B, S, -1)[:, -1])
        ops.argmax(self.logits, self.vocab, out=self.next_ids)
        return logits.view(B, S, self.vocab_pad)[:, :, : self.vocab], loss

    @torch.no_grad()
    def generate(self, batch_input_ids, batch_X_modals, max_new_tokens: int, use_graph: bool = True):
        embeds, _, _ = self.prepare_inputs(batch_input_ids, batch_X_modals)
        return self.generate_from_embeds(embeds, max_new_tokens, use_graph)
