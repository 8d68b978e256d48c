"""Mirror of the reference's `models/unified_llama.py` (UnifiedConfig / UnifiedModel / UnifiedForCausalLM over
LLaMA), backed by `crab_b200.engine.CrabEngine`.

What is kept: class names, `from_pretrained(path, config=, torch_dtype=)`, `get_model()`, `forward(...)`,
`generate(batch_input_ids, batch_labels, batch_X_modals, batch_task_names, **kw)`, `prepare_inputs_for_generation`,
`.npu()`, `.device`, the state-dict key names (`model.layers.N.self_attn.q_proj.weight` …) and real `nn.Linear` leaves
named q/k/v/o/gate/up/down_proj so `peft_hyper.get_peft_model` (scripts/quick_start.py:475-493) can wrap them with
its hyper-LoRA `Linear` (peft_hyper/tuners/lora.py:118-159).  What is replaced: all arithmetic.  The modules hold
parameters only; on first use (or after the weights change) they are packed into the engine's bf16 device buffers.

Reference: models/unified_llama.py:11-40 (classes), :47-161 (forward), :244-267 (generate), :365-382
(prepare_inputs_for_generation), :385-387 (device).
"""
from __future__ import annotations

import glob
import json
import os
from types import SimpleNamespace
from typing import List, Optional

import torch
from torch import nn
from transformers import LlamaConfig

from ..engine import CrabEngine, DecoderConfig
from .unified_arch import UnifiedMetaForCausalLM, UnifiedMetaModel, build_crab_config


def select_pred_embeddings(output_ids, mask_token_ids, prefill_tail: torch.Tensor, step_hidden: torch.Tensor):
    """The reference's pairing of generated tokens and hidden states (models/unified_llama.py:331-351): entry t of
    `output_hidden_states` (t = 0: the prompt pass, (S, D); t >= 1: one row) is kept when generated token t + 1 is a `<mask_i>`
    token; the kept rows are concatenated, more than six -> the last six, fewer than six -> no segmentation (returns None).
    `prefill_tail` holds the last <= 6 rows of the prompt pass (all that can survive the last-six rule), `step_hidden[t - 1]` the
    row of step t.  Returns a (6, D) tensor or None."""
    n = len(output_ids)
    rows, count = [], 0
    for t in range(n - 1):                       # zip(mask_list over output_ids[1:], hidden_states) stops at n - 1 entries
        if output_ids[t + 1] in mask_token_ids:
            if t == 0:
                rows.append(prefill_tail)        # stands for all S rows of the prompt pass: only its tail can reach the last six
                count += 10 ** 6                 # "at least six more rows than needed"
            else:
                rows.append(step_hidden[t - 1].unsqueeze(0))
                count += 1
    if count < 6:
        return None
    allrows = torch.cat(rows, 0)
    if allrows.shape[0] < 6:
        raise RuntimeError("prompt shorter than six positions: cannot apply the reference's last-six rule")
    return allrows[-6:]


class UnifiedConfig(LlamaConfig):
    model_type = "unified_llm"


class _NormWeight(nn.Module):
    def __init__(self, dim: int, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim, dtype=dtype), requires_grad=False)


class _Attn(nn.Module):
    def __init__(self, c, dtype, bias: bool):
        super().__init__()
        hd = getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads
        kv = getattr(c, "num_key_value_heads", None) or c.num_attention_heads
        self.q_proj = nn.Linear(c.hidden_size, c.num_attention_heads * hd, bias=bias, dtype=dtype)
        self.k_proj = nn.Linear(c.hidden_size, kv * hd, bias=bias, dtype=dtype)
        self.v_proj = nn.Linear(c.hidden_size, kv * hd, bias=bias, dtype=dtype)
        self.o_proj = nn.Linear(c.num_attention_heads * hd, c.hidden_size, bias=False, dtype=dtype)


class _Mlp(nn.Module):
    def __init__(self, c, dtype):
        super().__init__()
        self.gate_proj = nn.Linear(c.hidden_size, c.intermediate_size, bias=False, dtype=dtype)
        self.up_proj = nn.Linear(c.hidden_size, c.intermediate_size, bias=False, dtype=dtype)
        self.down_proj = nn.Linear(c.intermediate_size, c.hidden_size, bias=False, dtype=dtype)


class _Layer(nn.Module):
    def __init__(self, c, dtype, qkv_bias):
        super().__init__()
        self.self_attn = _Attn(c, dtype, qkv_bias)
        self.mlp = _Mlp(c, dtype)
        self.input_layernorm = _NormWeight(c.hidden_size, dtype)
        self.post_attention_layernorm = _NormWeight(c.hidden_size, dtype)


class UnifiedModel(UnifiedMetaModel, nn.Module):
    """Decoder parameter container (`model.*` keys of the reference checkpoints)."""

    config_class = UnifiedConfig
    qkv_bias = False

    def __init__(self, config, dtype=torch.float32):
        nn.Module.__init__(self)
        self.config = config
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, dtype=dtype)
        self.layers = nn.ModuleList([_Layer(config, dtype, self.qkv_bias) for _ in range(config.num_hidden_layers)])
        self.norm = _NormWeight(config.hidden_size, dtype)
        self.pad_token_id = getattr(config, "pad_token_id", None) or 0
        self._owner = None

    def _engine_stale(self):
        if self._owner is not None:
            self._owner()._engine = None


class UnifiedForCausalLM(UnifiedMetaForCausalLM, nn.Module):
    config_class = UnifiedConfig
    _model_cls = UnifiedModel

    def __init__(self, config, torch_dtype=None, max_ctx: int = 2048, **kwargs):
        nn.Module.__init__(self)
        dtype = torch_dtype or torch.float32
        with torch.device("cpu"):
            self.model = self._model_cls(config, dtype)
            self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False, dtype=dtype)
        import weakref

        self.model._owner = weakref.ref(self)
        self.config = config
        self.vocab_size = config.vocab_size
        self.pretraining_tp = getattr(config, "pretraining_tp", 1)
        self.max_ctx = max_ctx
        self.is_avs_task = False
        self._engine: Optional[CrabEngine] = None
        self._engine_fp = None
        self._target = torch.device("cpu")
        # what HF `generate` would read from the checkpoint: generation_config.json if present (from_pretrained), else the
        # model config's eos / pad ids.  quick_start.py calls generate(**sample, use_cache=True, max_new_tokens=N) and relies on it.
        self.generation_config = SimpleNamespace(max_new_tokens=20, eos_token_id=getattr(config, "eos_token_id", None),
                                                 pad_token_id=getattr(config, "pad_token_id", None), do_sample=False,
                                                 temperature=1.0, top_p=1.0, top_k=0)
        # any load_state_dict that reaches this module — including PeftModel.load_state_dict recursing from a wrapper
        # (scripts/quick_start.py:542) — invalidates the packed engine
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate_engine())

    def _invalidate_engine(self):
        self._engine = None

    def _fingerprint(self):
        """Cheap staleness check for in-place parameter edits: tensor identities + version counters."""
        return tuple((id(p), p._version) for p in self.parameters())

    # ---- construction / weights --------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, config=None, torch_dtype=None, **kwargs):
        """Build the container and load an HF LLaMA/Qwen2 checkpoint directory (safetensors or .bin shards)."""
        if config is None:
            config = cls.config_class.from_pretrained(pretrained_model_name_or_path)
        model = cls(config, torch_dtype=torch_dtype, **{k: v for k, v in kwargs.items() if k == "max_ctx"})
        files = sorted(glob.glob(os.path.join(str(pretrained_model_name_or_path), "*.safetensors")))
        sd = {}
        if files:
            from safetensors.torch import load_file

            for f in files:
                sd.update(load_file(f))
        else:
            for f in sorted(glob.glob(os.path.join(str(pretrained_model_name_or_path), "pytorch_model*.bin"))):
                sd.update(torch.load(f, map_location="cpu"))
        if sd:
            if getattr(config, "tie_word_embeddings", False) and "lm_head.weight" not in sd and "model.embed_tokens.weight" in sd:
                sd["lm_head.weight"] = sd["model.embed_tokens.weight"]      # tied checkpoints ship no lm_head tensor
            res = model.load_state_dict(sd, strict=False)
            lost = [k for k in res.missing_keys if k.startswith(("model.layers.", "model.embed_tokens", "model.norm", "lm_head"))]
            if lost:
                raise RuntimeError(f"checkpoint {pretrained_model_name_or_path} lacks {len(lost)} decoder tensors, e.g. {lost[:4]}")
        gc = os.path.join(str(pretrained_model_name_or_path), "generation_config.json")
        if os.path.exists(gc):
            with open(gc) as f:
                for k, v in json.load(f).items():
                    setattr(model.generation_config, k, v)
        return model

    @classmethod
    def from_engine(cls, config, engine: CrabEngine):
        """Wrap an already-packed engine (weights live only in its device buffers; the parameter containers are
        created on the `meta` device).  Used when weights are produced directly on the GPU."""
        with torch.device("meta"):
            model = cls.__new__(cls)
            nn.Module.__init__(model)
            model.model = cls._model_cls(config, torch.bfloat16)
            model.lm_head = nn.Linear(config.hidden_size, engine.vocab, bias=False, dtype=torch.bfloat16)
        model.config, model.vocab_size, model.max_ctx = config, engine.vocab, engine.cfg.max_ctx
        model.pretraining_tp, model.is_avs_task = 1, False
        model._engine, model._target = engine, engine.dev
        model._engine_fp = None
        # no checkpoint, no generation_config: fixed-length greedy runs unless the caller passes eos_token_id
        model.generation_config = SimpleNamespace(max_new_tokens=20, eos_token_id=None, pad_token_id=engine.cfg.pad_token_id,
                                                  do_sample=False, temperature=1.0, top_p=1.0, top_k=0)
        model.model.pad_token_id = engine.cfg.pad_token_id
        ids = engine.cfg.special_ids
        model.SPECIAL_TOKEN_2_IDS, model.IDS_2_SPECIAL_TOKEN = dict(ids), {v: k for k, v in ids.items()}
        return model

    def get_model(self):
        return self.model

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def resize_token_embeddings(self, new_num_tokens: int):
        """Grow embed_tokens / lm_head (new rows: mean of the old ones — initialisation is irrelevant for inference
        because the fine-tuned checkpoint overwrites them, scripts/quick_start.py:537-554)."""
        old_e, old_h = self.model.embed_tokens.weight.data, self.lm_head.weight.data
        if new_num_tokens == old_e.shape[0]:
            return self.model.embed_tokens
        n = min(old_e.shape[0], new_num_tokens)
        e = nn.Embedding(new_num_tokens, old_e.shape[1], dtype=old_e.dtype)
        h = nn.Linear(old_h.shape[1], new_num_tokens, bias=False, dtype=old_h.dtype)
        with torch.no_grad():
            e.weight[:] = old_e.float().mean(0).to(old_e.dtype)
            h.weight[:] = old_h.float().mean(0).to(old_h.dtype)
            e.weight[:n] = old_e[:n]
            h.weight[:n] = old_h[:n]
        self.model.embed_tokens, self.lm_head = e, h
        self.config.vocab_size = self.vocab_size = new_num_tokens
        self._engine = None
        return e

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = {(k[len("base_model.model."):] if k.startswith("base_model.model.") else k): v for k, v in state_dict.items()}
        self._engine = None
        return nn.Module.load_state_dict(self, sd, strict=strict)

    def _lora_hparams(self):
        """(r, alpha, nums) of the hyper-LoRA wrap peft_hyper put on the linears (peft_hyper/tuners/lora.py:260-297), read from
        the wrapped modules themselves; None when the decoder is not wrapped."""
        for m in self.model.layers[0].self_attn.modules() if len(self.model.layers) else ():
            if hasattr(m, "lora_A") and hasattr(m, "lora_route"):
                r = int(getattr(m, "r", m.lora_A.weight.shape[0]))
                nums = int(getattr(m, "lora_num", m.lora_route.weight.shape[0]))
                alpha = getattr(m, "lora_alpha", None)
                if alpha is None and getattr(m, "scaling", None) is not None:
                    alpha = float(m.scaling) * r
                return r, (16 if alpha is None else alpha), nums
        return None

    # ---- device handling: parameters stay on the host as the load-time copy; the engine owns device memory ------
    def cuda(self, device=None):
        self._target = torch.device("cuda", torch.cuda.current_device() if device is None else
                                    (device if isinstance(device, int) else torch.device(device).index or 0))
        return self

    def npu(self, device=None):  # scripts/quick_start.py:558 (the reference was developed on Ascend)
        return self.cuda(device)

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)) and torch.device(a).type == "cuda":
                return self.cuda(torch.device(a).index)
        return nn.Module.to(self, *args, **kwargs)

    @property
    def device(self):
        # a wrapper's nn.Module.cuda() moves the parameters without calling our .cuda(): follow them
        if self._target.type != "cuda" and self.lm_head.weight.is_cuda:
            self._target = self.lm_head.weight.device
        return self._target

    @property
    def dtype(self):
        return torch.bfloat16

    def decoder_config(self) -> DecoderConfig:
        c = self.config
        hd = getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads
        d = DecoderConfig(hidden=c.hidden_size, inter=c.intermediate_size, layers=c.num_hidden_layers,
                          heads=c.num_attention_heads, kv_heads=getattr(c, "num_key_value_heads", None) or c.num_attention_heads,
                          head_dim=hd, vocab=self.lm_head.weight.shape[0],
                          rope_theta=float(getattr(c, "rope_theta", None) or (getattr(c, "rope_parameters", None) or {}).get("rope_theta", 10000.0)),
                          eps=c.rms_norm_eps, qkv_bias=self.model.qkv_bias)
        hp = self._lora_hparams()
        if hp is not None:
            if (hp[0], hp[2]) != (8, 3):
                raise ValueError(f"hyper-LoRA geometry r={hp[0]}, lora_nums={hp[2]} is not supported: the B200 kernels are built for "
                                 "the reference's shipped r=8 x 3 experts (11 router/A rows, 24 z columns per linear)")
            d.lora_r, d.lora_alpha, d.lora_nums = hp
        return d

    def engine(self) -> CrabEngine:
        meta = self.lm_head.weight.is_meta           # from_engine(): the weights live only in the engine
        if self._engine is not None and not meta and self._engine_fp is not None and self._engine_fp != self._fingerprint():
            self._engine = None                      # a parameter was edited in place since the engine was packed
        if self._engine is None:
            if self.device.type != "cuda":
                raise RuntimeError("call .cuda() / .npu() first: crab_b200 has no CPU execution path")
            # state_dict() of this container includes any hyper-LoRA tensors peft_hyper attached to the linears
            self._engine = CrabEngine(self.state_dict(), build_crab_config(self.decoder_config(), self, self.max_ctx),
                                      self._target)
            self._engine_fp = self._fingerprint()
        return self._engine

    # ---- forward / generate -------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, batch_input_ids=None, batch_labels=None, batch_X_modals=None, batch_task_names=None,
                input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                logits_to_keep: int = 0, **kwargs):
        """Inference forward with the reference's contract (models/unified_llama.py:47-161): `batch_*` inputs go through
        prepare_multimodal_inputs (which also builds the -100-masked labels), `inputs_embeds` / `input_ids` start a new sequence;
        the result carries fp32 `.logits` for ALL positions (b, S, vocab) — `logits_to_keep=1` keeps only the last one, like HF —
        and `.loss`, the shifted cross-entropy, when labels are given.  A (b, 1) `input_ids` after that continues the sequence
        (decode step, :125-127) and returns (b, 1, vocab).  There is no backward: training is out of scope."""
        eng = self.engine()
        if input_ids is not None and input_ids.shape[1] == 1 and getattr(eng, "cur_len", 0) > 0 and inputs_embeds is None \
                and batch_input_ids is None and past_key_values is not None:
            if not getattr(eng, "_fwd_decode", False):
                eng.begin_decode(input_ids.shape[0], use_graph=False)
                eng._fwd_decode = True
            eng.next_ids.copy_(input_ids[:, 0].to(eng.dev))
            logits, _ = eng.decode_step()
            return SimpleNamespace(logits=logits.unsqueeze(1).clone(), past_key_values=True, loss=None)
        if inputs_embeds is None and batch_input_ids is not None:
            prep = self.prepare_multimodal_inputs(batch_input_ids, batch_labels, batch_X_modals, batch_task_names)
            inputs_embeds = prep["inputs_embeds"]
            labels = prep["labels"] if labels is None else labels
        elif inputs_embeds is None and input_ids is not None:
            b, s = input_ids.shape
            inputs_embeds = self.encode_ids(input_ids).view(b, s, -1)
        eng._fwd_decode = False
        if logits_to_keep == 1 and labels is None:
            logits, _ = eng.prefill(inputs_embeds.clone())
            return SimpleNamespace(logits=logits.unsqueeze(1).clone(), past_key_values=True, loss=None)
        logits, loss = eng.forward_logits(inputs_embeds, labels)
        if logits_to_keep:
            logits = logits[:, -logits_to_keep:]
        return SimpleNamespace(logits=logits, past_key_values=True, loss=loss)

    def _eos_ids(self, eos_token_id, ignore_eos: bool):
        """HF semantics: an explicit `eos_token_id` wins, else generation_config.eos_token_id (from the checkpoint's
        generation_config.json, else the model config).  `ignore_eos=True` (extension) forces a fixed-length run."""
        if ignore_eos:
            return None
        if eos_token_id is None:
            eos_token_id = getattr(self.generation_config, "eos_token_id", None)
        if eos_token_id is None:
            eos_token_id = getattr(self.config, "eos_token_id", None) if not self.lm_head.weight.is_meta else None
        return eos_token_id

    def _pad_id(self, pad_token_id=None):
        for v in (pad_token_id, getattr(self.generation_config, "pad_token_id", None), self.model.pad_token_id):
            if v is not None:
                return int(v)
        return 0

    @torch.no_grad()
    def generate(self, batch_input_ids=None, batch_labels=None, batch_X_modals=None, batch_task_names=None, *,
                 inputs_embeds=None, max_new_tokens: Optional[int] = None, eos_token_id=None, pad_token_id=None, do_sample=None,
                 temperature=None, top_p=None, top_k=None, ignore_eos: bool = False, generator: Optional[torch.Generator] = None,
                 **kwargs):
        """Returns only the new ids (b, <= max_new_tokens), like HF generate with inputs_embeds (models/unified_llama.py:244-267).
        EOS: rows that emit an EOS id are padded with pad_token_id afterwards and decoding stops right after the last row
        finishes; the EOS id defaults to the checkpoint's generation_config / config exactly as in HF — quick_start.py passes
        only `use_cache` and `max_new_tokens`.  Greedy by default; `do_sample` (argument or generation_config, e.g.
        LLaMA-2-chat's T=0.6 / top-p 0.9) switches the token choice to temperature / top-k / top-p sampling."""
        eng = self.engine()
        if inputs_embeds is None:
            inputs_embeds, _, _ = eng.prepare_inputs(batch_input_ids, batch_X_modals)
        n = max_new_tokens or self.generation_config.max_new_tokens
        gc = self.generation_config
        sample = bool(getattr(gc, "do_sample", False) if do_sample is None else do_sample)
        sampling = None
        if sample:
            sampling = dict(temperature=float(temperature if temperature is not None else getattr(gc, "temperature", 1.0) or 1.0),
                            top_p=float(top_p if top_p is not None else getattr(gc, "top_p", 1.0) or 1.0),
                            top_k=int(top_k if top_k is not None else getattr(gc, "top_k", 0) or 0), generator=generator)
        return eng.generate_from_embeds(inputs_embeds, n, eos_token_id=self._eos_ids(eos_token_id, ignore_eos),
                                        pad_token_id=self._pad_id(pad_token_id), sampling=sampling)

    @torch.no_grad()
    def generate_avs(self, batch_input_ids=None, batch_labels=None, batch_X_modals=None, batch_task_names=None, *,
                     max_new_tokens: Optional[int] = None, forced_output_ids: Optional[torch.Tensor] = None, **kwargs):
        """Mirror of the reference's `generate_avs` (models/unified_llama.py:270-361), one sample per call as there
        (`mask_list = ...tolist()[0]  # bs == 1`): greedy generation with the last layer's hidden states kept, the hidden states
        paired with the generated `<mask_i>` tokens -> `pred_embeddings`, then `postprocess_seg` (the segmentation head) with the
        ViT taps of the sample's image.  Returns {'output_ids': (1, n)} plus 'pred_masks': [(num_classes, 224, 224)] when six mask
        embeddings were produced.  `forced_output_ids` (1, n) teacher-forces the generated sequence (tests)."""
        eng = self.engine()
        if eng.seg is None:
            raise RuntimeError("generate_avs needs the segmentation branch: init_multimodal_modules(segment_branch=True) and its weights")
        assert len(batch_input_ids) == 1, "generate_avs handles one sample per call, as the reference does"
        inputs = self.prepare_multimodal_inputs(batch_input_ids, None, batch_X_modals, batch_task_names,   # labels are unused here
                                                return_multi_scale_features=True, return_gt_mask=False)
        n = max_new_tokens or self.generation_config.max_new_tokens
        # same stopping rule as generate(): the reference calls HF generate here too (unified_llama.py:322-330), so the
        # sequence — and with it the hidden states that can pair with <mask_i> tokens — ends at EOS
        ids = eng.generate_from_embeds(inputs["inputs_embeds"], n, capture_hidden=6,
                                       teacher_tokens=None if forced_output_ids is None else forced_output_ids.to(eng.dev),
                                       eos_token_id=None if forced_output_ids is not None else
                                       self._eos_ids(kwargs.get("eos_token_id"), bool(kwargs.get("ignore_eos", False))),
                                       pad_token_id=self._pad_id(kwargs.get("pad_token_id")))
        output_ids = ids if forced_output_ids is None else forced_output_ids.to(eng.dev)
        result = {"output_ids": output_ids}
        mask_ids = [self.SPECIAL_TOKEN_2_IDS[f"<mask_{i}>"] for i in range(6)]
        pred = select_pred_embeddings(output_ids[0].tolist(), mask_ids, eng.hidden_prefill_tail[0], eng.hidden_steps[:, 0])
        if pred is None:
            return result
        masks = eng.seg.forward(pred.unsqueeze(0).contiguous(), inputs["multi_scale_image_features"], list(batch_task_names))
        result["pred_masks"] = masks
        return result

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        """Kept for peft_hyper.PeftModelForCausalLM, which wraps this attribute (peft_model.py:518-520)."""
        if past_key_values is not None:
            return {"input_ids": input_ids[:, -1:], "past_key_values": past_key_values, **kwargs}
        return {"input_ids": None if inputs_embeds is not None else input_ids, "inputs_embeds": inputs_embeds, **kwargs}
