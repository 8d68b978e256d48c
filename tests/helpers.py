"""Shared helpers for the parity tests: golden cases -> oracle config, engine config, weights and inputs."""
from pathlib import Path

import torch

from oracle import crab_oracle as O
from oracle import synth

GOLDEN = Path(__file__).resolve().parent / "golden"


def oracle_cfg(case, special_ids=None) -> O.CrabCfg:
    lc = case["llama_cfg"]
    dec = O.DecoderCfg(hidden=lc["hidden_size"], inter=lc["intermediate_size"], layers=lc["num_hidden_layers"],
                       heads=lc["num_attention_heads"], kv_heads=lc["num_key_value_heads"],
                       head_dim=lc["hidden_size"] // lc["num_attention_heads"], vocab=lc["vocab_size"] + 17,
                       rope_theta=lc.get("rope_theta", 10000.0), eps=lc.get("rms_norm_eps", 1e-6),
                       qkv_bias=case.get("kind") == "qwen")
    clip = O.ClipCfg(hidden=case["clip"]["hidden"], heads=case["clip"]["heads"], layers=case["clip"]["layers"],
                     patch=case["patch_size"])
    beats = O.BeatsCfg(layers=case["beats"]["encoder_layers"])
    cfg = O.CrabCfg(decoder=dec, clip=clip, beats=beats, qformer=O.QformerCfg(),
                    select_layers=tuple(case["select_layer_list"]),
                    image_tokens=(case["image_size"] // case["patch_size"]) ** 2, base_vocab=lc["vocab_size"])
    if special_ids is not None:
        for k in ("<video>", "<audio>", "<image>", "<mask_5>"):
            assert cfg.special_ids[k] == special_ids[k]
    return cfg


def engine_cfg(case, ocfg: O.CrabCfg, max_ctx=512):
    from crab_b200 import engine as E

    d = ocfg.decoder
    return E.CrabConfig(
        decoder=E.DecoderConfig(hidden=d.hidden, inter=d.inter, layers=d.layers, heads=d.heads, kv_heads=d.kv_heads,
                                head_dim=d.head_dim, vocab=d.vocab, rope_theta=d.rope_theta, eps=d.eps,
                                qkv_bias=d.qkv_bias),
        clip=E.ClipConfig(hidden=ocfg.clip.hidden, inter=case["clip"]["inter"], heads=ocfg.clip.heads,
                          layers=ocfg.clip.layers, patch=ocfg.clip.patch, image=case["image_size"]),
        beats=E.BeatsConfig(layers=ocfg.beats.layers, ffn=case["beats"].get("encoder_ffn_embed_dim", 3072)),
        qformer=E.QformerConfig(inter=case["bert"].get("intermediate_size", 3072)),
        select_layers=tuple(case["select_layer_list"]), pad_token_id=0, max_ctx=max_ctx,
        special_ids=dict(ocfg.special_ids))


def case_inputs(case, ocfg):
    bs = case.get("bs", 1)
    ids, X = [], []
    for i in range(bs):
        plen = case.get("prompt_lens", (case["prompt_len"],) * bs)[i]
        v, a, t = synth.synth_inputs(case["input_seed"] + 1000 * i, frames=case["frames"], image=case["image_size"],
                                     audio_segs=case["audio_segs"], audio_len=case["audio_len"], prompt_len=plen,
                                     base_vocab=ocfg.base_vocab, video_id=ocfg.special_ids["<video>"],
                                     audio_id=ocfg.special_ids["<audio>"])
        ids.append(t)
        X.append({"<video>": v, "<audio>": a})
    return ids, X


def load_golden(name):
    g = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    case = g["case"]
    sd = O.strip_peft_prefix(synth.synth_state_dict(g["manifest"], case["weight_seed"]))
    ocfg = oracle_cfg(case, g["special_ids"])
    ids, X = case_inputs(case, ocfg)
    return g, case, sd, ocfg, ids, X


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()
