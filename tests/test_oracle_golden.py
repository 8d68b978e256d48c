"""Pin the CPU oracle (oracle/crab_oracle.py) against outputs of the REAL reference recorded in tests/golden/
(generated in the build container by oracle/make_golden.py).  Runs anywhere; no reference checkout needed."""
from pathlib import Path

import pytest
import torch

from oracle import crab_oracle as O
from oracle import synth

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = sorted(p.stem for p in GOLDEN.glob("llama_*.pt"))


def cfg_from_case(case, special_ids) -> O.CrabCfg:
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
    assert cfg.special_ids["<video>"] == special_ids["<video>"] and cfg.special_ids["<audio>"] == special_ids["<audio>"]
    assert cfg.special_ids["<mask_5>"] == special_ids["<mask_5>"]
    return cfg


def load_case(name):
    g = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    case = g["case"]
    sd = O.strip_peft_prefix(synth.synth_state_dict(g["manifest"], case["weight_seed"]))
    cfg = cfg_from_case(case, g["special_ids"])
    bs = case.get("bs", 1)
    ids, X = [], []
    for i in range(bs):
        plen = case.get("prompt_lens", (case["prompt_len"],) * bs)[i]
        v, a, t = synth.synth_inputs(case["input_seed"] + 1000 * i, frames=case["frames"], image=case["image_size"],
                                     audio_segs=case["audio_segs"], audio_len=case["audio_len"], prompt_len=plen,
                                     base_vocab=cfg.base_vocab, video_id=cfg.special_ids["<video>"],
                                     audio_id=cfg.special_ids["<audio>"])
        ids.append(t)
        X.append({"<video>": v, "<audio>": a})
    return g, case, sd, cfg, ids, X


def _close(a, b, tol=2e-4):
    err = (a.float() - b.float()).abs().max().item()
    scale = b.float().abs().max().item() + 1e-9
    assert err / scale < tol, f"rel max err {err / scale:.3e} (abs {err:.3e}, scale {scale:.3e})"


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    g, case, sd, cfg, ids, X = load_case(name)
    with torch.no_grad():
        taps = O.visual_encoder(sd, X[0]["<video>"].unsqueeze(0), cfg.clip, cfg.select_layers)
        for t, ref in zip(taps, g["vit_taps"]):
            _close(t[0], ref)
        _close(O.vl_projector(sd, taps[-1], cfg.qformer, cfg.image_tokens)[0], g["vl_out"])
        beats = O.audio_encoder(sd, X[0]["<audio>"].unsqueeze(0), cfg.beats)
        _close(beats[0], g["beats_out"])
        _close(O.al_projector(sd, beats, cfg.qformer)[0], g["al_out"])
        prep = O.prepare_multimodal_inputs(sd, ids, X, cfg)
        _close(prep["inputs_embeds"], g["inputs_embeds"])
        assert torch.equal(prep["attention_mask"], g["attention_mask"])
        assert torch.equal(prep["position_ids"], g["position_ids"])
        h, cache, hiddens = O.decoder_forward(sd, prep["inputs_embeds"], cfg.decoder, collect_hidden=True)
        # HF's hidden_states tuple: embeddings, then each layer's output; the last entry is post-final-norm
        for i, ref in enumerate(g["hidden_states"][:-1]):
            _close(hiddens[i][:, -4:], ref)
        _close(h[:, -4:], g["hidden_states"][-1])
        logits = O.lm_head(sd, h[:, -1])
        _close(logits, g["prefill_last_logits"])
        first = g["generated_ids"][:, :1]
        h1, cache = O.decoder_forward(sd, sd["model.embed_tokens.weight"][first[:, 0]].unsqueeze(1), cfg.decoder, cache)
        _close(O.lm_head(sd, h1[:, -1]), g["step1_logits"])
        gen, _ = O.greedy_generate(sd, prep["inputs_embeds"], cfg.decoder, case["new_tokens"])
        assert torch.equal(gen, g["generated_ids"]), (gen, g["generated_ids"])


def test_beats_bucket_table_properties():
    b = O.beats_relative_buckets(48, 48, 320, 800)
    assert b.shape == (48, 48) and int(b.min()) >= 0 and int(b.max()) < 320
    assert torch.equal(torch.diagonal(b), torch.zeros(48, dtype=torch.long))
    # bidirectional: positive offsets live in the upper half of the bucket range
    assert int(b[0, 1]) == 160 + 1 and int(b[1, 0]) == 1


def test_hyper_lora_reduces_to_linear_when_B_is_zero():
    torch.manual_seed(0)
    cfg = O.DecoderCfg(hidden=32, inter=64, layers=1, heads=1, kv_heads=1, head_dim=32, vocab=50)
    sd = {"l.weight": torch.randn(16, 32), "l.lora_A.weight": torch.randn(8, 32), "l.lora_route.weight": torch.randn(3, 32)}
    for i in range(3):
        sd[f"l.lora_B{i}.weight"] = torch.zeros(16, 8)
    x = torch.randn(2, 5, 32)
    assert torch.allclose(O.hyper_lora_linear(x, sd, "l", cfg), x @ sd["l.weight"].t())


def test_oracle_qwen_matches_reference():
    """Qwen2 backbone (q/k/v bias, GQA, rope_theta from config) vs the reference's unified_qwen + peft_hyper."""
    g = torch.load(GOLDEN / "qwen_small.pt", weights_only=False)
    case = g["case"]
    sd = O.strip_peft_prefix(synth.synth_state_dict(g["manifest"], case["weight_seed"]))
    lc = case["llama_cfg"]
    dec = O.DecoderCfg(hidden=lc["hidden_size"], inter=lc["intermediate_size"], layers=lc["num_hidden_layers"],
                       heads=lc["num_attention_heads"], kv_heads=lc["num_key_value_heads"],
                       head_dim=lc["hidden_size"] // lc["num_attention_heads"], vocab=lc["vocab_size"],
                       rope_theta=lc["rope_theta"], eps=lc["rms_norm_eps"], qkv_bias=True)
    with torch.no_grad():
        h, cache, hiddens = O.decoder_forward(sd, g["inputs_embeds"], dec, collect_hidden=True)
        for i, ref in enumerate(g["hidden_states"][:-1]):
            _close(hiddens[i][:, -4:], ref)
        _close(O.lm_head(sd, h[:, -1]), g["prefill_last_logits"])
        first = g["generated_ids"][:, 0]
        h1, cache = O.decoder_forward(sd, sd["model.embed_tokens.weight"][first].unsqueeze(1), dec, cache)
        _close(O.lm_head(sd, h1[:, -1]), g["step1_logits"])
        gen, _ = O.greedy_generate(sd, g["inputs_embeds"], dec, case["new_tokens"])
    assert torch.equal(gen, g["generated_ids"])
