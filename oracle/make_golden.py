"""TEST INFRASTRUCTURE — generate tests/golden/*.pt by running the REAL reference (`/root/reference`) in this
container (through oracle/ref_shims.py) on synthetic seeded weights and inputs.

    python -m oracle.make_golden            # rewrites tests/golden/

Each fixture holds: the case config, the state-dict manifest {key: shape} (weights are regenerated from
oracle/synth.py by name+seed, so no weights are committed), the input seed, and the reference's outputs
(inputs_embeds, per-stage encoder outputs, last-position logits, greedy ids).  tests/test_oracle_golden.py then
checks oracle/crab_oracle.py against these on any box (no reference needed).
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shims as R  # noqa: E402
from oracle import synth  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

CASES = {
    # small everything, but with the true head sizes (64 encoders / 128 decoder) and the hard-coded widths the
    # reference's init_multimodal_modules imposes (CLIP 1024, BEATs 768: models/unified_arch.py:78-85)
    "llama_small": dict(
        kind="llama",
        llama_cfg=dict(hidden_size=256, intermediate_size=384, num_hidden_layers=2, num_attention_heads=2,
                       num_key_value_heads=2, vocab_size=320, max_position_embeddings=1024, rms_norm_eps=1e-6,
                       rope_theta=10000.0),
        d_model=256, clip=dict(hidden=1024, inter=512, layers=3, heads=16), image_size=56, patch_size=14,
        select_layer_list=(1, 2, 2),
        beats=dict(encoder_layers=2, encoder_ffn_embed_dim=512),
        bert=dict(intermediate_size=512),
        frames=2, audio_segs=2, audio_len=98, prompt_len=24, new_tokens=8, weight_seed=7, input_seed=1,
    ),
    "llama_small_bs2": dict(
        kind="llama",
        llama_cfg=dict(hidden_size=256, intermediate_size=384, num_hidden_layers=2, num_attention_heads=2,
                       num_key_value_heads=2, vocab_size=320, max_position_embeddings=1024, rms_norm_eps=1e-6,
                       rope_theta=10000.0),
        d_model=256, clip=dict(hidden=1024, inter=512, layers=3, heads=16), image_size=56, patch_size=14,
        select_layer_list=(1, 2, 2),
        beats=dict(encoder_layers=2, encoder_ffn_embed_dim=512),
        bert=dict(intermediate_size=512),
        frames=2, audio_segs=2, audio_len=98, prompt_len=24, new_tokens=6, weight_seed=8, input_seed=2, bs=2,
        prompt_lens=(24, 31),
    ),
}


QWEN_CASES = {
    # Qwen2 backbone: q/k/v bias, grouped KV heads (4 query heads share 2 KV heads... here 4:2), rope_theta 1e6
    "qwen_small": dict(
        kind="qwen",
        llama_cfg=dict(hidden_size=512, intermediate_size=768, num_hidden_layers=2, num_attention_heads=4,
                       num_key_value_heads=2, vocab_size=352, max_position_embeddings=1024, rms_norm_eps=1e-6,
                       rope_theta=1000000.0, tie_word_embeddings=False),
        bs=2, seq_len=75, new_tokens=8, weight_seed=11, input_seed=5,
    ),
}


@torch.no_grad()
def run_qwen_case(name: str, case: dict) -> dict:
    from transformers import Qwen2ForCausalLM

    model = R.build_reference_qwen(qwen_cfg=case["llama_cfg"], lora=True)
    manifest = load_synth_weights(model, case["weight_seed"])
    g = torch.Generator(device="cpu").manual_seed(case["input_seed"])
    emb = torch.randn(case["bs"], case["seq_len"], case["llama_cfg"]["hidden_size"], generator=g)
    out = {"case": case, "manifest": manifest, "inputs_embeds": emb.clone()}
    inner = model.base_model.model  # the reference's UnifiedForCausalLM (Qwen)
    gen = Qwen2ForCausalLM.generate(inner, inputs_embeds=emb, max_new_tokens=case["new_tokens"], do_sample=False,
                                    eos_token_id=None, pad_token_id=0, use_cache=True)
    out["generated_ids"] = gen.clone()
    fo = model(inputs_embeds=emb, use_cache=True, output_hidden_states=True)
    out["prefill_last_logits"] = fo.logits[:, -1].float().clone()
    out["hidden_states"] = [h[:, -4:].clone() for h in fo.hidden_states]
    step = model(input_ids=gen[:, :1], past_key_values=fo.past_key_values, use_cache=True)
    out["step1_logits"] = step.logits[:, -1].float().clone()
    return out


def _beats_cfg(over: dict) -> dict:
    return dict(R.BEATS_CFG_PUBLIC, **over)


def load_synth_weights(model, seed: int):
    sd = model.state_dict()
    manifest = {k: tuple(v.shape) for k, v in sd.items()}
    new = synth.synth_state_dict(manifest, seed)
    for k, v in new.items():
        new[k] = v.to(sd[k].dtype)
    missing = model.load_state_dict(new, strict=True)
    return manifest


@torch.no_grad()
def run_case(name: str, case: dict) -> dict:
    model, tok = R.build_reference_model(
        llama_cfg=case["llama_cfg"], d_model=case["d_model"], clip=case["clip"], image_size=case["image_size"],
        patch_size=case["patch_size"], select_layer_list=case["select_layer_list"],
        beats_cfg=_beats_cfg(case["beats"]), bert=case["bert"], lora=True)
    manifest = load_synth_weights(model, case["weight_seed"])
    base_vocab = case["llama_cfg"]["vocab_size"]
    ids_map = model.SPECIAL_TOKEN_2_IDS
    bs = case.get("bs", 1)
    batch_ids, batch_X = [], []
    for i in range(bs):
        plen = case.get("prompt_lens", (case["prompt_len"],) * bs)[i]
        video, audio, ids = synth.synth_inputs(case["input_seed"] + 1000 * i, frames=case["frames"],
                                               image=case["image_size"], audio_segs=case["audio_segs"],
                                               audio_len=case["audio_len"], prompt_len=plen, base_vocab=base_vocab,
                                               video_id=ids_map["<video>"], audio_id=ids_map["<audio>"])
        batch_ids.append(ids)
        batch_X.append({"<video>": video, "<audio>": audio})
    inner = model.get_model()
    out = {"case": case, "manifest": manifest, "special_ids": dict(ids_map)}
    # stage dumps for sample 0
    v0, a0 = batch_X[0]["<video>"], batch_X[0]["<audio>"]
    vit_list = inner.visual_encoder(v0.unsqueeze(0))
    out["vit_taps"] = [t[0].clone() for t in vit_list]
    out["vl_out"] = inner.vl_projector(vit_list[-1])[0].clone()
    beats = inner.audio_encoder(a0.unsqueeze(0))
    out["beats_out"] = beats[0].clone()
    out["al_out"] = inner.al_projector(beats)[0].clone()
    prep = model.prepare_multimodal_inputs(batch_input_ids=batch_ids, batch_labels=[i.clone() for i in batch_ids],
                                           batch_X_modals=batch_X, batch_task_names=["avqa"] * bs)
    out["inputs_embeds"] = prep["inputs_embeds"].clone()
    out["attention_mask"] = prep["attention_mask"].clone()
    out["position_ids"] = prep["position_ids"].clone()
    # the reference's own generate(): HF greedy loop with inputs_embeds only
    gen = model.generate(batch_input_ids=batch_ids, batch_labels=[i.clone() for i in batch_ids], batch_X_modals=batch_X,
                         batch_task_names=["avqa"] * bs, use_cache=True, max_new_tokens=case["new_tokens"],
                         do_sample=False, eos_token_id=None, pad_token_id=0)
    out["generated_ids"] = gen.clone()
    # prefill logits + per-layer hidden states by a direct forward (what generate's step 0 runs)
    fo = model(inputs_embeds=prep["inputs_embeds"], use_cache=True, output_hidden_states=True)
    out["prefill_last_logits"] = fo.logits[:, -1].float().clone()
    out["hidden_states"] = [h[:, -4:].clone() for h in fo.hidden_states]  # last 4 positions of each layer
    # one teacher-forced decode step with the reference's own first token
    step = model(input_ids=gen[:, :1], past_key_values=fo.past_key_values, use_cache=True)
    out["step1_logits"] = step.logits[:, -1].float().clone()
    return out


def main():
    if not R.reference_available():
        raise SystemExit("reference checkout not found; goldens can only be regenerated in the build container")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    for name, case in list(CASES.items()) + list(QWEN_CASES.items()):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        out = run_qwen_case(name, case) if case["kind"] == "qwen" else run_case(name, case)
        # keep fixtures small: fp32 tensors only of modest size
        torch.save(out, GOLDEN / f"{name}.pt")
        sz = (GOLDEN / f"{name}.pt").stat().st_size
        print(f"{name}: wrote {sz / 1e6:.2f} MB; ids={out['generated_ids'].tolist()}")


if __name__ == "__main__":
    main()
