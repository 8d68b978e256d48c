"""TEST INFRASTRUCTURE — FULL-SHAPE golden: run the REAL reference (`/root/reference`, through oracle/ref_shims.py)
at the shapes BASELINE.json configs[0]/[1] and scripts/quick_start.py really use, and keep compact samples of every
stage so that the oracle and the CUDA path can both be pinned at those shapes on a box that has no reference.

    python -m oracle.make_golden_full        # ~10 min, ~20 GB RAM; writes tests/golden/full_llama7b.pt (~2 MB)

Shapes (SURVEY.md §8 head): CLIP ViT-L/14 at 224^2, 24 layers, taps hidden_states[14|22|23] (configs/unified_config.py:14);
BEATs 12 layers; both Q-Formers at bert-base widths; LLaMA-2-7B-dim decoder layers (hidden 4096, ff 11008, 32 heads),
vocab 32000 + 17, hyper-LoRA r=8 x 3 on all seven linears with NON-ZERO lora_B (oracle/synth.py).  The decoder depth is 4
layers (fp32 7B-dim layers are 0.8 GB each; every layer runs the same kernels at the same shapes).
Two input variants, same weights:
  "v8_a98"   8 frames,  audio (10, 98, 128)  -> T = 48, S = 64 + 574 = 638   (BASELINE configs[0]/[1])
  "v10_a198" 10 frames, audio (10, 198, 128) -> T = 96, S = 64 - 2 + 320 + 320 = 702   (what quick_start's dataset really
             produces: dataset/quick_start_dataset.py:83, 326-341, 715-731)
Per variant the fixture keeps, for sample rows chosen by `rows(n, k)`: the last ViT tap, VLProjector output, BEATs output,
ALProjector output, inputs_embeds; the last-position logits of the prompt pass and of 16 teacher-forced decode steps on a
fixed vocabulary subset (`logit_cols`) plus arg-max ids / top-2 margins over the full vocabulary; the reference's own
greedy ids; and the bf16 yardstick of SURVEY.md §7: the same decoder run by the reference in bf16 ("HF-bf16") against its
fp32 run (rel-L2 of the logits), so tests can assert  err(ours vs fp32) <= c * err(HF-bf16 vs fp32).
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shims as R  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.make_golden import _beats_cfg, load_synth_weights  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

CASE = dict(
    kind="llama",
    llama_cfg=dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=4, num_attention_heads=32,
                   num_key_value_heads=32, vocab_size=32000, max_position_embeddings=2048, rms_norm_eps=1e-6,
                   rope_theta=10000.0),
    d_model=4096, clip=dict(hidden=1024, inter=4096, layers=24, heads=16), image_size=224, patch_size=14,
    select_layer_list=(14, 22, 23), beats=dict(encoder_layers=12, encoder_ffn_embed_dim=3072),
    bert=dict(intermediate_size=3072), prompt_len=64, new_tokens=17, weight_seed=21,
    variants={"v8_a98": dict(frames=8, audio_segs=10, audio_len=98, input_seed=31),
              "v10_a198": dict(frames=10, audio_segs=10, audio_len=198, input_seed=32)},
    sample_rows=16, logit_stride=8,
)


def rows(n: int, k: int) -> torch.Tensor:
    """k row indices spread over [0, n) (deterministic; shared by the generator and the tests)."""
    return torch.linspace(0, n - 1, min(k, n)).round().long()


def sample(t: torch.Tensor, k: int) -> torch.Tensor:
    t = t.reshape(-1, t.shape[-1])
    return t[rows(t.shape[0], k)].float().clone()


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@torch.no_grad()
def main():
    if not R.reference_available():
        raise SystemExit("reference checkout not found; goldens can only be regenerated in the build container")
    case = CASE
    t0 = time.time()
    model, tok = R.build_reference_model(
        llama_cfg=case["llama_cfg"], d_model=case["d_model"], clip=case["clip"], image_size=case["image_size"],
        patch_size=case["patch_size"], select_layer_list=case["select_layer_list"],
        beats_cfg=_beats_cfg(case["beats"]), bert=case["bert"], lora=True)
    manifest = load_synth_weights(model, case["weight_seed"])
    print(f"reference built + synthetic weights loaded in {time.time() - t0:.0f} s", flush=True)
    ids_map = model.SPECIAL_TOKEN_2_IDS
    inner = model.get_model()
    K, n_new = case["sample_rows"], case["new_tokens"]
    out = {"case": case, "manifest": manifest, "special_ids": dict(ids_map), "variants": {}}
    keep = {}
    for vname, v in case["variants"].items():
        t0 = time.time()
        video, audio, ids = synth.synth_inputs(v["input_seed"], frames=v["frames"], image=case["image_size"],
                                               audio_segs=v["audio_segs"], audio_len=v["audio_len"],
                                               prompt_len=case["prompt_len"], base_vocab=case["llama_cfg"]["vocab_size"],
                                               video_id=ids_map["<video>"], audio_id=ids_map["<audio>"])
        X = [{"<video>": video, "<audio>": audio}]
        o = {}
        vit = inner.visual_encoder(video.unsqueeze(0))
        o["vit_tap_last"] = sample(vit[-1][0], K)
        o["vit_tap_first"] = sample(vit[0][0], K)
        o["vl_out"] = sample(inner.vl_projector(vit[-1])[0], K)
        beats = inner.audio_encoder(audio.unsqueeze(0))
        o["beats_out"] = sample(beats[0], K)
        o["al_out"] = sample(inner.al_projector(beats)[0], K)
        prep = model.prepare_multimodal_inputs(batch_input_ids=[ids], batch_labels=[ids.clone()], batch_X_modals=X,
                                               batch_task_names=["avqa"])
        emb = prep["inputs_embeds"]
        o["S"] = int(emb.shape[1])
        o["inputs_embeds"] = sample(emb[0], 2 * K)
        o["attention_mask"] = prep["attention_mask"].clone()
        o["position_ids"] = prep["position_ids"].clone()
        gen = model.generate(batch_input_ids=[ids], batch_labels=[ids.clone()], batch_X_modals=X, batch_task_names=["avqa"],
                             use_cache=True, max_new_tokens=n_new, do_sample=False, eos_token_id=None, pad_token_id=0)
        o["generated_ids"] = gen.clone()
        # teacher-forced run with the reference's own ids: last-position logits of the prompt pass and of every step
        fo = model(inputs_embeds=emb, use_cache=True)
        logits = [fo.logits[:, -1].float()]
        past = fo.past_key_values
        for s in range(n_new - 1):
            st = model(input_ids=gen[:, s:s + 1], past_key_values=past, use_cache=True)
            past = st.past_key_values
            logits.append(st.logits[:, -1].float())
        logits = torch.stack(logits, 0)[:, 0]                      # (n_new, V)
        assert torch.equal(logits.argmax(-1), gen[0]), "teacher-forced arg-max must reproduce generate()'s ids"
        cols = torch.arange(0, logits.shape[1], case["logit_stride"])
        top2 = logits.topk(2, dim=-1)
        o["logit_cols"] = cols
        o["logits_sub"] = logits[:, cols].clone()
        o["logits_norm"] = logits.norm(dim=-1).clone()
        o["top2_values"] = top2.values.clone()
        o["top2_indices"] = top2.indices.clone()
        out["variants"][vname] = o
        keep[vname] = (emb.clone(), gen.clone(), logits.clone())
        print(f"{vname}: S={o['S']} ids={gen.tolist()} ({time.time() - t0:.0f} s)", flush=True)
    # ---- the HF-bf16 yardstick: the reference's decoder in bf16 on the same inputs_embeds / teacher tokens ------------
    t0 = time.time()
    model = model.to(torch.bfloat16)
    for vname, (emb, gen, logits32) in keep.items():
        fo = model(inputs_embeds=emb.to(torch.bfloat16), use_cache=True)
        lg = [fo.logits[:, -1].float()]
        past = fo.past_key_values
        for s in range(n_new - 1):
            st = model(input_ids=gen[:, s:s + 1], past_key_values=past, use_cache=True)
            past = st.past_key_values
            lg.append(st.logits[:, -1].float())
        lg = torch.stack(lg, 0)[:, 0]
        o = out["variants"][vname]
        o["hf_bf16_logits_rel_l2"] = torch.tensor([rel_l2(lg[i], logits32[i]) for i in range(n_new)])
        o["hf_bf16_logits_max_abs"] = (lg - logits32).abs().amax(-1)
        o["hf_bf16_argmax"] = lg.argmax(-1)
        print(f"{vname}: HF-bf16 vs fp32 logits rel_l2 {o['hf_bf16_logits_rel_l2'].tolist()} ({time.time() - t0:.0f} s)", flush=True)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.save(out, GOLDEN / "full_llama7b.pt")
    print(f"wrote {(GOLDEN / 'full_llama7b.pt').stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
