"""CPU: the drop-in boundary exercised the way `scripts/quick_start.py:453-560` exercises it — `from_pretrained` on the
mirror class, the REFERENCE's own `peft_hyper.get_peft_model` wrapped around it, `init_multimodal_modules` on synthetic CLIP /
BEATs checkpoints, `initialize_MM_tokenizer`, `load_state_dict(finetune_weights, strict=False)` with the reference's key
names, then `model.generate(batch_input_ids=..., batch_X_modals=...)` through the PeftModel attribute tunnel.  The kernel
library is replaced by its CPU stand-in (tests/fake_ops.py); outputs are compared with the reference's golden run.
Needs the reference checkout (for `peft_hyper` and the checkpoint writers), so it only runs in the build container."""
import tempfile
from pathlib import Path

import pytest
import torch

import fake_ops
from helpers import load_golden, rel_l2
from oracle import ref_shims as R
from oracle import synth

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference checkout not present")


def _build(monkeypatch, segment_branch):
    from transformers import LlamaConfig

    from crab_b200 import engine
    from crab_b200.engine import QformerConfig
    from crab_b200.models import unified_arch
    from crab_b200.models.unified_llama import UnifiedForCausalLM

    monkeypatch.setattr(engine, "ops", fake_ops)
    monkeypatch.setattr(unified_arch, "ops", fake_ops, raising=False)
    monkeypatch.setattr(fake_ops, "MIN_K", 8)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    R.install_shims()
    from peft_hyper import LoraConfig, get_peft_model   # the reference's own wrapper

    g, case, sd_plain, ocfg, ids, X = load_golden("llama_small")
    ckpt = synth.synth_state_dict(g["manifest"], case["weight_seed"])          # reference key names (base_model.model.…)
    tmp = Path(tempfile.mkdtemp(prefix="crab_qs_"))
    clip_dir = R.make_clip_dir(tmp, image=case["image_size"], patch=case["patch_size"], **case["clip"])
    beats_pt = R.make_beats_ckpt(tmp, dict(R.BEATS_CFG_PUBLIC, **case["beats"]))
    llama_dir = tmp / "llama"
    config = LlamaConfig(**case["llama_cfg"])
    config.save_pretrained(llama_dir)

    # ---- scripts/quick_start.py:466-529, with the mirror class in place of models.unified_llama.UnifiedForCausalLM ----
    model = UnifiedForCausalLM.from_pretrained(str(llama_dir), config=config, torch_dtype=torch.bfloat16)
    peft_config = LoraConfig(task_type="CAUSAL_LM", target_modules="q_proj,k_proj,v_proj,o_proj,gate_proj,down_proj,up_proj".split(","),
                             inference_mode=False, r=8, lora_alpha=16, lora_dropout=0.05, lora_nums=3)
    model = get_peft_model(model, peft_config)
    base_vocab = config.vocab_size
    tok = R.FakeTokenizer(base_vocab)
    model.get_model().pad_token_id = 0
    model.get_model().init_multimodal_modules(
        visual_branch=True, audio_branch=True, segment_branch=segment_branch, d_model=case["d_model"], vit_ckpt_path=str(clip_dir),
        select_layer_list=list(case["select_layer_list"]), select_feature="patch", image_size=case["image_size"],
        patch_size=case["patch_size"], visual_query_token_nums=32, audio_query_token_nums=32, BEATs_ckpt_path=str(beats_pt),
        prompt_embed_dim=256, mask_decoder_transformer_depth=2, low_res_mask_size=112, avs_query_num=300, num_classes=1,
        query_generator_num_layers=2, dice_loss_weight=0.5, bce_loss_weight=2.0, use_vqgan=False,
        qformer_config=QformerConfig(inter=case["bert"].get("intermediate_size", 3072)))
    model.initialize_MM_tokenizer(tok, mask_token_nums=6, use_vqgan=False)
    assert len(tok) == base_vocab + 17 and model.SPECIAL_TOKEN_2_IDS["<video>"] == base_vocab + 3
    model.load_state_dict(ckpt, strict=False)
    if segment_branch:   # quick_start.py:545-554: a second checkpoint carries the seg_module weights
        seg_ckpt = synth.synth_state_dict({"base_model.model.model.seg_module." + k: v
                                           for k, v in unified_arch.seg_manifest(case["d_model"]).items()}, 77)
        model.load_state_dict(seg_ckpt, strict=False)
    model.eval()
    # ---- the engine is built from the wrapped model's state dict (CPU stand-in instead of .npu()) ----------------------------
    inner = model
    while not isinstance(inner, UnifiedForCausalLM):
        inner = inner.base_model if hasattr(inner, "base_model") and inner.base_model is not inner else inner.model
    inner._engine = engine.CrabEngine(inner.state_dict(), unified_arch.build_crab_config(inner.decoder_config(), inner, 512),
                                      torch.device("cpu"))
    return model, inner, g, case, ocfg, ids, X


def test_quick_start_construction_and_generate(monkeypatch):
    model, inner, g, case, ocfg, ids, X = _build(monkeypatch, segment_branch=False)
    assert inner._engine.lora and inner._engine.has_encoders and inner._engine.seg is None
    n_new = g["generated_ids"].shape[1]
    out = model.generate(batch_input_ids=ids, batch_labels=[None] * len(ids), batch_X_modals=X,
                         batch_task_names=["avqa"] * len(ids), use_cache=True, max_new_tokens=n_new)
    assert tuple(out.shape) == tuple(g["generated_ids"].shape)
    assert torch.equal(out[:, 0].cpu(), g["generated_ids"][:, 0])
    emb = model.prepare_multimodal_inputs(ids, None, X, ["avqa"] * len(ids))["inputs_embeds"]
    assert rel_l2(emb, g["inputs_embeds"]) < 3e-2


def test_quick_start_avs_branch(monkeypatch):
    """seg_branch=True: the seg_module containers take the reference's checkpoint keys and `model.generate_avs(**sample)`
    (quick_start.py:73-80) returns masks through the PeftModel tunnel."""
    from crab_b200 import seg

    monkeypatch.setattr(seg, "ops", fake_ops)
    model, inner, g, case, ocfg, ids, X = _build(monkeypatch, segment_branch=True)
    assert inner._engine.seg is not None
    gi = torch.Generator().manual_seed(3)
    image = torch.randn(1, 3, case["image_size"], case["image_size"], generator=gi)
    prompt = torch.randint(3, ocfg.base_vocab, (12,), generator=gi)
    prompt[4] = ocfg.special_ids["<image>"]
    m = [ocfg.special_ids[f"<mask_{i}>"] for i in range(6)]
    forced = torch.tensor([7, ocfg.special_ids["<mask_start>"]] + m + [ocfg.special_ids["<mask_end>"]]).view(1, -1)
    res = model.generate_avs(batch_input_ids=[prompt], batch_labels=[None], batch_X_modals=[{"<image>": image}],
                             batch_task_names=["s4"], use_cache=True, max_new_tokens=forced.shape[1], forced_output_ids=forced)
    assert torch.equal(res["output_ids"], forced) and tuple(res["pred_masks"][0].shape) == (1, 224, 224)
    assert torch.isfinite(res["pred_masks"][0]).all()
