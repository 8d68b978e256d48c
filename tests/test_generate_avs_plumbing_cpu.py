"""CPU: `generate_avs` end to end — engine plumbing (ViT taps out of prepare_inputs, last-layer hidden
states captured from prefill and every decode step, the reference's mask-token pairing and last-six rule, the segmentation
head) — with the kernel library replaced by its CPU stand-in (tests/fake_ops.py) and compared with the oracle restatement of
models/unified_llama.py:270-361 on seeded weights.  The same flow runs on the GPU in tests/test_seg_gpu.py."""
import pytest
import torch

import fake_ops
from helpers import engine_cfg, load_golden
from oracle import seg_oracle as S
from oracle import synth


def _setup(monkeypatch):
    from crab_b200 import engine, seg
    from crab_b200.models import unified_arch

    monkeypatch.setattr(engine, "ops", fake_ops)
    monkeypatch.setattr(seg, "ops", fake_ops)
    monkeypatch.setattr(fake_ops, "MIN_K", 8)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    D = ocfg.decoder.hidden
    seg_sd = synth.synth_state_dict({"model.seg_module." + k: v for k, v in unified_arch.seg_manifest(D).items()}, 77)
    sd = dict(sd)
    sd.update(seg_sd)
    grid = case["image_size"] // case["patch_size"]
    eng = engine.CrabEngine(sd, engine_cfg(case, ocfg), torch.device("cpu"))
    # an image-only AVS prompt: text, <image>, text
    gi = torch.Generator().manual_seed(3)
    image = torch.randn(1, 3, case["image_size"], case["image_size"], generator=gi)
    prompt = torch.randint(3, ocfg.base_vocab, (12,), generator=gi)
    prompt[4] = ocfg.special_ids["<image>"]
    return eng, sd, ocfg, grid, prompt, {"<image>": image}


@pytest.mark.parametrize("first_is_mask", [False, True])
def test_generate_avs_plumbing_matches_the_oracle(monkeypatch, first_is_mask):
    from crab_b200.models.unified_llama import select_pred_embeddings

    eng, sd, ocfg, grid, prompt, X = _setup(monkeypatch)
    assert eng.seg is not None and eng.seg.grid == grid
    m = [ocfg.special_ids[f"<mask_{i}>"] for i in range(6)]
    # forced generated sequences: "<mask_start> m0..m5 <mask_end>" starting at token 1 (the usual answer), and the corner case
    # where token 1 is already a mask token, which drags the prompt pass's hidden states into the last-six rule
    start, end = ocfg.special_ids["<mask_start>"], ocfg.special_ids["<mask_end>"]
    forced = torch.tensor(([start] + m + [end, 5]) if first_is_mask else ([7, start] + m + [end]))
    n = forced.numel()
    ref = S.generate_avs(sd, prompt, X, ocfg, n, "s4", forced_output_ids=forced, grid=grid)
    assert ref["pred_embeddings"] is not None
    taps = tuple(eng.cfg.select_layers[:2])
    emb, _, _ = eng.prepare_inputs([prompt], [X], want_image_taps=taps)
    assert len(eng.image_taps) == 1 and len(eng.image_taps[0]) == 2
    out = eng.generate_from_embeds(emb, n, capture_hidden=6, teacher_tokens=forced.view(1, -1))
    assert tuple(eng.hidden_steps.shape) == (n - 1, 1, ocfg.decoder.hidden) and eng.hidden_prefill_tail.shape[1] == 6
    pred = select_pred_embeddings(forced.tolist(), m, eng.hidden_prefill_tail[0], eng.hidden_steps[:, 0])
    assert pred is not None and tuple(pred.shape) == (6, ocfg.decoder.hidden)
    rel = ((pred.float() - ref["pred_embeddings"][0]).norm() / ref["pred_embeddings"][0].norm()).item()
    print("pred_embeddings rel_l2:", rel)
    assert rel < 3e-2
    feats = [t[: grid * grid].unsqueeze(0) for t in eng.image_taps[0]]
    masks = eng.seg.forward(pred.unsqueeze(0).contiguous(), feats, ["s4"])
    r = ref["pred_masks"][0]
    rel = ((masks[0] - r).norm() / r.norm()).item()
    print("pred_masks rel_l2:", rel)
    assert tuple(masks[0].shape) == (1, 224, 224) and rel < 8e-2


def test_select_pred_embeddings_rules():
    from crab_b200.models.unified_llama import select_pred_embeddings

    D = 8
    tail = torch.arange(6 * D, dtype=torch.float32).view(6, D)
    steps = 100 + torch.arange(9 * D, dtype=torch.float32).view(9, D)
    m = [50, 51, 52, 53, 54, 55]
    assert select_pred_embeddings([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], m, tail, steps) is None          # no mask tokens
    assert select_pred_embeddings([1, 2, 50, 51, 52, 53, 54, 9, 10, 11], m, tail, steps) is None     # only five
    got = select_pred_embeddings([1, 2, 50, 51, 52, 53, 54, 55, 10, 11], m, tail, steps)            # tokens 2..7 -> hidden 1..6
    assert torch.equal(got, steps[0:6])
    got = select_pred_embeddings([1, 50, 51, 52, 53, 54, 55, 50, 10, 11], m, tail, steps)           # seven (incl. prompt pass) -> last six
    assert torch.equal(got, torch.cat([steps[0:6]], 0))
    got = select_pred_embeddings([1, 50, 51, 52, 9, 9, 9, 9, 9, 9], m, tail, steps)                 # prompt pass + 2 steps
    assert torch.equal(got, torch.cat([tail[-4:], steps[0:2]], 0))


def test_mirror_generate_avs_call_matches_the_oracle(monkeypatch):
    """The public call `UnifiedForCausalLM.generate_avs(batch_input_ids=..., batch_X_modals=..., batch_task_names=...)`."""
    from crab_b200.models.unified_llama import UnifiedConfig, UnifiedForCausalLM

    eng, sd, ocfg, grid, prompt, X = _setup(monkeypatch)
    d = ocfg.decoder
    model = UnifiedForCausalLM.from_engine(UnifiedConfig(hidden_size=d.hidden, intermediate_size=d.inter, num_hidden_layers=d.layers,
                                                         num_attention_heads=d.heads, num_key_value_heads=d.kv_heads,
                                                         vocab_size=d.vocab), eng)
    m = [ocfg.special_ids[f"<mask_{i}>"] for i in range(6)]
    forced = torch.tensor([7, ocfg.special_ids["<mask_start>"]] + m + [ocfg.special_ids["<mask_end>"]])
    ref = S.generate_avs(sd, prompt, X, ocfg, forced.numel(), "avss", forced_output_ids=forced, grid=grid)
    res = model.generate_avs(batch_input_ids=[prompt], batch_labels=None, batch_X_modals=[X], batch_task_names=["avss"],
                             max_new_tokens=forced.numel(), forced_output_ids=forced.view(1, -1))
    assert torch.equal(res["output_ids"].cpu(), forced.view(1, -1))
    assert tuple(res["pred_masks"][0].shape) == (71, 224, 224)
    rel = ((res["pred_masks"][0] - ref["pred_masks"][0]).norm() / ref["pred_masks"][0].norm()).item()
    print("mirror generate_avs pred_masks rel_l2:", rel)
    assert rel < 8e-2
    # without mask tokens in the generated sequence there is no segmentation output (reference :341-351)
    res2 = model.generate_avs(batch_input_ids=[prompt], batch_labels=None, batch_X_modals=[X], batch_task_names=["avss"],
                              max_new_tokens=4)
    assert "pred_masks" not in res2 and tuple(res2["output_ids"].shape) == (1, 4)


def test_select_pred_embeddings_equals_the_literal_reference_loop():
    """Random generated sequences: the tail-based selection equals the reference's literal code (models/unified_llama.py:
    331-351: zip(mask_list, hidden_states), cat along the sequence axis, last six) evaluated on full hidden states."""
    from crab_b200.models.unified_llama import select_pred_embeddings

    g = torch.Generator().manual_seed(0)
    D, S, mask_ids = 4, 9, [90, 91, 92, 93, 94, 95]
    for trial in range(300):
        n = int(torch.randint(2, 14, (1,), generator=g))
        ids = [int(v) for v in torch.randint(85, 99, (n,), generator=g)]
        hidden = [torch.randn(1, S, D, generator=g)] + [torch.randn(1, 1, D, generator=g) for _ in range(n - 1)]
        # --- literal restatement of the reference lines
        mask_list = [int(t in mask_ids) for t in ids[1:]]
        pred = [hs for item, hs in zip(mask_list, hidden) if item == 1]
        ref = None
        if pred:
            cat = torch.cat(pred, dim=1)
            if cat.shape[1] > 6:
                ref = cat[:, -6:]
            elif cat.shape[1] == 6:
                ref = cat
        # --- ours: only the prompt pass's last six rows and one row per step are kept by the engine
        got = select_pred_embeddings(ids, mask_ids, hidden[0][0, -6:], torch.cat(hidden[1:], 0)[:, 0] if n > 1 else torch.empty(0, D))
        if ref is None:
            assert got is None, (trial, ids)
        else:
            assert got is not None and torch.equal(got, ref[0]), (trial, ids)
