"""GPU: the segmentation head through the C ABI — the five helper kernels against torch evaluations of the
same ops, and the whole `SegHead` against the oracle on the reference's golden weights (tests/golden/seg_small.pt)."""
import math
from pathlib import Path

import pytest
import torch

from oracle import seg_oracle as S
from oracle import synth
from oracle.make_seg_golden import seg_inputs

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "seg_small.pt"


def _g(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


@pytest.mark.parametrize("Nq,Nk,hd", [(300, 300, 32), (300, 256, 16), (1024, 300, 16), (300, 1024, 16), (7, 1, 32)])
def test_small_attn(cuda_dev, Nq, Nk, hd):
    from crab_b200 import ops

    H = 8
    g = _g(Nq + Nk + hd)
    q = torch.randn(Nq, H * hd + 8, generator=g).to(torch.bfloat16).to(cuda_dev)     # wider rows: strides are honoured
    k = torch.randn(Nk, H * hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    v = torch.randn(Nk, H * hd, generator=g).to(torch.bfloat16).to(cuda_dev)
    o = torch.zeros(Nq, H * hd, dtype=torch.bfloat16, device=cuda_dev)
    ops.small_attn(q, k, v, o, H, hd)
    sp = lambda t: t[:, : H * hd].float().reshape(t.shape[0], H, hd).transpose(0, 1)  # noqa: E731
    ref = (torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / math.sqrt(hd), -1) @ sp(v)).transpose(0, 1).reshape(Nq, H * hd)
    assert (o.float() - ref).abs().max().item() < 2e-2


def test_elementwise_rowmean_im2col_bilinear(cuda_dev):
    from crab_b200 import ops

    g = _g(3)
    a = torch.randn(50, 40, generator=g).to(torch.bfloat16).to(cuda_dev)
    b = torch.randn(50, 40, generator=g).to(torch.bfloat16).to(cuda_dev)
    gate = torch.randn(50, generator=g).to(cuda_dev)
    f = a.float()
    assert torch.equal(ops.elementwise(a, ops.EW_ADD, b=b), (f + b.float()).to(torch.bfloat16))
    assert torch.equal(ops.elementwise(a, ops.EW_ADD, b=b[:1]), (f + b[:1].float()).to(torch.bfloat16))
    assert torch.equal(ops.elementwise(a, ops.EW_RELU), torch.relu(f).to(torch.bfloat16))
    assert (ops.elementwise(a, ops.EW_GELU).float() - torch.nn.functional.gelu(f)).abs().max().item() < 1e-2
    ref = (torch.sigmoid(gate).unsqueeze(1) + 1) * f
    assert (ops.elementwise(a, ops.EW_GATE, gate=gate).float() - ref).abs().max().item() < 2e-2
    wide = torch.zeros(50, 64, dtype=torch.bfloat16, device=cuda_dev)                # strided output view
    ops.elementwise(a, ops.EW_RELU, out=wide[:, :40])
    assert torch.equal(wide[:, :40], torch.relu(f).to(torch.bfloat16)) and wide[:, 40:].abs().max().item() == 0
    x32 = torch.randn(33, 72, generator=g).to(cuda_dev)
    assert (ops.row_mean_f32(x32, 71) - x32[:, :71].mean(1)).abs().max().item() < 1e-5
    img = torch.randn(6 * 5, 16, generator=g).to(torch.bfloat16).to(cuda_dev)
    ref = torch.nn.functional.unfold(img.float().reshape(1, 6, 5, 16).permute(0, 3, 1, 2), 3, padding=1)[0]
    ref = ref.reshape(16, 9, 30).permute(2, 1, 0).reshape(30, 144)
    assert torch.equal(ops.im2col3x3(img, 6, 5).float(), ref)
    m = torch.randn(32 * 32, 8, generator=g).to(cuda_dev)
    nchw = m[:, :5].reshape(1, 32, 32, 5).permute(0, 3, 1, 2)
    for (ho, wo) in ((112, 112), (224, 224), (20, 50)):
        ref = torch.nn.functional.interpolate(nchw, (ho, wo), mode="bilinear", align_corners=False)[0]
        assert (ops.bilinear_f32(m, 32, 32, ho, wo, 5, nchw_out=True) - ref).abs().max().item() < 1e-5
        tok = ops.bilinear_f32(m, 32, 32, ho, wo, 5)
        assert (tok.reshape(ho, wo, 5).permute(2, 0, 1) - ref).abs().max().item() < 1e-5
    acc = torch.ones(112 * 112, 5, device=cuda_dev)
    ops.bilinear_f32(m, 32, 32, 112, 112, 5, out=acc, alpha=0.5, beta=1.0)
    ref = 1.0 + 0.5 * torch.nn.functional.interpolate(nchw, (112, 112), mode="bilinear", align_corners=False)[0]
    assert (acc.reshape(112, 112, 5).permute(2, 0, 1) - ref).abs().max().item() < 1e-5


def test_seg_head_vs_oracle(cuda_dev):
    """SegModule.forward on the GPU (s4 and avss heads) vs the CPU oracle, reference golden weights; also vs the reference's own
    masks stored in the fixture."""
    from crab_b200.seg import SegHead

    g = torch.load(GOLD)
    sd = synth.synth_state_dict(g["manifest"], g["weight_seed"])
    pred, feats = seg_inputs(g["input_seed"], g["d_model"])
    head = SegHead(sd, cuda_dev)
    out = head.forward(pred.to(torch.bfloat16).to(cuda_dev), [f.to(torch.bfloat16).to(cuda_dev) for f in feats], g["tasks"])
    torch.cuda.synchronize()
    assert tuple(out[0].shape) == (1, 224, 224) and tuple(out[1].shape) == (71, 224, 224)
    with torch.no_grad():
        ref = S.seg_module_forward(sd, pred.to(torch.bfloat16).float(), [f.to(torch.bfloat16).float() for f in feats], g["tasks"])
    for o, r, name in zip(out, ref, g["tasks"]):
        rel = ((o.cpu() - r).norm() / r.norm()).item()
        print(f"seg head {name}: rel_l2 vs oracle = {rel:.3e}")
        assert rel < 5e-2, rel
    rel = ((out[0].cpu() - g["mask_s4"]).norm() / g["mask_s4"].norm()).item()
    assert rel < 5e-2, rel
    # the binary masks the evaluation thresholds (mask > 0) agree with the reference's on all but the uncertain band
    agree = ((out[0].cpu() > 0) == (g["mask_s4"] > 0)).float().mean().item()
    assert agree > 0.99, agree


def test_generate_avs_end_to_end_vs_oracle(cuda_dev):
    """`UnifiedForCausalLM.generate_avs` on the GPU (ViT taps, hidden-state capture from prefill and the decode graph, the
    reference's mask-token pairing, segmentation head) vs the oracle restatement of models/unified_llama.py:270-361, with the
    generated sequence teacher-forced to contain the six mask tokens.  Same flow as tests/test_generate_avs_plumbing_cpu.py."""
    from helpers import engine_cfg, load_golden

    from crab_b200.engine import CrabEngine
    from crab_b200.models import unified_arch
    from crab_b200.models.unified_llama import UnifiedConfig, UnifiedForCausalLM

    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    D = ocfg.decoder.hidden
    sd = dict(sd)
    sd.update(synth.synth_state_dict({"model.seg_module." + k: v for k, v in unified_arch.seg_manifest(D).items()}, 77))
    grid = case["image_size"] // case["patch_size"]
    eng = CrabEngine(sd, engine_cfg(case, ocfg), cuda_dev)
    gi = torch.Generator().manual_seed(3)
    image = torch.randn(1, 3, case["image_size"], case["image_size"], generator=gi)
    prompt = torch.randint(3, ocfg.base_vocab, (12,), generator=gi)
    prompt[4] = ocfg.special_ids["<image>"]
    Xs = {"<image>": image}
    d = ocfg.decoder
    model = UnifiedForCausalLM.from_engine(UnifiedConfig(hidden_size=d.hidden, intermediate_size=d.inter, num_hidden_layers=d.layers,
                                                         num_attention_heads=d.heads, num_key_value_heads=d.kv_heads,
                                                         vocab_size=d.vocab), eng)
    m = [ocfg.special_ids[f"<mask_{i}>"] for i in range(6)]
    forced = torch.tensor([7, ocfg.special_ids["<mask_start>"]] + m + [ocfg.special_ids["<mask_end>"]])
    with torch.no_grad():
        ref = S.generate_avs(sd, prompt, Xs, ocfg, forced.numel(), "s4", forced_output_ids=forced, grid=grid)
    res = model.generate_avs(batch_input_ids=[prompt], batch_labels=None, batch_X_modals=[Xs], batch_task_names=["s4"],
                             max_new_tokens=forced.numel(), forced_output_ids=forced.view(1, -1))
    torch.cuda.synchronize()
    assert torch.equal(res["output_ids"].cpu(), forced.view(1, -1))
    out = res["pred_masks"][0].cpu()
    rel = ((out - ref["pred_masks"][0]).norm() / ref["pred_masks"][0].norm()).item()
    print(f"generate_avs pred_masks rel_l2 vs oracle = {rel:.3e}")
    assert tuple(out.shape) == (1, 224, 224) and rel < 8e-2, rel
    # ViT taps are the reference's multi-scale features
    with torch.no_grad():
        from oracle import crab_oracle as O
        taps = O.visual_encoder(sd, image.unsqueeze(0), ocfg.clip, ocfg.select_layers)
    for k in range(2):
        got = eng.image_taps[0][k].float().cpu()
        assert ((got - taps[k][0]).norm() / taps[k][0].norm()).item() < 2e-2
