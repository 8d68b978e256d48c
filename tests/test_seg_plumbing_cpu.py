"""CPU: the host-side plumbing of the segmentation head (crab_b200/seg.py) — weight packing, token-major layouts, the folded
query generator, pixel-shuffle indices, K / N padding — run against a CPU stand-in of the kernel library (tests/fake_ops.py,
which enforces the C ABI's argument checks) and compared with the oracle on the reference's golden weights.  The kernels
themselves are checked on the GPU (tests/test_seg_gpu.py)."""
from pathlib import Path

import torch

import fake_ops
from oracle import seg_oracle as S
from oracle import synth
from oracle.make_seg_golden import seg_inputs

GOLD = Path(__file__).resolve().parent / "golden" / "seg_small.pt"


def test_fake_kernels_use_the_reference_formulas():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(32 * 32, 8, generator=g)
    for (ho, wo) in ((112, 112), (50, 70), (16, 16)):
        ref = torch.nn.functional.interpolate(x.reshape(1, 32, 32, 8).permute(0, 3, 1, 2), (ho, wo), mode="bilinear", align_corners=False)[0]
        got = fake_ops.bilinear_f32(x, 32, 32, ho, wo, 8, nchw_out=True)
        assert (got - ref).abs().max().item() < 1e-5
    img = torch.randn(6 * 5, 8, generator=g).to(torch.bfloat16)
    ref = torch.nn.functional.unfold(img.float().reshape(1, 6, 5, 8).permute(0, 3, 1, 2), 3, padding=1)[0]   # [(c, ky, kx), hw]
    ref = ref.reshape(8, 9, 30).permute(2, 1, 0).reshape(30, 72)                                              # -> [hw, (tap, c)]
    assert torch.equal(fake_ops.im2col3x3(img, 6, 5).float(), ref)


def test_seg_head_plumbing_matches_the_oracle(monkeypatch):
    from crab_b200 import seg

    monkeypatch.setattr(seg, "ops", fake_ops)
    g = torch.load(GOLD)
    sd = synth.synth_state_dict(g["manifest"], g["weight_seed"])
    pred, feats = seg_inputs(g["input_seed"], g["d_model"])
    head = seg.SegHead(sd, torch.device("cpu"))
    out = head.forward(pred.to(torch.bfloat16), [f.to(torch.bfloat16) for f in feats], g["tasks"])
    assert tuple(out[0].shape) == (1, 224, 224) and tuple(out[1].shape) == (71, 224, 224)
    with torch.no_grad():
        ref = S.seg_module_forward(sd, pred.to(torch.bfloat16).float(), [f.to(torch.bfloat16).float() for f in feats], g["tasks"])
    for o, r in zip(out, ref):
        rel = ((o - r).norm() / r.norm()).item()
        print("rel_l2 vs oracle:", rel)
        assert rel < 5e-2, rel
