"""GPU parity for the front-end kernels through the C ABI: Kaldi fbank and the fused uint8
rescale/normalise/patchify, against the oracle, the reference's golden outputs and size-independent properties."""
import math
from pathlib import Path

import pytest
import torch

from oracle import frontend_oracle as F

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden" / "frontend.pt"
FB_MEAN_TOL = 2e-5  # mean abs error on the normalised log-mel values


def _check_fbank(out, ref32, wave):
    """Two fp32 FFTs (pocketfft in the reference, radix-2 in shared memory here) round differently; bins 100 dB below a
    strong tone, or frames riding on a large DC drift, show it after the log (the reference itself is up to 9e-4 away from
    the float64 evaluation on the drift waveform).  So: tiny mean error against the reference, and a worst-case error
    against float64 in the same class as the reference's own (measured on B200: 0.5-5x of it, rounding luck)."""
    out = out.cpu()
    f64 = F.fbank_preprocess(wave.cpu(), dtype=torch.float64)
    e_ref = (ref32.double() - f64).abs().max().item()
    e_out = (out.double() - f64).abs().max().item()
    assert (out - ref32).abs().mean().item() < FB_MEAN_TOL
    assert e_out <= max(8.0 * e_ref, 2e-4), (e_out, e_ref)
    assert (out - ref32).abs().max().item() <= 2.0 * (e_out + e_ref) + 1e-6


def test_fbank_matches_reference_golden(cuda_dev):
    from crab_b200.dataset import audio_processor as A

    gold = torch.load(GOLD)
    w = F.synth_waveforms(gold["wave_seed"])
    out = A.preprocess(w.to(cuda_dev))
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (4, 98, 128)
    _check_fbank(out, gold["fbank"], w)
    # waveforms without a dominant tone agree to fp32 round-off outright
    assert (out[0].cpu() - gold["fbank"][0]).abs().max().item() < 1e-4
    wr = w[:2, : gold["ragged_len"]].contiguous()
    _check_fbank(A.preprocess(wr), gold["fbank_ragged"], wr)       # host input, ragged length
    # strided rows (a view into a longer buffer) go through wave_stride
    _check_fbank(A.preprocess(w.to(cuda_dev)[:, :8000]), F.fbank_preprocess(w[:, :8000]), w[:, :8000])


def test_fbank_full_batch_properties(cuda_dev):
    """BASELINE size (32 samples x 10 one-second segments): frame-shift equivariance, gain law, oracle on a sample."""
    from crab_b200.dataset import audio_processor as A

    g = torch.Generator().manual_seed(3)
    w = (0.2 * torch.randn(320, 16000 + 160, generator=g)).to(cuda_dev)
    a = A.preprocess(w[:, :16000])
    b = A.preprocess(w[:, 160:16160])
    assert tuple(a.shape) == (320, 98, 128)
    assert (a[:, 1:] - b[:, :-1]).abs().max().item() < 1e-6          # shifting by one hop shifts the frames by one
    c = A.preprocess(2.0 * w[:, :16000])
    floor = (math.log(F.EPS) - 15.41663) / (2 * 6.55582)                # empty mel filters (e.g. bin 3) sit on the log floor
    live = a > floor + 1e-3
    assert live.float().mean().item() > 0.98
    gain = (c - a - math.log(4.0) / (2 * 6.55582))[live].abs().max().item()
    assert gain < 1e-5, gain                                            # power x4 -> log-mel + log 4
    ref = F.fbank_preprocess(w[7:9, :16000].cpu())
    assert (a[7:9].cpu() - ref).abs().max().item() < 1e-4


def test_patchify_u8_and_normalize_u8(cuda_dev):
    from crab_b200 import ops
    from crab_b200.dataset.image_processor import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD, ClipImageProcessorB200

    gold = torch.load(GOLD)
    frames = F.synth_frames(gold["frame_seed"], 1)
    pv = ClipImageProcessorB200().preprocess(frames)["pixel_values"]
    assert pv.is_cuda and (pv.cpu() - gold["pixel_values"]).abs().max().item() < 2e-6  # HF CLIPImageProcessor output
    f3 = F.synth_frames(9, 3).to(cuda_dev)
    kpad = (3 * 14 * 14 + 7) // 8 * 8
    fused = ops.patchify_u8(f3, 14, kpad, OPENAI_CLIP_MEAN, OPENAI_CLIP_STD)
    two_step = ops.patchify(ops.normalize_u8(f3, OPENAI_CLIP_MEAN, OPENAI_CLIP_STD), 14, kpad)
    assert torch.equal(fused, two_step)                                 # same arithmetic, same rounding
    ref = F.clip_pixel_values(f3.cpu())                                 # oracle, then the reference's im2col order
    ref = ref.unfold(2, 14, 14).unfold(3, 14, 14).permute(0, 2, 3, 1, 4, 5).reshape(3 * 256, 588).to(torch.bfloat16)
    d = (fused[:, :588].float().cpu() - ref.float()).abs()
    assert (d > 0).float().mean().item() < 1e-3 and d.max().item() <= 2 ** -6   # <= 1 bf16 ulp on a few boundary cases
    assert fused[:, 588:].abs().max().item() == 0


def test_engine_accepts_raw_uint8_frames(cuda_dev):
    """generate() fed with decoder-format uint8 frames gives the same ids as with the processor's fp32 pixel_values."""
    from helpers import engine_cfg, load_golden

    from crab_b200.engine import CrabEngine

    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    eng = CrabEngine(sd, engine_cfg(case, ocfg), cuda_dev)
    size = case["image_size"]
    u8 = F.synth_frames(21, case["frames"], size)
    X_u8 = [{"<video>": u8, "<audio>": X[0]["<audio>"]}]
    X_f32 = [{"<video>": F.clip_pixel_values(u8), "<audio>": X[0]["<audio>"]}]
    e1, _, _ = eng.prepare_inputs(ids, X_u8)
    e1 = e1.clone()
    e2, _, _ = eng.prepare_inputs(ids, X_f32)
    rel = ((e1.float() - e2.float()).norm() / e2.float().norm()).item()
    assert rel < 2e-3, rel
    o1 = eng.generate(ids, X_u8, 6)
    o2 = eng.generate(ids, X_f32, 6)
    assert torch.equal(o1, o2)


@pytest.mark.parametrize("H,W", [(300, 400), (480, 360), (150, 224), (224, 300), (640, 480), (96, 131)])
def test_resize_crop_is_pillow_bit_exact(cuda_dev, H, W):
    """Shortest-edge bicubic resize + centre crop on the GPU == PIL.Image.resize(BICUBIC) + crop, byte for byte (up- and
    down-scaling, either orientation, one-axis-only cases), and the processor's pixel_values follow from it."""
    import numpy as np
    from PIL import Image

    from crab_b200.dataset.image_processor import ClipImageProcessorB200, resize_output_size

    rng = np.random.default_rng(H * 1000 + W)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    frames[0, : H // 2] = 255           # saturated regions: the bicubic overshoot must clip exactly as Pillow clips
    frames[0, H // 2:] = 0
    proc = ClipImageProcessorB200()
    got = proc.to_device_uint8(torch.from_numpy(frames)).cpu().numpy()
    nh, nw = resize_output_size(H, W, 224)
    top, left = (nh - 224) // 2, (nw - 224) // 2
    for i in range(3):
        ref = np.asarray(Image.fromarray(frames[i]).resize((nw, nh), Image.BICUBIC))[top:top + 224, left:left + 224]
        assert np.array_equal(got[i], ref), (i, np.abs(got[i].astype(int) - ref.astype(int)).max())
        assert np.array_equal(F.clip_resize_crop(frames[i]), ref)
    pv = proc.preprocess(torch.from_numpy(frames))["pixel_values"].cpu()
    assert (pv - F.clip_pixel_values(torch.from_numpy(got))).abs().max().item() < 2e-6
