"""CPU checks for the front-end row: the oracle restatement against outputs of the REAL reference
`dataset/audio_processor.preprocess` / HF CLIPImageProcessor (tests/golden/frontend.pt, oracle/make_frontend_golden.py),
and the product's host-side filter tables against the ones torchaudio builds for the reference."""
from pathlib import Path

import pytest
import torch

from oracle import frontend_oracle as F

GOLD = Path(__file__).resolve().parent / "golden" / "frontend.pt"


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def test_oracle_fbank_matches_the_reference(gold):
    w = F.synth_waveforms(gold["wave_seed"])
    fb = F.fbank_preprocess(w)
    assert fb.shape == gold["fbank"].shape == (4, 98, 128)
    assert (fb - gold["fbank"]).abs().max().item() <= 1e-5
    # the log floor is exercised (digital silence): log(eps) normalised
    floor = (torch.log(torch.tensor(F.EPS)) - 15.41663) / (2 * 6.55582)
    assert (gold["fbank"][2] == floor).float().mean().item() > 0.5


def test_oracle_fbank_ragged_length(gold):
    w = F.synth_waveforms(gold["wave_seed"])[:2, : gold["ragged_len"]]
    fb = F.fbank_preprocess(w)
    assert fb.shape == gold["fbank_ragged"].shape == (2, 1 + (gold["ragged_len"] - 400) // 160, 128)
    assert (fb - gold["fbank_ragged"]).abs().max().item() <= 1e-5


def test_oracle_pixel_values_match_hf_clip_processor(gold):
    pv = F.clip_pixel_values(F.synth_frames(gold["frame_seed"], 1))
    assert pv.shape == gold["pixel_values"].shape
    assert (pv - gold["pixel_values"]).abs().max().item() <= 2e-6


def test_host_tables_are_the_reference_tables(gold):
    from crab_b200.dataset import audio_processor as A

    assert torch.equal(A.mel_banks(), gold["mel_banks"])
    assert torch.equal(F.mel_filterbank()[:, :256], gold["mel_banks"])
    assert torch.equal(A.povey_window(), gold["window"])
    start, off, w = A.sparse_mel_rows(gold["mel_banks"])
    dense = torch.zeros_like(gold["mel_banks"])
    for m in range(128):
        n = int(off[m + 1] - off[m])
        dense[m, int(start[m]): int(start[m]) + n] = w[int(off[m]): int(off[m + 1])]
    assert torch.equal(dense, gold["mel_banks"])  # the sparse rows lose nothing


def test_oracle_resize_is_pillow_bit_exact_and_tables_agree():
    """The resize restatement against Pillow itself (the routine transformers 4.37's CLIPImageProcessor calls), bit for bit,
    including the committed golden; the product's vectorised coefficient tables against the oracle's loops."""
    import numpy as np
    from PIL import Image

    from crab_b200.dataset import image_processor as P

    rng = np.random.default_rng(0)
    for (H, W, oh, ow) in [(37, 53, 24, 34), (480, 640, 224, 298), (100, 80, 280, 224), (231, 517, 224, 501), (224, 300, 224, 300),
                           (224, 224, 224, 224)]:
        im = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(im).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(F.pil_resize_bicubic(im, oh, ow), ref), (H, W, oh, ow)
        for a, b in ((W, ow), (H, oh)):
            bt, kk = P.pil_bicubic_tables(a, b)
            for i, (lo, k) in enumerate(F.pil_coeffs(a, b)):
                assert bt[i, 0] == lo and bt[i, 1] == len(k) and kk[i, : len(k)].tolist() == k and not kk[i, len(k):].any()
    assert P.resize_output_size(300, 400, 224) == (224, 298) and P.resize_output_size(480, 360, 224) == (298, 224)
    gold = torch.load(GOLD)
    got = F.clip_resize_crop(F.synth_frames(gold["resize_seed"], 1, 8)[0].numpy().repeat(45, 0).repeat(61, 1)[:300, :400])
    assert np.array_equal(got, gold["resize_crop_300x400"].numpy())


def test_frontend_refuses_cpu_and_bad_shapes():
    from crab_b200.dataset import audio_processor as A
    from crab_b200.dataset.image_processor import ClipImageProcessorB200, frames_to_uint8_thwc
    from crab_b200.lib import CrabError

    with pytest.raises(CrabError):
        A.preprocess(torch.zeros(2, 16000), device=torch.device("cpu"))
    with pytest.raises(CrabError):
        A.preprocess(torch.zeros(2, 100), device=torch.device("cpu"))
    with pytest.raises(CrabError):
        frames_to_uint8_thwc(torch.zeros(2, 224, 224, 3))  # not uint8
    with pytest.raises(CrabError):
        ClipImageProcessorB200(device=torch.device("cpu")).preprocess(torch.zeros(1, 100, 100, 3, dtype=torch.uint8))
