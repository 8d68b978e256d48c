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
