"""Mirror of the reference's `dataset/audio_processor.py::preprocess` (lines 29-41) on the B200.

Same name, arguments and result as the reference function — waveforms in [-1, 1] at 16 kHz in, normalised 128-bin Kaldi
log-mel filterbank frames out — but the whole pipeline (framing, DC removal, pre-emphasis, Povey window, 512-point FFT,
mel filters, log, normalisation) is one CUDA kernel (`crab_kaldi_fbank`, crab_b200/csrc/frontend.cu) and the result stays
on the device, ready for `UnifiedForCausalLM.generate(batch_X_modals=[{'<audio>': fbank, ...}])`.

The filter tables below follow `torchaudio.compliance.kaldi.get_mel_banks` / `_feature_window_function` (the third-party
routine the reference calls) operation by operation in fp32, so the weights are the ones the reference multiplies with.
There is no CPU fallback: a non-CUDA `device` raises.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple, Union

import torch

from .. import ops
from ..lib import CrabError

SAMPLE_RATE = 16000
FRAME_LENGTH_MS, FRAME_SHIFT_MS = 25.0, 10.0
WINDOW, SHIFT, PADDED = 400, 160, 512
NUM_MEL = 128
LOW_FREQ, HIGH_FREQ = 20.0, 0.0

_tables: Dict[Tuple[int, int], Tuple[torch.Tensor, ...]] = {}


def _mel_scale_scalar(freq: float) -> float:
    return 1127.0 * math.log(1.0 + freq / 700.0)


def mel_banks(num_bins: int = NUM_MEL, padded: int = PADDED, sample_freq: float = SAMPLE_RATE, low_freq: float = LOW_FREQ,
              high_freq: float = HIGH_FREQ) -> torch.Tensor:
    """fp32 [num_bins, padded // 2] triangular filters on the mel scale (Kaldi `MelBanks`, no VTLN)."""
    num_fft_bins = padded // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / padded
    mel_low, mel_high = _mel_scale_scalar(low_freq), _mel_scale_scalar(high_freq)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down))


def povey_window(n: int = WINDOW) -> torch.Tensor:
    return torch.hann_window(n, periodic=False, dtype=torch.float32).pow(0.85)


def sparse_mel_rows(banks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Dense [n_mel, n_bins] -> (start[n_mel], off[n_mel + 1], packed weights): every row's non-zero span."""
    starts, offs, ws = [], [0], []
    for row in banks:
        nz = torch.nonzero(row > 0).flatten()
        if nz.numel() == 0:
            starts.append(0)
            offs.append(offs[-1])
            continue
        a, b = int(nz[0]), int(nz[-1]) + 1
        starts.append(a)
        ws.append(row[a:b])
        offs.append(offs[-1] + (b - a))
    return (torch.tensor(starts, dtype=torch.int32), torch.tensor(offs, dtype=torch.int32),
            torch.cat(ws).to(torch.float32).contiguous())


def _device_tables(dev: torch.device):
    key = (dev.index or 0, NUM_MEL)
    if key not in _tables:
        start, off, w = sparse_mel_rows(mel_banks())
        k = torch.arange(PADDED // 2, dtype=torch.float64) * (2.0 * math.pi / PADDED)
        twiddle = torch.stack([torch.cos(k), -torch.sin(k)], 1).to(torch.float32).contiguous()  # exp(-2 pi i k / 512)
        _tables[key] = (povey_window().to(dev), twiddle.to(dev), start.to(dev), off.to(dev), w.to(dev))
    return _tables[key]


def preprocess(source: Union[torch.Tensor, Sequence[torch.Tensor]], fbank_mean: float = 15.41663, fbank_std: float = 6.55582,
               device: Optional[torch.device] = None) -> torch.Tensor:
    """source: (n, L) waveforms (or a list of equal-length 1-D waveforms) -> (n, 1 + (L - 400) // 160, 128) fp32 on the GPU
    == ((kaldi.fbank(w * 2**15, num_mel_bins=128, ...) - fbank_mean) / (2 * fbank_std)) for every w."""
    if not isinstance(source, torch.Tensor):
        source = torch.stack([torch.as_tensor(s) for s in source], 0)
    if source.dim() == 1:
        source = source.unsqueeze(0)
    if source.dim() != 2:
        raise CrabError(f"preprocess wants (n, L) waveforms, got shape {tuple(source.shape)}")
    if source.shape[1] < WINDOW:
        raise CrabError(f"waveforms shorter than one {WINDOW}-sample frame")
    dev = device or (source.device if source.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if dev.type != "cuda":
        raise CrabError("crab_b200.dataset.audio_processor.preprocess runs on the GPU only (no CPU path)")
    wave = source.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
    window, twiddle, start, off, w = _device_tables(dev)
    return ops.kaldi_fbank(wave, window, twiddle, start, off, w, in_scale=float(2 ** 15), mean=fbank_mean, std2=2.0 * fbank_std)
