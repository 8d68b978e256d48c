"""TEST INFRASTRUCTURE — generate tests/golden/frontend.pt from the REAL reference front-end in this container.

    python -m oracle.make_frontend_golden

Audio: imports `/root/reference/dataset/audio_processor.py` (its module-level `import librosa` is satisfied by an empty
stub: `preprocess` never touches it) and runs the reference's own `preprocess` on seeded waveforms.
Video: the reference calls the HF `CLIPImageProcessor` of the vision tower on PIL frames decoded at 224x224
(dataset/quick_start_dataset.py:303-315); the installed transformers' CLIPImageProcessor (clip-vit-large-patch14 settings)
is run on seeded uint8 frames.
"""
from __future__ import annotations

import importlib.util
import sys
import types

sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def main():
    stub = "librosa" not in sys.modules
    if stub:
        sys.modules["librosa"] = types.ModuleType("librosa")
    spec = importlib.util.spec_from_file_location("ref_audio_processor", REF / "dataset" / "audio_processor.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if stub:
        del sys.modules["librosa"]  # transformers probes optional packages through importlib and trips over a spec-less stub
    from oracle.frontend_oracle import synth_frames, synth_waveforms
    wav = synth_waveforms(11)
    fb = mod.preprocess(wav)                       # the reference's own function
    wav2 = wav[:2, :12345].contiguous()            # ragged length (frames = 1 + (L - 400) // 160 = 75)
    fb2 = mod.preprocess(wav2)

    from PIL import Image
    from transformers import CLIPImageProcessor
    proc = CLIPImageProcessor(do_resize=True, size={"shortest_edge": 224}, resample=3, do_center_crop=True,
                              crop_size={"height": 224, "width": 224}, do_rescale=True, rescale_factor=1 / 255,
                              do_normalize=True, image_mean=[0.48145466, 0.4578275, 0.40821073],
                              image_std=[0.26862954, 0.26130258, 0.27577711], do_convert_rgb=True)
    frames = synth_frames(5, 1)
    pil = [Image.fromarray(f.numpy()) for f in frames]
    pv = proc.preprocess(pil, return_tensors="pt")["pixel_values"].to(torch.float32)
    # non-224 input: Pillow itself (what transformers 4.37's CLIPImageProcessor.resize calls), shortest edge 224 + centre crop
    import numpy as np
    big = synth_frames(9, 1, 8)[0].numpy().repeat(45, 0).repeat(61, 1)[:300, :400]   # 300 x 400 blocky image with hard edges
    oh, ow = 224, int(224 * 400 / 300)
    rz = np.asarray(Image.fromarray(big).resize((ow, oh), Image.BICUBIC))
    top, left = (oh - 224) // 2, (ow - 224) // 2
    resize_crop = torch.from_numpy(rz[top:top + 224, left:left + 224].copy())
    import torchaudio.compliance.kaldi as K
    mel_banks, _ = K.get_mel_banks(128, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    window = K._feature_window_function(K.POVEY, 400, 0.42, torch.device('cpu'), torch.float32)
    out = dict(resize_seed=9, resize_crop_300x400=resize_crop, mel_banks=mel_banks.clone(), window=window.clone(), wave_seed=11, fbank=fb.clone(), fbank_ragged=fb2.clone(), ragged_len=int(wav2.shape[1]),
               frame_seed=5, pixel_values=pv.clone(),
               versions=dict(torch=str(torch.__version__), torchaudio=str(__import__("torchaudio").__version__),
                             transformers=str(__import__("transformers").__version__)))
    torch.save(out, ROOT / "tests" / "golden" / "frontend.pt")
    print({k: (tuple(v.shape) if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
