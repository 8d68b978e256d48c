"""The UNMODIFIED reference entry script — /root/reference/scripts/quick_start.py::inference (:453-585) with its own
inference_ntp loop (:30-50) — executed against crab_b200 through the four shim modules INTEGRATION.md §1 tells a maintainer
to add, injected here via sys.modules:

    models.unified_llama   -> crab_b200.models.unified_llama        configs.unified_config -> crab_b200.configs.unified_config
    models.unified_arch    -> crab_b200.models.unified_arch         models.unified_qwen    -> crab_b200.models.unified_qwen

Everything else the script touches is the reference's own code (peft_hyper.get_peft_model, utils.avss_utils, its argument
parsing, DataLoader, its generate / batch_decode / print loop) except what needs files or packages that do not exist offline:
`dataset.quick_start_dataset` (jsonlines / librosa / decord + the AVQA files) is a stub that yields ONE synthetic sample through a
collator of the reference's shape (dataset/quick_start_dataset.py:624-707), `utils.util` re-exports this package's
`prepare_sample` (utils/util.py:33-47; the reference's file imports jsonlines), `LlamaTokenizer.from_pretrained` returns a
stand-in (no tokenizer files), checkpoints are synthetic.  The kernel library is replaced by its CPU stand-in (tests/fake_ops.py)
because this container has no GPU; tests/test_full_shape_gpu.py::test_mirror_model_on_gpu_full_width runs the same mirror
object on the B200.  Needs the reference checkout: runs in the build container only."""
import importlib.util
import sys
import tempfile
import types
from pathlib import Path

import pytest
import torch

import fake_ops
from oracle import ref_shims as R
from oracle import synth

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference checkout not present")
REF = Path("/root/reference")


class _Tok:
    """Stand-in for LlamaTokenizer: whitespace 'tokens' hashed into the vocabulary, the MM tokens get real ids via add_tokens."""

    pad_token_id, eos_token_id = None, 2

    def __init__(self, n):
        self.n, self.added = n, {}

    @classmethod
    def from_pretrained(cls, path, **kw):
        import json

        return cls(json.load(open(Path(path) / "config.json"))["vocab_size"])

    def __len__(self):
        return self.n

    def add_tokens(self, toks, special_tokens=False):
        for t in toks:
            self.added[t] = self.n
            self.n += 1
        return len(toks)

    def tokenize(self, text):
        return text.split()

    def convert_tokens_to_ids(self, toks):
        return [self.added.get(t, 3 + (sum(map(ord, t)) % 300)) for t in toks]

    def batch_decode(self, ids, skip_special_tokens=False):
        return [" ".join(str(int(t)) for t in row) for row in ids]


def test_unmodified_quick_start_inference(monkeypatch, capsys):
    import transformers
    from transformers import LlamaConfig

    from crab_b200 import engine
    from crab_b200.configs import unified_config as our_cfg
    from crab_b200.eval import driver
    from crab_b200.models import unified_arch, unified_llama, unified_qwen

    # ---- no GPU here: kernel library -> CPU stand-in, device plumbing -> cpu --------------------------------------------------
    monkeypatch.setattr(engine, "ops", fake_ops)
    monkeypatch.setattr(unified_arch, "ops", fake_ops, raising=False)
    monkeypatch.setattr(fake_ops, "MIN_K", 8)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "manual_seed", lambda s: None)
    monkeypatch.setattr(torch.cuda, "manual_seed_all", lambda s: None)
    built = {}

    def cpu_engine(self):
        if self._engine is None:
            self._engine = engine.CrabEngine(self.state_dict(), unified_arch.build_crab_config(self.decoder_config(), self, 1024),
                                             torch.device("cpu"))
            built["engine"] = self._engine
        return self._engine
    monkeypatch.setattr(unified_llama.UnifiedForCausalLM, "engine", cpu_engine)
    monkeypatch.setattr(unified_llama.UnifiedForCausalLM, "cuda", lambda self, device=None: self)     # .npu() lands here
    monkeypatch.setattr(unified_llama.UnifiedForCausalLM, "device", property(lambda self: torch.device("cpu")))
    R.install_shims()          # accelerate stubs etc. for the reference's own peft_hyper; puts /root/reference on sys.path

    # ---- the four INTEGRATION §1 shims + the offline stand-ins --------------------------------------------------------------
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
        return m
    models_pkg = mod("models", unified_llama=unified_llama, unified_arch=unified_arch, unified_qwen=unified_qwen)
    models_pkg.__path__ = []
    for n, m in (("models.unified_llama", unified_llama), ("models.unified_arch", unified_arch), ("models.unified_qwen", unified_qwen)):
        monkeypatch.setitem(sys.modules, n, m)
    cfg_pkg = mod("configs", unified_config=our_cfg)
    cfg_pkg.__path__ = []
    monkeypatch.setitem(sys.modules, "configs.unified_config", our_cfg)

    tmp = Path(tempfile.mkdtemp(prefix="crab_qs_script_"))
    g = torch.Generator().manual_seed(11)
    state = {}

    class _Dataset(list):
        pass

    def get_dataset_collator(data_args, tokenizer, image_processor=None, mode="test", test_name="test_s", **kw):
        """One AVQA-shaped sample: 2 frames, 2 one-second fbank segments, an instruction with one <video> and one <audio>."""
        assert data_args.avqa_task and mode == "test"
        inst = {"instruction": "w1 w2 <video> w3 w4 <audio> what is playing ?", "output": "piano", "task_name": "avqa",
                "video": torch.randn(2, 3, 224, 224, generator=g), "audio": 0.5 * torch.randn(2, 98, 128, generator=g),
                "video_path": "v.mp4", "audio_path": "a.wav"}
        state["tokenizer"] = tokenizer

        def collate(instances):   # the reference collator's shape (dataset/quick_start_dataset.py:624-707)
            out = {"batch_input_ids": [], "batch_labels": [], "batch_X_modals": [], "batch_metadata": [], "batch_task_names": []}
            for it in instances:
                ids = tokenizer.convert_tokens_to_ids(tokenizer.tokenize(it["instruction"]))
                out["batch_input_ids"].append(torch.tensor(ids, dtype=torch.long))
                out["batch_labels"].append(torch.tensor([-100] * len(ids), dtype=torch.long))
                out["batch_X_modals"].append({"<video>": it["video"], "<audio>": it["audio"]})
                out["batch_metadata"].append({"instruction": it["instruction"], "output": it["output"]})
                out["batch_task_names"].append(it["task_name"])
            return out
        return _Dataset([inst]), collate
    ds_pkg = mod("dataset", quick_start_dataset=None)
    ds_pkg.__path__ = []
    ds_pkg.quick_start_dataset = mod("dataset.quick_start_dataset", get_dataset_collator=get_dataset_collator, get_v2_pallete=lambda **k: None)
    utils_pkg = mod("utils")
    utils_pkg.__path__ = [str(REF / "utils")]          # utils.avss_utils is the reference's own file
    utils_pkg.util = mod("utils.util", set_seed=lambda seed=42: torch.manual_seed(seed), find_all_linear_names=lambda m: [],
                         prepare_sample=lambda data: driver.prepare_sample(data, device="cpu"), write2json=driver.write_jsonl,
                         load_ckpt=lambda p: torch.load(p, map_location="cpu"))
    utils_pkg.deepspeed_utils = mod("utils.deepspeed_utils")
    # no tokenizer files offline: LlamaTokenizer.from_pretrained (quick_start.py:495-500) hands out the stand-in
    monkeypatch.setattr(transformers.LlamaTokenizer, "from_pretrained", classmethod(lambda cls, path, **kw: _Tok.from_pretrained(path)))
    # transformers 5.5's TrainingArguments.__post_init__ sets up accelerate's PartialState; accelerate is not installed offline and
    # the script only reads plain fields (fp16 / bf16 / *_branch / loss weights) from it
    monkeypatch.setattr(transformers.TrainingArguments, "__post_init__", lambda self: None)

    # ---- synthetic checkpoints: d_model is hard-wired to 4096 in the script (:457), everything else is kept small ------------
    llama_cfg = dict(hidden_size=4096, intermediate_size=256, num_hidden_layers=1, num_attention_heads=32, num_key_value_heads=32,
                     vocab_size=320, max_position_embeddings=2048, rms_norm_eps=1e-6, rope_theta=10000.0, eos_token_id=2, pad_token_id=0)
    llama_dir = tmp / "llama"
    LlamaConfig(**llama_cfg).save_pretrained(llama_dir)
    clip_dir = R.make_clip_dir(tmp, image=224, patch=14, hidden=1024, inter=256, layers=24, heads=16)
    beats_pt = R.make_beats_ckpt(tmp, dict(R.BEATS_CFG_PUBLIC, encoder_layers=2, encoder_ffn_embed_dim=256))
    ckpt_dir = tmp / "ckpt"
    ckpt_dir.mkdir()
    # finetune_weights.bin with the reference's key names (projectors + hyper-LoRA), as quick_start.py:537-542 loads it
    dcfg = engine.DecoderConfig(hidden=4096, inter=256, layers=1, heads=32, kv_heads=32, head_dim=128, vocab=337)
    man = {"base_model.model." + k: v for k, v in unified_arch.decoder_manifest(dcfg).items() if "lora_" in k}
    q = engine.QformerConfig()
    for kind, width in (("visual", 1024), ("audio", 768)):
        pre = "base_model.model.model." + ("vl_projector." if kind == "visual" else "al_projector.")
        man.update({pre + k: v for k, v in unified_arch.projector_manifest(kind, q, width, 4096, 32).items()})
    torch.save(synth.synth_state_dict(man, 5), ckpt_dir / "finetune_weights.bin")

    monkeypatch.setattr(sys, "argv", ["quick_start.py", "--model_name_or_path", str(llama_dir), "--vit_ckpt_path", str(clip_dir),
                                      "--BEATs_ckpt_path", str(beats_pt), "--ckpt_dir", str(ckpt_dir), "--avqa_task", "True",
                                      "--visual_branch", "True", "--audio_branch", "True", "--device", "cpu", "--output_dir", str(tmp / "out"),
                                      "--bf16", "False"])
    # ---- load and run the script's own source, untouched -----------------------------------------------------------------------
    spec = importlib.util.spec_from_file_location("ref_quick_start", REF / "scripts" / "quick_start.py")
    qs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qs)
    qs.inference()
    out = capsys.readouterr().out
    eng = built["engine"]
    assert eng.lora and eng.has_encoders and eng.cfg.decoder.hidden == 4096
    assert "load ckpt from" in out and "'predict':" in out
    # the script's generate(**sample, use_cache=True, max_new_tokens=500) ran with the checkpoint's EOS default (ids end at EOS or
    # at 500 tokens) and its tokenizer.batch_decode printed one record
    rec = out[out.index("'predict':"):]
    pred = rec.split("'")[3]
    n_tok = len(pred.split())
    assert 1 <= n_tok <= 500, rec[:400]
    if n_tok < 500:
        assert pred.split()[-1] == "2", "generation stopped early only at the EOS id of the checkpoint's config"
    tok = state["tokenizer"]
    assert len(tok) == 320 + 17 and eng.cfg.special_ids["<video>"] == 323
