"""Host-side logic that needs no GPU: the reference-facing surface (config dataclasses, tokenizer bookkeeping,
parameter naming) and the data-parallel sharding helpers (world_size-2 gloo)."""
import os
from pathlib import Path

import pytest
import torch

from helpers import GOLDEN, engine_cfg, load_golden

ROOT = Path(__file__).resolve().parent.parent


def test_config_dataclasses_parse_like_the_reference_cli():
    import transformers

    from crab_b200.configs.unified_config import DataArguments, InferenceArguments, ModelArguments

    p = transformers.HfArgumentParser([ModelArguments, DataArguments, InferenceArguments])
    m, d, i = p.parse_args_into_dataclasses(args=["--llm_name", "llama", "--avqa_task", "True", "--device", "cuda:0",
                                                  "--video_frame_nums", "10", "--vit_ckpt_path", "/x"])
    assert (m.llm_name, m.vit_ckpt_path, m.patch_size, m.visual_query_token_nums) == ("llama", "/x", 14, 32)
    assert d.avqa_task is True and d.video_frame_nums == 10 and i.device == "cuda:0"
    assert ModelArguments.select_layer_list == [14, 22, 23]
    assert "select_layer_list" not in {f.name for f in __import__("dataclasses").fields(ModelArguments)}


def test_manifest_names_match_the_reference_state_dict():
    """Every tensor name/shape the engine reads exists, with that shape, in the REAL reference's state dict
    (recorded in the golden fixture when it was generated from /root/reference)."""
    from crab_b200.models.unified_arch import full_manifest

    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    ref = {k[len("base_model.model."):]: tuple(v) for k, v in g["manifest"].items()}
    ours = full_manifest(engine_cfg(case, ocfg), d_model=case["d_model"])
    missing = [k for k in ours if k not in ref and "post_layernorm" not in k]
    assert not missing, missing[:5]
    bad = [k for k in ours if k in ref and tuple(ours[k]) != ref[k]]
    assert not bad, [(k, ours[k], ref[k]) for k in bad[:5]]


def test_tokenizer_bookkeeping_matches_reference_ids():
    from transformers import LlamaConfig

    from crab_b200.models.unified_arch import special_token_ids
    from crab_b200.models.unified_llama import UnifiedForCausalLM

    g = torch.load(GOLDEN / "llama_small.pt", weights_only=False)

    class Tok:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def add_tokens(self, t, special_tokens=False):
            self.n += len(t)
            return len(t)

    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=1, vocab_size=320)
    m = UnifiedForCausalLM(cfg)
    m.initialize_MM_tokenizer(Tok(320), mask_token_nums=6)
    assert m.SPECIAL_TOKEN_2_IDS == g["special_ids"] == special_token_ids(320)
    assert m.lm_head.weight.shape[0] == 337 and m.get_model().embed_tokens.weight.shape[0] == 337
    assert m.KEYS == ["<image>", "<video>", "<audio>"]
    names = dict(m.named_modules())
    for leaf in ("q_proj", "k_proj", "v_proj", "o_proj"):
        assert isinstance(names[f"model.layers.0.self_attn.{leaf}"], torch.nn.Linear)
    for leaf in ("gate_proj", "up_proj", "down_proj"):
        assert isinstance(names[f"model.layers.0.mlp.{leaf}"], torch.nn.Linear)
    with pytest.raises(RuntimeError):
        m.engine()  # no CUDA target selected: must not silently run anywhere else


def test_shard_range_properties():
    from crab_b200.parallel import shard_batch, shard_range

    for n in (0, 1, 7, 32, 255, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard_batch(list(range(10)), 1, 3) == [4, 5, 6]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    from crab_b200.parallel import gather_rows, shard_range

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 5  # ragged: shards of 3 and 2
    lo, hi = shard_range(n_total, rank, world)
    local = (torch.arange(lo, hi)[:, None] * 10 + torch.arange(4)[None, :]).long()  # "generated ids" of my samples
    full = gather_rows(local, n_total, rank, world)
    q.put((rank, full.tolist()))
    dist.destroy_process_group()


def test_all_gather_of_generated_ids_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    want = (torch.arange(5)[:, None] * 10 + torch.arange(4)[None, :]).tolist()
    assert res[0] == want and res[1] == want


def test_bench_roofline_constants_match_the_survey():
    """bench.py's algorithmic FLOP count per sample reproduces SURVEY.md 8(d): 15.86 TFLOP at S = 1086 for LLaMA-7B dims (the
    figure `phases.prefill_frac_of_tensor_roofline` divides by), and both backbones resolve to their checkpoint dims."""
    import types

    import bench

    a = types.SimpleNamespace(backbone="llama", layers=0)
    b = bench.backbone(a)
    assert (b["hidden"], b["inter"], b["layers"], b["heads"], b["kv_heads"], b["vocab"]) == (4096, 11008, 32, 32, 32, 32017)
    assert abs(bench.algorithmic_prefill_tflop(b, 1086) - 15.86) < 0.01
    q = bench.backbone(types.SimpleNamespace(backbone="qwen", layers=0))
    assert (q["hidden"], q["inter"], q["layers"], q["heads"], q["kv_heads"], q["vocab"]) == (3584, 18944, 28, 28, 4, 152081)
    assert q["qkv_bias"] and q["rope_theta"] == 1e6
    assert bench.backbone(types.SimpleNamespace(backbone="qwen", layers=3))["layers"] == 3
