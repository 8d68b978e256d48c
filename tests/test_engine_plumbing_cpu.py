"""CPU: the engine's host-side plumbing on the main path (weight packing incl. the hyper-LoRA K-extension and SwiGLU row
interleave, splice planning, KV-cache bookkeeping, the decode-step call sequence) run through the CPU stand-in of the kernel
library (tests/fake_ops.py) and compared with the reference's golden outputs.  The kernels themselves are GPU-tested."""
import pytest
import torch

import fake_ops
from helpers import engine_cfg, load_golden, rel_l2


@pytest.mark.parametrize("name", ["llama_small", "llama_small_bs2"])
def test_engine_on_the_stand_in_library_reproduces_the_reference(monkeypatch, name):
    from crab_b200 import engine

    monkeypatch.setattr(engine, "ops", fake_ops)
    monkeypatch.setattr(fake_ops, "MIN_K", 8)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    g, case, sd, ocfg, ids, X = load_golden(name)
    eng = engine.CrabEngine(sd, engine_cfg(case, ocfg), torch.device("cpu"))
    emb = g["inputs_embeds"].to(torch.bfloat16)   # the reference's own inputs_embeds: decoder parity isolated from the encoders
    n_new = g["generated_ids"].shape[1]
    out, logits = eng.generate_from_embeds(emb.clone(), n_new, return_logits=True, teacher_tokens=g["generated_ids"])
    assert rel_l2(logits[0], g["prefill_last_logits"]) < 3e-2 and rel_l2(logits[1], g["step1_logits"]) < 3e-2
    assert torch.equal(out[:, 0], g["generated_ids"][:, 0])
    # encoders, bridges and the splice (CLIP, BEATs incl. the Toeplitz pos-conv and gated bias, Q-Formers, left padding)
    assert rel_l2(eng.encode_video(X[0]["<video>"]), g["vl_out"]) < 3e-2
    assert rel_l2(eng.encode_audio(X[0]["<audio>"]), g["al_out"]) < 3e-2
    e2, mask, pos = eng.prepare_inputs(ids, X)
    assert tuple(e2.shape) == tuple(g["inputs_embeds"].shape) and rel_l2(e2, g["inputs_embeds"]) < 3e-2
    assert torch.equal(mask, g["attention_mask"]) and torch.equal(pos, g["position_ids"])


def test_flag_ring_and_split_policy(monkeypatch):
    """Host-side bookkeeping of the fused decode linears: flag slots are handed round a ring (launch k raises slot k and zeroes slot
    k - 1; the first launch of a step zeroes the last slot of the previous one), every statistics-carrying launch of a step gets a
    distinct slot, and the decode step hands out exactly as many slots as it has such launches; short contexts skip the KV split."""
    from crab_b200 import engine

    monkeypatch.setattr(engine, "ops", fake_ops)
    monkeypatch.setattr(fake_ops, "MIN_K", 8)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    eng = engine.CrabEngine(sd, engine_cfg(case, ocfg), torch.device("cpu"))
    n = 7
    slot = eng._flag_ring(n)
    kws = [slot() for _ in range(n)]
    ring = eng._buf("dec_flag_ring_%d" % n, (n, 32), torch.int32, zero=True)
    for k, kw in enumerate(kws):
        assert kw["flags"].data_ptr() == ring[k].data_ptr() and kw["flags_clear"].data_ptr() == ring[(k - 1) % n].data_ptr()
        assert kw["stats_scratch"].numel() >= 8 * 36 * 32 and kw["stats_scratch"].dtype == torch.float32
    assert len({kw["flags"].data_ptr() for kw in kws}) == n
    with pytest.raises(StopIteration):
        slot()
    # the decode step asks for one slot per statistics-carrying launch: qkv + gate/up (+ o, + down with LoRA) per layer, + lm_head
    asked = []
    real = eng._flag_ring

    def spy(m):
        asked.append(m)
        return real(m)
    monkeypatch.setattr(eng, "_flag_ring", spy)
    emb = g["inputs_embeds"].to(torch.bfloat16)
    eng.generate_from_embeds(emb.clone(), 3)
    L = len(eng.layers)
    assert asked and all(m in (L * 3 + 1, L * 4 + 1, L * 2 + 1) for m in asked), asked
    # KV split policy: contexts of at most 256 keys over the whole request run unsplit (no combine launch)
    eng.prefill(emb.clone())
    eng.begin_decode(emb.shape[0], use_graph=False, max_len=200)
    assert eng._dec_args[1] == 1
    eng.begin_decode(emb.shape[0], use_graph=False, max_len=2000)
    assert eng._dec_args[1] >= 1
