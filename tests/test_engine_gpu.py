"""End-to-end parity of the CUDA path (crab_b200.engine, through the C ABI) against (a) the golden outputs of the
real reference and (b) the CPU oracle, stage by stage, on the small golden cases.

Tolerance contract: the GPU path stores bf16 and accumulates fp32; against the fp32 reference the
expected error is a few bf16 ulps per op, compounding with depth.  Each stage asserts a relative L2 bound of 2x the value
measured on B200 (printed with pytest -s; measured: CLIP tap 5.8e-3, VL 8.8e-3, BEATs 7.1e-3, AL 8.3e-3, inputs_embeds
7e-3, logits 6.6e-3 / 7.0e-3).  Token ids: teacher-forced arg-max must agree wherever the reference's top-2 logit margin
exceeds 4x the measured logit error; free-running ids must equal the reference's up to the first non-decisive step.
The same checks at BASELINE's full shapes live in tests/test_full_shape_gpu.py.
"""
import pytest
import torch

from helpers import GOLDEN, engine_cfg, load_golden, rel_l2
from oracle import crab_oracle as O

pytestmark = pytest.mark.gpu
CASES = sorted(p.stem for p in GOLDEN.glob("llama_*.pt"))


@pytest.fixture(scope="module", params=CASES)
def setup(request, cuda_dev):
    from crab_b200.engine import CrabEngine

    g, case, sd, ocfg, ids, X = load_golden(request.param)
    eng = CrabEngine(sd, engine_cfg(case, ocfg), cuda_dev)
    return g, case, sd, ocfg, ids, X, eng


def test_encoders_vs_reference(setup, cuda_dev):
    g, case, sd, ocfg, ids, X, eng = setup
    v = X[0]["<video>"].to(cuda_dev)
    tokens = ocfg.image_tokens + 1
    clip = eng.clip_forward(v).view(v.shape[0], tokens, -1)[:, 1:].reshape(-1, ocfg.clip.hidden)
    e = rel_l2(clip, g["vit_taps"][-1])
    print(f"clip last tap rel_l2={e:.3e}")
    assert e < 1.2e-2
    vl = eng.encode_video(v)
    e = rel_l2(vl, g["vl_out"])
    print(f"vl_projector rel_l2={e:.3e}")
    assert e < 1.8e-2
    a = X[0]["<audio>"].to(cuda_dev)
    b, T = eng.beats_forward(a)
    e = rel_l2(b.view(a.shape[0], T, -1), g["beats_out"])
    print(f"beats rel_l2={e:.3e}")
    assert e < 1.5e-2
    al = eng.encode_audio(a)
    e = rel_l2(al, g["al_out"])
    print(f"al_projector rel_l2={e:.3e}")
    assert e < 1.7e-2


def test_prepare_inputs_vs_reference(setup):
    g, case, sd, ocfg, ids, X, eng = setup
    emb, mask, pos = eng.prepare_inputs(ids, X)
    assert tuple(emb.shape) == tuple(g["inputs_embeds"].shape)
    assert torch.equal(mask, g["attention_mask"]) and torch.equal(pos, g["position_ids"])
    e = rel_l2(emb, g["inputs_embeds"])
    print(f"inputs_embeds rel_l2={e:.3e}")
    assert e < 1.5e-2
    # text rows are exact bf16 roundings of the embedding table (pure gather)
    tab = sd["model.embed_tokens.weight"].to(torch.bfloat16)
    assert torch.equal(emb[0, -3:].cpu(), tab[ids[0][-3:]])


def test_prefill_and_decode_vs_reference(setup, cuda_dev):
    g, case, sd, ocfg, ids, X, eng = setup
    # feed the REFERENCE's inputs_embeds so decoder parity is isolated from encoder error
    emb = g["inputs_embeds"].to(cuda_dev).to(torch.bfloat16)
    ref_ids = g["generated_ids"]
    n_new = ref_ids.shape[1]
    with torch.no_grad():
        _, ref_logits = O.greedy_generate(sd, g["inputs_embeds"], ocfg.decoder, n_new, teacher_tokens=ref_ids)
    out, logits = eng.generate_from_embeds(emb.clone(), n_new, use_graph=True, return_logits=True,
                                           teacher_tokens=ref_ids.to(cuda_dev))
    logits = logits.cpu()
    e0 = rel_l2(logits[0], g["prefill_last_logits"])
    e1 = rel_l2(logits[1], g["step1_logits"])
    print(f"prefill logits rel_l2={e0:.3e}; step-1 logits rel_l2={e1:.3e}")
    assert e0 < 1.4e-2 and e1 < 1.4e-2
    err = (logits - ref_logits).abs().max().item()
    top2 = ref_logits.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]  # (n, b)
    decisive = margin > 4 * err
    agree = logits.argmax(-1) == ref_logits.argmax(-1)
    print(f"max |dlogit|={err:.3e}; decisive steps {int(decisive.sum())}/{decisive.numel()}; agree {int(agree.sum())}")
    assert bool(agree[decisive].all())
    # no-graph path must give bit-identical results to the graph path
    out2, logits2 = eng.generate_from_embeds(emb.clone(), n_new, use_graph=False, return_logits=True,
                                             teacher_tokens=ref_ids.to(cuda_dev))
    assert torch.equal(out, out2) and torch.equal(logits2.cpu(), logits)


def test_generate_end_to_end(setup):
    g, case, sd, ocfg, ids, X, eng = setup
    n_new = g["generated_ids"].shape[1]
    ref_ids = g["generated_ids"]
    out = eng.generate(ids, X, n_new).cpu()
    match = (out == ref_ids).float().mean().item()
    print(f"free-running greedy ids: {out.tolist()} vs reference {ref_ids.tolist()} (match {match:.2f})")
    assert out.shape == ref_ids.shape
    # every row must reproduce the reference's ids up to its first NON-decisive step: a step is decisive when the fp32
    # oracle's top-2 margin (teacher-forced on the reference's ids) exceeds 4x the max |dlogit| measured for this engine
    with torch.no_grad():
        _, ref_logits = O.greedy_generate(sd, g["inputs_embeds"], ocfg.decoder, n_new, teacher_tokens=ref_ids)
    _, logits = eng.generate_from_embeds(g["inputs_embeds"].to(eng.dev).to(torch.bfloat16), n_new, return_logits=True,
                                         teacher_tokens=ref_ids.to(eng.dev))
    err = (logits.cpu() - ref_logits).abs().max().item()
    top2 = ref_logits.topk(2, dim=-1).values
    decisive = ((top2[..., 0] - top2[..., 1]) > 4 * err).t()          # (b, n)
    for b in range(out.shape[0]):
        bad = (~decisive[b]).nonzero()
        k = int(bad[0]) if bad.numel() else n_new
        assert torch.equal(out[b, :k], ref_ids[b, :k]), (b, k, out[b].tolist(), ref_ids[b].tolist())


def test_qwen_decoder_vs_reference(cuda_dev):
    """Qwen2 backbone through the same kernels: qkv bias in the GEMM epilogue, GQA in flash / decode attention, rope
    theta 1e6; compared with the reference's unified_qwen + peft_hyper outputs (tests/golden/qwen_small.pt)."""
    from crab_b200 import engine as E
    from oracle import synth

    g = torch.load(GOLDEN / "qwen_small.pt", weights_only=False)
    case = g["case"]
    sd = O.strip_peft_prefix(synth.synth_state_dict(g["manifest"], case["weight_seed"]))
    lc = case["llama_cfg"]
    hd = lc["hidden_size"] // lc["num_attention_heads"]
    dec_o = O.DecoderCfg(hidden=lc["hidden_size"], inter=lc["intermediate_size"], layers=lc["num_hidden_layers"],
                         heads=lc["num_attention_heads"], kv_heads=lc["num_key_value_heads"], head_dim=hd,
                         vocab=lc["vocab_size"], rope_theta=lc["rope_theta"], eps=lc["rms_norm_eps"], qkv_bias=True)
    cfg = E.CrabConfig(decoder=E.DecoderConfig(hidden=dec_o.hidden, inter=dec_o.inter, layers=dec_o.layers, heads=dec_o.heads,
                                               kv_heads=dec_o.kv_heads, head_dim=hd, vocab=dec_o.vocab,
                                               rope_theta=dec_o.rope_theta, eps=dec_o.eps, qkv_bias=True), max_ctx=256)
    eng = E.CrabEngine(sd, cfg, cuda_dev, load_encoders=False)
    ref_ids = g["generated_ids"]
    n_new = ref_ids.shape[1]
    with torch.no_grad():
        _, ref_logits = O.greedy_generate(sd, g["inputs_embeds"], dec_o, n_new, teacher_tokens=ref_ids)
    out, logits = eng.generate_from_embeds(g["inputs_embeds"].to(cuda_dev).to(torch.bfloat16), n_new, return_logits=True,
                                           teacher_tokens=ref_ids.to(cuda_dev))
    logits = logits.cpu()
    e0, e1 = rel_l2(logits[0], g["prefill_last_logits"]), rel_l2(logits[1], g["step1_logits"])
    print(f"qwen prefill logits rel_l2={e0:.3e}; step-1 rel_l2={e1:.3e}")
    assert e0 < 1.5e-2 and e1 < 1.5e-2   # measured 7e-3
    err = (logits - ref_logits).abs().max().item()
    top2 = ref_logits.topk(2, dim=-1).values
    decisive = (top2[..., 0] - top2[..., 1]) > 4 * err
    agree = logits.argmax(-1) == ref_logits.argmax(-1)
    assert bool(agree[decisive].all())
    free = eng.generate_from_embeds(g["inputs_embeds"].to(cuda_dev).to(torch.bfloat16), n_new).cpu()
    print(f"qwen free-running match {(free == ref_ids).float().mean().item():.2f}")


def test_full_width_decoder_layer_vs_oracle(cuda_dev):
    """7B-dim layers (D=4096, F=11008, 32 heads x 128) at the bench's tile shapes, 2 layers, bs 3, S=200: prefill and two
    decode steps against the CPU oracle on the same seeded weights."""
    from crab_b200 import engine as E
    from crab_b200.models.unified_arch import decoder_manifest
    from oracle import synth

    dcfg = E.DecoderConfig(layers=2, vocab=2048)
    sd = synth.synth_state_dict(decoder_manifest(dcfg), 3)
    dec_o = O.DecoderCfg(layers=2, vocab=2048)
    g = torch.Generator(device="cpu").manual_seed(9)
    emb = torch.randn(3, 200, 4096, generator=g)
    eng = E.CrabEngine(sd, E.CrabConfig(decoder=dcfg, max_ctx=256), cuda_dev, load_encoders=False)
    with torch.no_grad():
        ref_ids, ref_logits = O.greedy_generate(sd, emb.to(torch.bfloat16).float(), dec_o, 3)
    out, logits = eng.generate_from_embeds(emb.to(cuda_dev).to(torch.bfloat16), 3, return_logits=True,
                                           teacher_tokens=ref_ids.to(cuda_dev))
    e = [rel_l2(logits[i], ref_logits[i]) for i in range(3)]
    print("full-width logits rel_l2 per step:", ["%.3e" % v for v in e])
    assert max(e) < 2e-2


def test_full_width_qwen7b_layer_vs_oracle(cuda_dev):
    """BASELINE configs[4] shapes: Qwen2-7B-dim layers (D=3584, F=18944, 28 query / 4 kv heads x 128, qkv bias, theta 1e6)
    with hyper-LoRA, 1 layer, bs 2, S=160: prefill (CTA-pair / tcgen05 flash with GQA) and two decode steps (cluster row
    kernels on 18944-column rows, fused GQA decode attention, split-K streaming GEMMs) against the CPU oracle."""
    from crab_b200 import engine as E
    from crab_b200.models.unified_arch import decoder_manifest
    from oracle import synth

    dcfg = E.DecoderConfig(hidden=3584, inter=18944, layers=1, heads=28, kv_heads=4, head_dim=128, vocab=2048,
                           rope_theta=1e6, qkv_bias=True)
    sd = synth.synth_state_dict(decoder_manifest(dcfg), 5)
    dec_o = O.DecoderCfg(hidden=3584, inter=18944, layers=1, heads=28, kv_heads=4, head_dim=128, vocab=2048, rope_theta=1e6,
                         qkv_bias=True)
    g = torch.Generator(device="cpu").manual_seed(10)
    emb = torch.randn(2, 160, 3584, generator=g)
    eng = E.CrabEngine(sd, E.CrabConfig(decoder=dcfg, max_ctx=192), cuda_dev, load_encoders=False)
    with torch.no_grad():
        ref_ids, ref_logits = O.greedy_generate(sd, emb.to(torch.bfloat16).float(), dec_o, 3)
    out, logits = eng.generate_from_embeds(emb.to(cuda_dev).to(torch.bfloat16), 3, return_logits=True,
                                           teacher_tokens=ref_ids.to(cuda_dev))
    e = [rel_l2(logits[i], ref_logits[i]) for i in range(3)]
    print("qwen-7B-width logits rel_l2 per step:", ["%.3e" % v for v in e])
    assert max(e) < 2e-2


def test_batched_generate_with_eos_and_unequal_prompts(cuda_dev):
    """The batched-evaluation call pattern (scripts/finetune/inference_hyper_lora.py: generate(**sample,
    max_new_tokens=500) on a batch with unequal prompt lengths): rows stop at EOS, finished rows emit the pad id, decoding ends
    right after the last row finishes (HF semantics), and tokens before EOS equal the unconstrained greedy run."""
    from crab_b200.engine import CrabEngine
    from crab_b200.models.unified_llama import UnifiedConfig, UnifiedForCausalLM

    g, case, sd, ocfg, ids, X = load_golden("llama_small_bs2")
    assert len({int(t.numel()) for t in ids}) > 1, "the golden case must have unequal prompt lengths"
    eng = CrabEngine(sd, engine_cfg(case, ocfg), cuda_dev)
    lc = case["llama_cfg"]
    model = UnifiedForCausalLM.from_engine(UnifiedConfig(hidden_size=lc["hidden_size"], intermediate_size=lc["intermediate_size"],
                                                         num_hidden_layers=lc["num_hidden_layers"],
                                                         num_attention_heads=lc["num_attention_heads"],
                                                         num_key_value_heads=lc["num_key_value_heads"],
                                                         vocab_size=lc["vocab_size"] + 17), eng)
    n = 24
    free = model.generate(batch_input_ids=ids, batch_labels=None, batch_X_modals=X, batch_task_names=["avqa"] * len(ids),
                          max_new_tokens=n).cpu()
    assert tuple(free.shape) == (len(ids), n)
    # choose EOS ids so that row 0 stops at step 2 and row 1 at step 5
    eos = [int(free[0, 2]), int(free[1, 5])]
    pad = int(model.model.pad_token_id or 0)
    exp = free.clone()
    stop = []
    for r in range(exp.shape[0]):
        hit = [i for i in range(n) if int(exp[r, i]) in eos]
        k = hit[0] if hit else n - 1
        exp[r, k + 1:] = pad
        stop.append(k)
    exp = exp[:, : max(stop) + 1]
    out = model.generate(batch_input_ids=ids, batch_labels=None, batch_X_modals=X, batch_task_names=["avqa"] * len(ids),
                         max_new_tokens=n, eos_token_id=eos).cpu()
    assert torch.equal(out, exp), (out.tolist(), exp.tolist())


def test_call_graphs_replay_equals_eager(cuda_dev):
    """The launch-bound call sites outside the decode loop (small encoder batches, short prefills) run as one CUDA-graph replay
    per shape: with inputs that CHANGE between calls the replays must give the eager path's bits (static input buffers refreshed,
    ViT taps handed over, KV cache and position state as after an eager prefill)."""
    from crab_b200.engine import CrabEngine

    g, case, sd, ocfg, ids, X = load_golden("llama_small")
    eng = CrabEngine(sd, engine_cfg(case, ocfg), cuda_dev)
    assert eng.call_graphs
    gen = torch.Generator().manual_seed(9)
    X2 = [{k: v + 0.25 * torch.randn(v.shape, generator=gen) for k, v in x.items()} for x in X]
    outs = {}
    for mode in ("graph", "graph_again", "eager"):
        eng.call_graphs = mode != "eager"
        res = []
        for Xi in (X, X2, X):
            emb, _, _ = eng.prepare_inputs(ids, [{k: v.to(cuda_dev) for k, v in x.items()} for x in Xi])
            logits, nxt = eng.prefill(emb.clone())
            eng.begin_decode(emb.shape[0])
            step_logits, _ = eng.decode_step()
            res.append((emb.clone(), logits.clone(), step_logits.clone(), eng.cur_len))
        outs[mode] = res
    assert any(k[0] == "prefill" for k in eng._call_graphs) and any(k[0] in ("video", "audio") for k in eng._call_graphs)
    for mode in ("graph", "graph_again"):
        for a, b in zip(outs[mode], outs["eager"]):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and a[3] == b[3]
    assert not torch.equal(outs["eager"][0][1], outs["eager"][1][1])   # the second input really differs
