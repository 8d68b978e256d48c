#!/usr/bin/env python
"""bench.py — Crab AV-prompt hot path on B200: prefill + greedy-decode tokens/s (BASELINE.json metric).

A "step" is ONE pass of the hot path over ONE batch of synthetic AVQA-shaped input:
    CLIP ViT-L/14 over 8x224^2 frames + BEATs over 10x1 s fbank segments + both Q-Former bridges
    -> splice into a 512-token prompt (S = 1086) -> LLaMA-2-7B-dim decoder prefill (hyper-LoRA on all 7 linears)
    -> 128 greedy tokens through a CUDA-graph decode step.
Workload at N=1: BASELINE.json configs[2] "bs32 synthetic AVQA (10s audio, 8 frames, 512-tok prompt) bf16 on 1xB200";
at N>1 every rank runs its own 32 samples (weak scaling: configs[3] at N=8) and the generated ids are all-gathered.

    python bench.py --gpus 1 --steps 3 --warmup 3                    # our arm
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0  # reference arm: CPU oracle port, bounded sample
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # one rank per GPU

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_traffic():
    """dram__bytes_read+write per launch from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {k: float(d[k]) for k in FALLBACK_PEAKS}, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# synthetic weights, generated tensor-by-tensor on the device from (name, seed): random init of the real architecture
# ------------------------------------------------------------------------------------------------------------------
class LazySynthSD:
    """dict-like state dict: tensors are produced on demand on `device` so 7B parameters never sit on the host."""

    def __init__(self, manifest, seed, device):
        self.m, self.seed, self.dev = manifest, seed, device

    def __contains__(self, k):
        return k in self.m

    def items(self):
        for k in self.m:
            yield k, self[k]

    def get(self, k, default=None):
        return self[k] if k in self.m else default

    def __getitem__(self, key):
        shape = tuple(self.m[key])
        h = int.from_bytes(hashlib.sha256(f"{self.seed}:{key}".encode()).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
        g = torch.Generator(device=self.dev).manual_seed(h)
        r = torch.randn(shape, generator=g, device=self.dev, dtype=torch.float32)
        last = key.rsplit(".", 1)[-1]
        n = math.prod(shape)
        if "lora_B" in key:
            return 0.05 * r
        if "lora_A" in key or "lora_route" in key:
            return r / math.sqrt(shape[-1])
        if last == "weight_g":
            return 0.5 + 0.1 * r.abs()
        if last == "grep_a":
            return 1.0 + 0.2 * r
        if "relative_attention_bias" in key or "query_tokens" in key or "class_embedding" in key or "position_embedding" in key:
            return 0.5 * r
        if "embed_tokens" in key:
            return r
        if len(shape) == 1:
            return 0.05 * r if last == "bias" else 1.0 + 0.1 * r
        return r / math.sqrt(n // shape[0])


def make_inputs(bs, frames, audio_segs, audio_len, prompt_len, base_vocab, ids_map, rank):
    """Pinned host tensors, per-sample seeds 1000+i."""
    ids, X = [], []
    for i in range(bs):
        g = torch.Generator(device="cpu").manual_seed(1000 + i + 100000 * rank)
        video = torch.randn(frames, 3, 224, 224, generator=g).pin_memory()
        audio = (0.5 * torch.randn(audio_segs, audio_len, 128, generator=g)).pin_memory()
        t = torch.randint(3, base_vocab, (prompt_len,), generator=g)
        t[10] = ids_map["<video>"]
        t[20] = ids_map["<audio>"]
        ids.append(t)
        X.append({"<video>": video, "<audio>": audio})
    return ids, X


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 7 and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# Decoder backbones of the two benchmark configurations (BASELINE.json configs[2] / configs[4])
BACKBONES = {
    "llama": dict(hidden=4096, inter=11008, layers=32, heads=32, kv_heads=32, head_dim=128, base_vocab=32000, rope_theta=1e4,
                  qkv_bias=False),   # LLaMA-2-7B-chat dims (models/unified_llama.py)
    "qwen": dict(hidden=3584, inter=18944, layers=28, heads=28, kv_heads=4, head_dim=128, base_vocab=152064, rope_theta=1e6,
                 qkv_bias=True),     # Qwen2-7B dims (models/unified_qwen.py)
}


def backbone(args):
    b = dict(BACKBONES[args.backbone])
    if args.layers:
        b["layers"] = args.layers
    b["vocab"] = b["base_vocab"] + 17  # initialize_MM_tokenizer adds 17 tokens (models/unified_arch.py:409-459)
    return b


def algorithmic_prefill_tflop(b, S):
    """Minimum FLOPs per sample up to the first token: encoders + bridges, decoder linears 2*params*S, causal
    attention 2*S^2*nq per layer, hyper-LoRA (11*in + 24*out MACs per token and linear), last-position lm_head."""
    nq, nk, D, F, L = b["heads"] * b["head_dim"], b["kv_heads"] * b["head_dim"], b["hidden"], b["inter"], b["layers"]
    lin = D * (nq + 2 * nk) + nq * D + 3 * D * F
    lora = 11 * (3 * D + nq + 2 * D + F) + 24 * ((nq + 2 * nk) + D + 2 * F + D)
    dec = 2.0 * S * L * (lin + lora) + 2.0 * S * S * nq * L + 2.0 * D * b["vocab"]
    return 1.387 + dec / 1e12


# CPU arm: the oracle port on the host cores, on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------------
def cpu_port_run(args, threads):
    """Times oracle/crab_oracle.py (the CPU restatement of the reference's path; the reference itself is Python +
    HF and cannot travel to the GPU box).  Bounded sample: ONE sample of the batch, `cpu_layers` of each stack, and
    `cpu_steps` decode steps, scaled to the full depth / 128 steps (every layer of a stack costs the same).
    Returns (tokens_per_s, sample_description, detail)."""
    from oracle import crab_oracle as O
    from oracle import synth

    torch.set_num_threads(threads)
    Ls, steps = args.cpu_layers, args.cpu_steps
    b = backbone(args)
    dec = O.DecoderCfg(hidden=b["hidden"], inter=b["inter"], layers=Ls, heads=b["heads"], kv_heads=b["kv_heads"],
                       head_dim=b["head_dim"], vocab=b["vocab"], rope_theta=b["rope_theta"], qkv_bias=b["qkv_bias"])
    cfg = O.CrabCfg(decoder=dec, clip=O.ClipCfg(layers=Ls), beats=O.BeatsCfg(layers=Ls), qformer=O.QformerCfg(),
                    select_layers=(Ls,), image_tokens=256, base_vocab=b["base_vocab"])
    from crab_b200.engine import BeatsConfig, ClipConfig, CrabConfig, DecoderConfig, QformerConfig
    from crab_b200.models.unified_arch import full_manifest

    ecfg = CrabConfig(decoder=DecoderConfig(hidden=b["hidden"], inter=b["inter"], layers=Ls, heads=b["heads"],
                                            kv_heads=b["kv_heads"], head_dim=b["head_dim"], vocab=b["vocab"],
                                            rope_theta=b["rope_theta"], qkv_bias=b["qkv_bias"]),
                      clip=ClipConfig(layers=Ls), beats=BeatsConfig(layers=Ls), qformer=QformerConfig())
    man = full_manifest(ecfg)
    sd = synth.synth_state_dict(man, 42)
    video, audio, ids = synth.synth_inputs(1000, frames=8, image=224, audio_segs=10, audio_len=98,
                                           prompt_len=args.prompt_len, base_vocab=b["base_vocab"],
                                           video_id=cfg.special_ids["<video>"], audio_id=cfg.special_ids["<audio>"])
    with torch.no_grad():
        t0 = time.perf_counter()
        taps = O.visual_encoder(sd, video.unsqueeze(0), cfg.clip, cfg.select_layers)
        t_clip = time.perf_counter() - t0
        t0 = time.perf_counter()
        vl = O.vl_projector(sd, taps[-1], cfg.qformer, 256)[0]
        t_vl = time.perf_counter() - t0
        t0 = time.perf_counter()
        be = O.audio_encoder(sd, audio.unsqueeze(0), cfg.beats)
        t_beats = time.perf_counter() - t0
        t0 = time.perf_counter()
        al = O.al_projector(sd, be, cfg.qformer)[0]
        t_al = time.perf_counter() - t0
        emb = sd["model.embed_tokens.weight"]
        x = torch.cat([emb[ids[:10]], vl, emb[ids[11:20]], al, emb[ids[21:]]], 0).unsqueeze(0)
        S = x.shape[1]
        t0 = time.perf_counter()
        h, cache = O.decoder_forward(sd, x, dec)
        t_pf = time.perf_counter() - t0
        t0 = time.perf_counter()
        logits = O.lm_head(sd, h[:, -1])
        t_head = time.perf_counter() - t0
        nxt = logits.argmax(-1)
        t0 = time.perf_counter()
        for _ in range(steps):
            h, cache = O.decoder_forward(sd, emb[nxt].unsqueeze(1), dec, cache)
            nxt = O.lm_head(sd, h[:, -1]).argmax(-1)
        t_dec = (time.perf_counter() - t0) / steps
    # scale the sampled depth to the full stacks: CLIP 23 layers, BEATs 12, decoder 32 (+ lm_head once per step)
    t_dec_layers = max(t_dec - t_head, 0.0)
    t_prefill = t_clip * 23 / Ls + t_vl + t_beats * 12 / Ls + t_al + t_pf * b["layers"] / Ls + t_head
    t_step = t_dec_layers * b["layers"] / Ls + t_head
    t_total = t_prefill + 127 * t_step
    tok_s = (S + args.new_tokens) / t_total
    detail = {"prefill_tok_s": S / t_prefill, "decode_tok_s": 1.0 / t_step, "t_prefill_s": t_prefill, "t_decode_step_s": t_step,
              "S": S, "measured": {"clip_s": t_clip, "vl_s": t_vl, "beats_s": t_beats, "al_s": t_al, "prefill_s": t_pf,
                                   "lm_head_s": t_head, "decode_step_s": t_dec}}
    sample = (f"1 of {args.bs} samples (fp32, {threads} threads): {Ls} of 23 CLIP / {Ls} of 12 BEATs / {Ls} of {b['layers']} decoder "
              f"layers at full width, S={S}, {steps} decode steps; times scaled by layer count and to 128 tokens")
    return tok_s, sample, detail


_emit = lambda text: print(text, flush=True)  # main() re-points this at the real stdout


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    for _ in range(max(args.warmup, 0)):
        cpu_port_run(args, threads)
    t0 = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        vals.append(cpu_port_run(args, threads))
    wall = time.perf_counter() - t0
    v, sample, detail = sorted(vals, key=lambda z: z[0])[len(vals) // 2]
    line = {"metric": "AV-prompt prefill+decode tokens/sec", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": max(args.steps, 1), "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": workload_config(args, note="CPU oracle port of the reference path (oracle/crab_oracle.py); "
                                                 "bs-1 bounded sample, see cpu_baseline.sample"),
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample, **detail},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


def workload_config(args, note=None):
    b = backbone(args)
    kv_gb = args.bs * (args.prompt_len + 574 + args.new_tokens) * 2 * b["kv_heads"] * b["head_dim"] * 2 * b["layers"] / 1e9
    c = {"workload": f"bs{args.bs}_avqa_10s-audio_8x224-video_{args.prompt_len}tok-prompt_{args.new_tokens}new_{args.backbone}7b-dims_hyperlora",
         "per_gpu_batch": args.bs, "frames": 8, "audio_segments": 10, "prompt_len": args.prompt_len,
         "seq_len_after_splice": args.prompt_len + 574, "new_tokens": args.new_tokens, "decoder_layers": b["layers"],
         "backbone": args.backbone,
         "l2_policy": f"working set per step (~14 GB weights + {kv_gb:.1f} GB KV) >> 126 MB L2; no explicit flush"}
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="crab_b200", choices=["crab_b200", "reference"])
    ap.add_argument("--bs", type=int, default=32)
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--layers", type=int, default=0, help="decoder layers (0 = the backbone's own depth)")
    ap.add_argument("--backbone", default="llama", choices=sorted(BACKBONES), help="llama = configs[2], qwen = configs[4]")
    ap.add_argument("--cpu-layers", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-pass", action="store_true",
                    help="for ncu launch lists: stop after the timed resident steps (NVTX range 'crab_timed'); prints no JSON")
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON): library chatter on fd 1 (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _emit
    _emit = lambda text: (real_stdout.write(text + "\n"), real_stdout.flush())

    if args.impl == "reference":
        run_reference_arm(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (crab_b200 arm) needs a B200: there is no CPU fallback. Use --impl reference for the CPU arm.")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from crab_b200 import ops
    from crab_b200.engine import BeatsConfig, ClipConfig, CrabConfig, CrabEngine, DecoderConfig, QformerConfig
    from crab_b200.models.unified_arch import full_manifest, special_token_ids
    if args.backbone == "qwen":
        from crab_b200.models.unified_qwen import UnifiedConfig, UnifiedForCausalLM
    else:
        from crab_b200.models.unified_llama import UnifiedConfig, UnifiedForCausalLM

    peaks, peaks_src = load_peaks()
    S = args.prompt_len + 574
    max_ctx = (S + args.new_tokens + 7) // 8 * 8
    b = backbone(args)
    n_layers = b["layers"]
    ids_map = special_token_ids(b["base_vocab"])
    cfg = CrabConfig(decoder=DecoderConfig(hidden=b["hidden"], inter=b["inter"], layers=n_layers, heads=b["heads"],
                                           kv_heads=b["kv_heads"], head_dim=b["head_dim"], vocab=b["vocab"],
                                           rope_theta=b["rope_theta"], qkv_bias=b["qkv_bias"]),
                     clip=ClipConfig(), beats=BeatsConfig(), qformer=QformerConfig(), max_ctx=max_ctx, special_ids=ids_map)
    sd = LazySynthSD(full_manifest(cfg), 42, dev)
    t0 = time.time()
    eng = CrabEngine(sd, cfg, dev)
    hf_cfg = UnifiedConfig(hidden_size=b["hidden"], intermediate_size=b["inter"], num_hidden_layers=n_layers,
                           num_attention_heads=b["heads"], num_key_value_heads=b["kv_heads"], vocab_size=b["vocab"])
    model = UnifiedForCausalLM.from_engine(hf_cfg, eng)  # the public API object a quick_start user holds
    torch.cuda.synchronize()
    t_load = time.time() - t0

    ids, X_host = make_inputs(args.bs, 8, 10, 98, args.prompt_len, b["base_vocab"], ids_map, rank)
    X_dev = [{k: v.to(dev) for k, v in x.items()} for x in X_host]
    h2d = sum(v.numel() * v.element_size() for x in X_host for v in x.values()) + sum(t.numel() * 8 for t in ids)
    d2h = args.bs * args.new_tokens * 8
    tokens_per_step = args.bs * (S + args.new_tokens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(ev=None):
        """inputs resident in HBM; the three phases are bracketed by CUDA events on the launching stream."""
        if ev is not None:
            ev[0].record()
        embeds, _, _ = eng.prepare_inputs(ids, X_dev)
        if ev is not None:
            ev[1].record()
        _, nxt = eng.prefill(embeds)
        out = torch.empty((args.bs, args.new_tokens), device=dev, dtype=torch.int64)
        out[:, 0].copy_(nxt)
        if ev is not None:
            ev[2].record()
        eng.begin_decode_cached(args.bs)
        for s_ in range(1, args.new_tokens):
            _, nxt = eng.decode_step()
            out[:, s_].copy_(nxt)
        if ev is not None:
            ev[3].record()
        if world > 1:
            gathered = torch.empty((world * args.bs, args.new_tokens), device=dev, dtype=torch.int64)
            dist.all_gather_into_tensor(gathered, out)
            out = gathered
        return out

    def step_e2e():
        """through the public API with HOST inputs: H2D of this step's inputs and D2H of the generated ids inside."""
        out = model.generate(batch_input_ids=ids, batch_labels=None, batch_X_modals=X_host, batch_task_names=["avqa"] * args.bs,
                             use_cache=True, max_new_tokens=args.new_tokens)
        return out.cpu()

    # ---- warm-up (also builds the decode graph once) -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        ref_out = step_resident()
    barrier()

    # ---- timed: resident ---------------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ops.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    ops.start_kernel_timing()
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    torch.cuda.nvtx.range_push("crab_timed")  # ncu --nvtx --nvtx-include "crab_timed/" profiles exactly the timed steps
    for i in range(args.steps):
        out = step_resident(evs[i])
    torch.cuda.nvtx.range_pop()
    e_end.record()
    barrier()
    if args.profile_pass:
        sys.stderr.write(f"profile pass: {args.steps} timed step(s), {e_start.elapsed_time(e_end) / args.steps:.1f} ms/step under the profiler\n")
        return
    kt = ops.stop_kernel_timing()
    launches = (ops.launch_count() - n0) // args.steps
    ms_total = e_start.elapsed_time(e_end)
    clocks = sampler.stop()
    deterministic = bool(torch.equal(out[: args.bs] if world > 1 else out, ref_out[: args.bs] if world > 1 else ref_out))
    t_enc = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    t_pf = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    t_dec = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    tm = torch.tensor([ms_total, t_enc, t_pf, t_dec], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total, t_enc, t_pf, t_dec = tm.tolist()
    ms_step = ms_total / args.steps
    value = world * tokens_per_step / (ms_step / 1e3)

    # ---- timed: e2e through the public API -----------------------------------------------------------------------------
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_e2e = step_e2e()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * tokens_per_step / t_e2e.item()
    e2e_same = bool(torch.equal(out_e2e.to(dev), out[: args.bs] if world > 1 else out))

    # ---- timed: e2e from RAW inputs: host waveforms + uint8 frames -> GPU fbank / fused normalise ----------
    from crab_b200.dataset import audio_processor as AP
    g_raw = torch.Generator(device="cpu").manual_seed(77 + rank)
    wave_host = (0.2 * torch.randn(args.bs * 10, 16000, generator=g_raw)).pin_memory()
    frames_host = [torch.randint(0, 256, (8, 224, 224, 3), generator=g_raw, dtype=torch.uint8).pin_memory() for _ in range(args.bs)]
    h2d_raw = wave_host.numel() * 4 + sum(f.numel() for f in frames_host) + sum(t.numel() * 8 for t in ids)

    def step_e2e_raw():
        fb = AP.preprocess(wave_host.to(dev, non_blocking=True))               # (bs*10, 98, 128) fp32 on the device
        X_raw = [{"<video>": frames_host[i], "<audio>": fb[i * 10:(i + 1) * 10]} for i in range(args.bs)]
        out = model.generate(batch_input_ids=ids, batch_labels=None, batch_X_modals=X_raw, batch_task_names=["avqa"] * args.bs,
                             use_cache=True, max_new_tokens=args.new_tokens)
        return out.cpu()

    step_e2e_raw()
    barrier()
    ops.start_kernel_timing()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_raw()
    torch.cuda.synchronize()
    t_raw = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    kraw = ops.stop_kernel_timing()
    if world > 1:
        dist.all_reduce(t_raw, op=dist.ReduceOp.MAX)
    frontend = {"e2e_raw_value": world * tokens_per_step / t_raw.item(), "unit": "tokens/s", "h2d_bytes_per_step": h2d_raw,
                "inputs": "pinned host fp32 waveforms (bs*10 x 16000) + uint8 frames (bs x 8x224x224x3)",
                "fbank_ms_per_step": kraw.get("crab_kaldi_fbank", {}).get("ms", 0.0) / args.steps,
                "patchify_u8_ms_per_step": kraw.get("crab_patchify_u8", {}).get("ms", 0.0) / args.steps}

    # ---- per-kernel split of one eager (un-graphed) decode step: same kernels as the graph -------------------------------
    embeds, _, _ = eng.prepare_inputs(ids, X_dev)
    eng.prefill(embeds)
    # take the split at the MEAN context of the timed decode (the cache rows up to there hold the previous run's keys:
    # same bytes, same timing), so the attention share and kv_bytes below refer to the same context length
    eng.cur_len = S + (args.new_tokens - 1) // 2 - 2
    eng.begin_decode(args.bs, use_graph=False)
    eng.decode_step()
    # Eager launches of 3-30 us kernels are host-bound (ctypes + two event records per launch): the device would idle between
    # kernels and the event intervals would measure the host.  Queue ~40 ms of filler GEMMs first so the host runs ahead and
    # the timed kernels execute back to back on the device.
    fa = torch.empty((16384, 4096), device=dev, dtype=torch.bfloat16).normal_()
    fw = torch.empty((8192, 4096), device=dev, dtype=torch.bfloat16).normal_()
    fo = torch.empty((16384, 8192), device=dev, dtype=torch.bfloat16)
    torch.cuda.synchronize()
    for _ in range(40):
        ops.gemm(fa, fw, out=fo)
    ops.start_kernel_timing()
    for _ in range(4):
        eng.decode_step()
    kd = ops.stop_kernel_timing()
    del fa, fw, fo
    for d in kd.values():
        for k in ("ms", "flops", "bytes"):
            d[k] /= 4
        d["launches"] //= 4

    if rank == 0:
        # ---------------- roofline of the dominant kernel -----------------------------------------------------------
        for d in kt.values():
            for k in ("ms", "flops", "bytes"):
                d[k] /= args.steps
            d["launches"] //= args.steps
        graph_step_ms = t_dec / max(args.new_tokens - 1, 1)
        # Decode split: the streaming kernels (>= 20 us each: GEMMs, attention) keep their measured device time; the event
        # pair around a 3 us kernel mostly measures launch gaps, so the light kernels share whatever the graph step has left.
        def _heavy(tag):
            return ("gemm" in tag) or ("attn" in tag) or ("chain" in tag)
        heavy_ms = sum(d["ms"] for k, d in kd.items() if _heavy(k))
        light_ms = sum(d["ms"] for k, d in kd.items() if not _heavy(k))
        if heavy_ms < graph_step_ms and light_ms > 0:
            light_scale = (graph_step_ms - heavy_ms) / light_ms
            for k, d in kd.items():
                if not _heavy(k):
                    d["ms"] *= light_scale
        dec_eager_ms = sum(d["ms"] for d in kd.values())
        shares = {k: d["ms"] for k, d in kt.items()}                      # encoders + prefill, measured in the timed region
        for k, d in kd.items():                                           # decode kernels: per-step split x number of steps
            shares["decode:" + k] = d["ms"] / max(dec_eager_ms, 1e-9) * t_dec
        top = max(shares, key=shares.get)
        ctx_mean = S + 1 + (args.new_tokens - 1) / 2.0
        kv_bytes = args.bs * ctx_mean * 2 * b["kv_heads"] * b["head_dim"] * 2 * n_layers
        wbytes = sum(L[k].numel() * 2 for L in eng.layers for k in ("wqkv", "wo", "wgu", "wd")) + eng.lm_head.numel() * 2
        decode_bytes = wbytes + kv_bytes
        if top.startswith("decode:"):
            d = kd[top[len("decode:"):]]
            scale = graph_step_ms / max(dec_eager_ms, 1e-9)
            if "gemm" in top or "chain" in top:
                ach = d["bytes"] / (d["ms"] * scale * 1e-3) / 1e9
            else:
                ach = kv_bytes / (d["ms"] * scale * 1e-3) / 1e9 if "attn_decode" in top else 0.0
            roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": None, "launches_per_step": d["launches"] * (args.new_tokens - 1),
                    "avg_launch_ms": d["ms"] * scale / max(d["launches"], 1)}
        else:
            d = kt[top]
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None,
                    "launches_per_step": d["launches"], "avg_launch_ms": d["ms"] / max(d["launches"], 1)}
        roof["peak_source"] = peaks_src
        roof["share_of_step"] = shares[top] / ms_step
        traffic = load_traffic()
        tkey = top[len("decode:"):] if top.startswith("decode:") else top
        if tkey in traffic:
            roof["traffic"] = traffic[tkey]["traffic_bytes_per_launch"]
            roof["traffic_note"] = f"ncu dram bytes of one launch, {traffic[tkey]['shape']} ({traffic.get('_source', '')})"
        # the other kernels that matter, same arithmetic (per kernel class, averaged over its launches in the step)
        def _roof(tag, dct, scale=1.0, bound="tensor"):
            d_ = dct.get(tag)
            if not d_ or d_["ms"] <= 0:
                return None
            t_ms = d_["ms"] * scale
            if bound == "tensor":
                a_ = d_["flops"] / (t_ms * 1e-3) / 1e12
                return {"bound": "tensor", "achieved": a_, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": a_ / peaks["bf16_tflops_sustained"], "ms_per_step": t_ms}
            a_ = d_["bytes"] / (t_ms * 1e-3) / 1e9
            return {"bound": "hbm", "achieved": a_, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": a_ / peaks["hbm_gbs"],
                    "ms_per_step": t_ms}
        dscale = graph_step_ms / max(dec_eager_ms, 1e-9)
        kd_step = {k: dict(v, ms=v["ms"] * (args.new_tokens - 1), bytes=v["bytes"] * (args.new_tokens - 1)) for k, v in kd.items()}
        for tag in ("crab_attn_decode", "crab_attn_decode_fused"):
            if tag in kd_step:
                kd_step[tag]["bytes"] = kv_bytes * (args.new_tokens - 1)
        rooflines = {
            "prefill_gemm_tcgen05<256>": _roof("gemm_bf16_tcgen05<256>", kt),
            "prefill_flash_attn_tcgen05<128>": _roof("crab_flash_attn_tcgen05<128>", kt),
            "decode_gemm_chain": _roof("crab_decode_chain", kd_step, dscale, "hbm"),
            "decode_attention": _roof("crab_attn_decode_fused" if "crab_attn_decode_fused" in kd_step else "crab_attn_decode",
                                      kd_step, dscale, "hbm"),
        }
        gemm_ms = sum(d["ms"] for k, d in kt.items() if k.startswith("gemm"))
        gemm_fl = sum(d["flops"] for k, d in kt.items() if k.startswith("gemm"))
        phases = {
            "encoders_bridge_splice_ms": t_enc, "decoder_prefill_ms": t_pf, "decode_127_steps_ms": t_dec,
            "prefill_tok_s": world * args.bs * S / ((t_enc + t_pf) / 1e3),
            "decoder_prefill_tok_s": world * args.bs * S / (t_pf / 1e3),
            "decode_tok_s": world * args.bs * (args.new_tokens - 1) / (t_dec / 1e3),
            "decode_step_ms": graph_step_ms,
            "prefill_gemm_tflops": gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
            "prefill_gemm_frac_of_peak": (gemm_fl / (gemm_ms * 1e-3) / 1e12) / peaks["bf16_tflops_sustained"] if gemm_ms else None,
            "prefill_algorithmic_tflop": algorithmic_prefill_tflop(b, S) * args.bs,
            "prefill_frac_of_tensor_roofline": (algorithmic_prefill_tflop(b, S) * 1e12 * args.bs / ((t_enc + t_pf) / 1e3))
            / (peaks["bf16_tflops_sustained"] * 1e12),
            "decode_bytes_per_step": decode_bytes, "decode_gbs": decode_bytes / (graph_step_ms * 1e-3) / 1e9,
            "decode_frac_of_hbm_roofline": decode_bytes / (graph_step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "kernel_ms_per_step": {k: round(v, 3) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])},
            "decode_split_note": "decode kernels: device time of 4 un-graphed steps at the mean context with the host running ahead; "
                                 "GEMM / attention keep their measured time, light kernels share the rest of the graph step. HBM peak = "
                                 "copy bandwidth (read + write); a read-only stream such as the decode attention can exceed it slightly",
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = os.cpu_count() or 1
                v, sample, detail = cpu_port_run(args, threads)
                cpu = {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample, **detail}
            except Exception as e:  # the CPU leg must never take the GPU numbers down with it
                cpu = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        line = {
            "metric": "AV-prompt prefill+decode tokens/sec", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "crab_b200.models.unified_llama.UnifiedForCausalLM.generate", "ids_equal_resident_run": e2e_same},
            "gpu_launches": int(launches), "roofline": roof, "rooflines": rooflines, "cpu_baseline": cpu, "phases": phases,
            "frontend": frontend,
            "deterministic_across_steps": deterministic, "weights": "random-init (seeded), generated on device",
            "load_s": round(t_load, 1),
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
