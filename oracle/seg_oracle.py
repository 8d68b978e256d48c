"""TEST INFRASTRUCTURE — CPU restatement of the reference's segmentation head, the post-processing half of
`UnifiedForCausalLM.generate_avs` (models/unified_llama.py:270-361 -> `postprocess_seg`, models/unified_arch.py:162-176).
Only tests/ may import this; the product path (crab_b200/seg.py, engine.py) never does.

Functional torch on a state dict with the reference's key names (`seg_module.*` of `UnifiedMetaModel`):

  seg_module_forward        <- SegModule.forward            models/multimodal_encoder.py:368-435 (inference branch, gt_mask=None)
  mask_decoder_predict      <- MaskDecoderMultiScale.predict_masks                    :1083-1143
  two_way_transformer       <- TwoWayTransformer.forward / TwoWayAttentionBlock.forward :1209-1331
  sam_attention             <- Attention.forward                                       :1368-1393
  query_generator           <- QueryGenerator.forward / AttentionLayer.forward         :1412-1445
                               (every layer is applied to the ORIGINAL queries, so only the last layer's output survives — kept)
  position_embedding_random <- PositionEmbeddingRandom.forward                         :822-835
  layer_norm_2d             <- LayerNorm2d.forward                                     :613-618

Pinning: tests/golden/seg_small.pt holds outputs of the REAL `SegModule` class imported from /root/reference
(oracle/make_seg_golden.py) on seeded weights (oracle/synth.py) and inputs; tests/test_seg_oracle_golden.py checks this file
against them.  Status: pinned.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd: SD, p: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm_2d(sd: SD, p: str, x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[p + ".weight"][:, None, None] * x + sd[p + ".bias"][:, None, None]


def position_embedding_random(gauss: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """-> (C, h, w) with C = 2 * gauss.shape[1]."""
    grid = torch.ones((h, w), dtype=gauss.dtype)
    y = (grid.cumsum(0) - 0.5) / h
    x = (grid.cumsum(1) - 0.5) / w
    c = (2 * torch.stack([x, y], -1) - 1) @ gauss
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], -1).permute(2, 0, 1)


def sam_attention(sd: SD, p: str, q, k, v, heads: int = 8):
    q, k, v = _lin(sd, p + ".q_proj", q), _lin(sd, p + ".k_proj", k), _lin(sd, p + ".v_proj", v)

    def split(t):
        b, n, c = t.shape
        return t.reshape(b, n, heads, c // heads).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), -1)
    o = (a @ v).transpose(1, 2)
    return _lin(sd, p + ".out_proj", o.reshape(o.shape[0], o.shape[1], -1))


def two_way_block(sd: SD, p: str, queries, keys, query_pe, key_pe, skip_first_layer_pe: bool):
    if skip_first_layer_pe:
        queries = sam_attention(sd, p + ".self_attn", queries, queries, queries)
    else:
        q = queries + query_pe
        queries = queries + sam_attention(sd, p + ".self_attn", q, q, queries)
    queries = _ln(sd, p + ".norm1", queries)
    q, k = queries + query_pe, keys + key_pe
    queries = _ln(sd, p + ".norm2", queries + sam_attention(sd, p + ".cross_attn_token_to_image", q, k, keys))
    mlp = _lin(sd, p + ".mlp.lin2", F.relu(_lin(sd, p + ".mlp.lin1", queries)))  # TwoWayTransformer passes activation=ReLU
    queries = _ln(sd, p + ".norm3", queries + mlp)
    q, k = queries + query_pe, keys + key_pe
    keys = _ln(sd, p + ".norm4", keys + sam_attention(sd, p + ".cross_attn_image_to_token", k, q, queries))
    return queries, keys


def two_way_transformer(sd: SD, p: str, image_embedding, image_pe, point_embedding, depth: int = 2):
    """image_embedding (B, C, h, w), image_pe same shape, point_embedding (B, N, C) -> (queries (B, N, C), keys (B, hw, C))."""
    keys = image_embedding.flatten(2).permute(0, 2, 1)
    key_pe = image_pe.flatten(2).permute(0, 2, 1).to(keys)
    queries = point_embedding
    for i in range(depth):
        queries, keys = two_way_block(sd, f"{p}.layers.{i}", queries, keys, point_embedding, key_pe, i == 0)
    q, k = queries + point_embedding, keys + key_pe
    queries = queries + sam_attention(sd, p + ".final_attn_token_to_image", q, k, keys)
    return _ln(sd, p + ".norm_final_attn", queries), keys


def _mha(sd: SD, p: str, q, k, v, heads: int = 8):
    """nn.MultiheadAttention(batch_first=True) forward, no masks, no dropout."""
    e = q.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q, w[:e], b[:e])
    k = F.linear(k, w[e:2 * e], b[e:2 * e])
    v = F.linear(v, w[2 * e:], b[2 * e:])

    def split(t):
        bsz, n, c = t.shape
        return t.reshape(bsz, n, heads, c // heads).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), -1)
    o = (a @ v).transpose(1, 2)
    return _lin(sd, p + ".out_proj", o.reshape(o.shape[0], o.shape[1], -1))


def query_generator(sd: SD, p: str, avs_query, sparse, num_layers: int = 2):
    out = avs_query
    for i in range(num_layers):  # reference quirk: each layer consumes the ORIGINAL avs_query (multimodal_encoder.py:1441-1444)
        lp = f"{p}.layers.{i}"
        q = _ln(sd, lp + ".norm1", avs_query + _mha(sd, lp + ".self_attn", avs_query, avs_query, avs_query))
        q = _ln(sd, lp + ".norm2", q + _mha(sd, lp + ".cross_attn", q, sparse, sparse))
        out = _ln(sd, lp + ".norm3", q + _lin(sd, lp + ".ffn.2", F.gelu(_lin(sd, lp + ".ffn.0", q))))
    return out


def _mlp3(sd: SD, p: str, x, conv: bool):
    for i in range(3):
        w, b = sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"]
        x = F.conv2d(x, w, b) if conv else F.linear(x, w, b)
        if i < 2:
            x = F.relu(x)
    return x


def _upscale(sd: SD, p: str, x):
    """ConvTranspose2d(k=2, s=2) -> LayerNorm2d -> GELU."""
    x = F.conv_transpose2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=2)
    return F.gelu(layer_norm_2d(sd, p + ".1", x))


def mask_decoder_predict(sd: SD, p: str, image_embeddings, image_pe, sparse, dense, level: int, previous_masks, task: str,
                         query_layers: int = 2, depth: int = 2):
    b = sparse.shape[0]
    avs_q = sd[p + ".avs_query_tokens.weight"].unsqueeze(0).expand(b, -1, -1)
    n_query = avs_q.shape[1]
    tokens = query_generator(sd, p + ".query_generator", avs_q, sparse, query_layers)
    tokens = tokens + sd[p + ".level_embed.weight"][level].view(1, 1, -1)
    src = torch.repeat_interleave(image_embeddings, b, dim=0)
    if level > 0:
        src = _upscale(sd, p + ".upsample_2x", src)
        h, w = src.shape[-2:]
        prev = previous_masks.mean(dim=1)
        src = (torch.repeat_interleave(prev[:, None], 256, dim=1).sigmoid() + 1) * src
        image_pe = position_embedding_random(sd[p + ".pe1.positional_encoding_gaussian_matrix"], h, w).unsqueeze(0)
        dense = F.interpolate(dense.float(), size=(h, w), mode="bilinear", align_corners=False).to(dense)
    src = src + dense
    pos = torch.repeat_interleave(image_pe, b, dim=0)
    bb, c, h, w = src.shape
    hs, keys = two_way_transformer(sd, f"{p}.transformer.{level}", src, pos, tokens, depth)
    q_out = _mlp3(sd, p + ".hyper_mlp", hs[:, :n_query], conv=False)             # (b, Q, C/8)
    up = _upscale(sd, p + ".output_upscaling", keys.transpose(1, 2).reshape(bb, c, h, w))  # (b, C/8, 2h, 2w)
    b2, c2, h2, w2 = up.shape
    masks = (q_out @ up.view(b2, c2, h2 * w2)).view(b2, -1, h2, w2)                # (b, Q, 2h, 2w)
    masks = _mlp3(sd, p + ".hyper_mlp_out", masks, conv=True)                      # (b, C/8, 2h, 2w)
    head = ".avss_classifier.weight" if task == "avss" else ".ms3_s4_classfier.weight"
    return F.conv2d(masks, sd[p + head])


def image_feature_neck(sd: SD, p: str, x):
    x = F.conv2d(x, sd[p + ".0.weight"])
    x = layer_norm_2d(sd, p + ".1", x)
    x = F.conv2d(x, sd[p + ".2.weight"], padding=1)
    return layer_norm_2d(sd, p + ".3", x)


def seg_module_forward(sd: SD, pred_embeddings: torch.Tensor, multi_scale_feats: Sequence[torch.Tensor], task_names: List[str],
                       p: str = "seg_module", grid: int = 16, scales: int = 2, tokens_per_scale: int = 3, low_res: int = 112,
                       image_size: int = 224) -> List[torch.Tensor]:
    """pred_embeddings (bs, scales * tokens_per_scale, d_model): last-layer hidden states at the <mask_i> steps;
    multi_scale_feats: `scales` tensors (bs, n_img * grid^2, 1024) of ViT taps -> list of (num_classes, 224, 224) masks."""
    x = _lin(sd, p + ".text_hidden_fcs.0.2", F.relu(_lin(sd, p + ".text_hidden_fcs.0.0", pred_embeddings)))
    bs, n, dim = x.shape
    obj = n // (scales * tokens_per_scale)
    x = x.reshape(bs, obj, scales, tokens_per_scale, dim)
    fused = torch.zeros(bs, obj, scales, dim, dtype=x.dtype)
    for i in range(tokens_per_scale):
        fused = fused + (1.0 / tokens_per_scale) * x[:, :, :, i]     # multiseg_scalar: unregistered Parameters, always 1/3
    grids = []
    for f in multi_scale_feats:
        b_, n_, d_ = f.shape
        g = f.reshape(b_, n_ // (grid * grid), grid, grid, d_).permute(0, 1, 4, 2, 3)[:, 0]
        grids.append(g)
    grids = torch.stack(grids, 1)                                    # (bs, level, 1024, grid, grid)
    dense_pe = position_embedding_random(sd[p + ".pe_layer.positional_encoding_gaussian_matrix"], grid, grid).unsqueeze(0)
    out = []
    for i in range(bs):
        sparse = fused[i]                                            # (obj, scales, 256)
        dense = sd[p + ".no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(sparse.shape[0], -1, grid, grid)
        img = image_feature_neck(sd, p + ".image_feature_neck", grids[i])   # (level, 256, grid, grid)
        ncls = 71 if task_names[i] == "avss" else 1
        low = torch.zeros(sparse.shape[0], ncls, low_res, low_res, dtype=img.dtype)
        prev = None
        for l in range(scales):
            prev = mask_decoder_predict(sd, p + ".mask_decoder", img[l].unsqueeze(0), dense_pe.to(img), sparse[:, l].unsqueeze(1),
                                        dense, l, prev, task_names[i])
            low = low + (1.0 / scales) * F.interpolate(prev.float(), (low_res, low_res), mode="bilinear",
                                                       align_corners=False).to(prev)       # multiscale_scalar: always 1/2
        mask = F.interpolate(low.float(), (image_size, image_size), mode="bilinear", align_corners=False).to(low)
        out.append(mask[0])
    return out


# ---- generate_avs: the LLM half (models/unified_llama.py:270-361) on top of oracle/crab_oracle.py ----------------------------
def generate_avs(sd: SD, input_ids: torch.Tensor, X_modals: dict, cfg, max_new_tokens: int, task: str,
                 forced_output_ids: torch.Tensor = None, grid: int = 16):
    """One sample, as the reference (`bs == 1`): prepare inputs with the ViT taps of the '<image>' (select_layers[0], [1];
    models/unified_arch.py:243-247), greedy generation keeping the last layer's final-normed hidden states of every forward
    pass (HF `output_hidden_states`), pair entry t with generated token t + 1 being a `<mask_i>` token (:331-340), keep the last
    six rows (:346-348), run the segmentation head.  `forced_output_ids` (n,) teacher-forces the generated sequence.
    Returns dict(output_ids, pred_embeddings, pred_masks)."""
    from oracle import crab_oracle as O

    prep = O.prepare_multimodal_inputs(sd, [input_ids], [X_modals], cfg)
    taps = O.visual_encoder(sd, X_modals["<image>"].unsqueeze(0), cfg.clip, cfg.select_layers)
    feats = [taps[0][:, : grid * grid], taps[1][:, : grid * grid]]
    emb = prep["inputs_embeds"].to(sd["model.embed_tokens.weight"].dtype)
    h, cache = O.decoder_forward(sd, emb, cfg.decoder)
    hidden = [h]                                            # hidden_states[0][-1]: (1, S, D)
    ids = []
    logits = O.lm_head(sd, h[:, -1])
    for step in range(max_new_tokens):
        nxt = logits.argmax(-1) if forced_output_ids is None else forced_output_ids[step:step + 1]
        ids.append(nxt)
        if step + 1 == max_new_tokens:
            break
        h, cache = O.decoder_forward(sd, sd["model.embed_tokens.weight"][nxt].unsqueeze(1), cfg.decoder, cache)
        hidden.append(h)                                    # (1, 1, D)
        logits = O.lm_head(sd, h[:, -1])
    output_ids = torch.stack(ids, 1)                        # (1, n)
    mask_ids = [cfg.special_ids[f"<mask_{i}>"] for i in range(6)]
    flags = [int(t) in mask_ids for t in output_ids[0, 1:].tolist()]
    pred = [hs for f, hs in zip(flags, hidden) if f]
    res = {"output_ids": output_ids, "pred_embeddings": None, "pred_masks": None}
    if not pred:
        return res
    pred = torch.cat(pred, dim=1)
    if pred.shape[1] > 6:
        pred = pred[:, -6:]
    elif pred.shape[1] < 6:
        return res
    res["pred_embeddings"] = pred
    res["pred_masks"] = seg_module_forward(sd, pred, feats, [task], p="model.seg_module", grid=grid)
    return res
