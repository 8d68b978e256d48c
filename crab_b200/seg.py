"""Segmentation head on the B200: the `SegModule` half of `generate_avs`
(models/multimodal_encoder.py:268-543 SegModule, :891-1143 MaskDecoderMultiScale, :1163-1393 TwoWayTransformer / Attention,
:1396-1445 QueryGenerator) as a sequence of C-ABI kernel calls.

Layout: every feature map is token-major — one bf16 row per pixel, channels contiguous — so
  * 1x1 convs, hyper-MLPs, class heads and linears are `crab_gemm_bf16` calls,
  * ConvTranspose2d(k=2, s=2) is a GEMM whose output columns are ordered (dy, dx, c) followed by a row gather (pixel shuffle),
  * the 3x3 conv is `crab_im2col3x3` + GEMM, LayerNorm2d is the row LayerNorm (eps 1e-6),
  * attention (8 heads x 16 / 32 channels, <= 1024 keys) is `crab_small_attn`,
  * bilinear resizing / multi-scale accumulation run on fp32 token-major maps (`crab_bilinear_f32`).
Input-independent parts are folded at load time: the query generator's self-attention acts on the constant query embeddings
(and only its LAST layer reaches the output — each layer is fed the original queries, :1441-1444), its cross-attention has a
single key (softmax == 1, so it is the linear map out_proj(v_proj(s))), and the positional-encoding grids are tables.

Python here is plumbing only; there is no torch arithmetic on the data path (load-time table building aside).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Sequence

import torch

from . import ops

SD = Dict[str, torch.Tensor]
E, HEADS, NQ, NQ_PAD, KMIN = 256, 8, 300, 304, 64


def _bf(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


class _Lin:
    __slots__ = ("w", "b")

    def __init__(self, w, b=None):
        self.w, self.b = w, b


def _pe_table(gauss: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """PositionEmbeddingRandom.forward (:822-835) as a token-major [h*w, 2*F] fp32 table (built once at load)."""
    g = gauss.detach().float().cpu()
    grid = torch.ones((h, w))
    y = (grid.cumsum(0) - 0.5) / h
    x = (grid.cumsum(1) - 0.5) / w
    c = (2 * torch.stack([x, y], -1) - 1) @ g
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], -1).reshape(h * w, -1)


def _shuffle_index(h: int, w: int, dev) -> torch.Tensor:
    """Row gather that turns a conv-transpose GEMM output [(y, x), (dy, dx, c)] viewed as [(y, x, dy, dx), c] into the
    token-major map at (2h, 2w): destination token (2y + dy, 2x + dx) reads source row ((y * w + x) * 4 + dy * 2 + dx)."""
    oy = torch.arange(2 * h).view(-1, 1).expand(2 * h, 2 * w)
    ox = torch.arange(2 * w).view(1, -1).expand(2 * h, 2 * w)
    src = ((oy // 2) * w + (ox // 2)) * 4 + (oy % 2) * 2 + (ox % 2)
    return src.reshape(-1).to(torch.int64).to(dev)


class SegHead:
    def __init__(self, sd: SD, device, prefix: str = "seg_module", grid: int = 16, scales: int = 2, tokens_per_scale: int = 3,
                 low_res: int = 112, image_size: int = 224):
        self.dev, self.grid, self.scales, self.tps, self.low_res, self.image = device, grid, scales, tokens_per_scale, low_res, image_size
        dev, p = device, prefix + "."
        g = lambda k: sd[p + k]  # noqa: E731

        def lin(name, bias=True):
            return _Lin(_bf(g(name + ".weight"), dev), _f32(g(name + ".bias"), dev) if bias else None)

        def ln(name):
            return (_f32(g(name + ".weight"), dev), _f32(g(name + ".bias"), dev))

        # ---- sparse prompt: text_hidden_fcs = Linear -> ReLU -> Linear; the mean over the 3 tokens of a scale is folded into
        # the second GEMM (A = [scales, 3*D] view of the hidden rows, B = [W2 | W2 | W2] / 3)
        self.fc0 = lin("text_hidden_fcs.0.0")
        w2 = g("text_hidden_fcs.0.2.weight").float()
        self.fc2 = _Lin(_bf(torch.cat([w2] * tokens_per_scale, 1) / tokens_per_scale, dev), _f32(g("text_hidden_fcs.0.2.bias"), dev))
        self.d_model = w2.shape[1]
        # ---- image_feature_neck
        self.neck0 = _Lin(_bf(g("image_feature_neck.0.weight").reshape(E, -1), dev))
        self.neck_ln1 = ln("image_feature_neck.1")
        self.neck2 = _Lin(_bf(g("image_feature_neck.2.weight").permute(0, 2, 3, 1).reshape(E, 9 * E), dev))  # columns (ky, kx, c)
        self.neck_ln3 = ln("image_feature_neck.3")
        self.no_mask = _bf(g("no_mask_embed.weight").reshape(1, E), dev)
        md = "mask_decoder."
        self.pe = {grid: _bf(_pe_table(g("pe_layer.positional_encoding_gaussian_matrix"), grid, grid), dev),
                   2 * grid: _bf(_pe_table(g(md + "pe1.positional_encoding_gaussian_matrix"), 2 * grid, 2 * grid), dev)}
        self.level_embed = _bf(g(md + "level_embed.weight"), dev)  # [scales, E]
        # ---- query generator, last layer only; constant self-attention branch folded at load (fp32)
        n_layers = 1 + max(int(k.split(".layers.")[1].split(".")[0]) for k in sd if k.startswith(p + md + "query_generator.layers."))
        qp = f"{md}query_generator.layers.{n_layers - 1}."
        avs = g(md + "avs_query_tokens.weight").float()
        wi, bi = g(qp + "self_attn.in_proj_weight").float(), g(qp + "self_attn.in_proj_bias").float()
        q, k_, v = (avs @ wi[i * E:(i + 1) * E].T + bi[i * E:(i + 1) * E] for i in range(3))
        hd = E // HEADS
        sp = lambda t: t.reshape(NQ, HEADS, hd).transpose(0, 1)  # noqa: E731
        att = torch.softmax(sp(q) @ sp(k_).transpose(-1, -2) / math.sqrt(hd), -1) @ sp(v)
        att = att.transpose(0, 1).reshape(NQ, E) @ g(qp + "self_attn.out_proj.weight").float().T + g(qp + "self_attn.out_proj.bias").float()
        q1 = torch.nn.functional.layer_norm(avs + att, (E,), g(qp + "norm1.weight").float(), g(qp + "norm1.bias").float())
        self.q1 = _bf(q1, dev)
        wc, bc = g(qp + "cross_attn.in_proj_weight").float(), g(qp + "cross_attn.in_proj_bias").float()
        wo, bo = g(qp + "cross_attn.out_proj.weight").float(), g(qp + "cross_attn.out_proj.bias").float()
        self.qg_cross = _Lin(_bf(wo @ wc[2 * E:], dev), _f32(wo @ bc[2 * E:] + bo, dev))  # one key: softmax == 1
        self.qg_ln2, self.qg_ln3 = ln(qp + "norm2"), ln(qp + "norm3")
        self.qg_ffn0, self.qg_ffn2 = lin(qp + "ffn.0"), lin(qp + "ffn.2")
        # ---- two-way transformers (one per scale)
        self.tw = []
        for l in range(scales):
            tp = f"{md}transformer.{l}."
            depth = 1 + max(int(k.split(".layers.")[1].split(".")[0]) for k in sd if k.startswith(p + tp + "layers."))

            def attn(name):
                return dict(q=lin(name + ".q_proj"), k=lin(name + ".k_proj"), v=lin(name + ".v_proj"), o=lin(name + ".out_proj"))

            layers = [dict(self_attn=attn(f"{tp}layers.{i}.self_attn"), t2i=attn(f"{tp}layers.{i}.cross_attn_token_to_image"),
                           i2t=attn(f"{tp}layers.{i}.cross_attn_image_to_token"), lin1=lin(f"{tp}layers.{i}.mlp.lin1"),
                           lin2=lin(f"{tp}layers.{i}.mlp.lin2"), n1=ln(f"{tp}layers.{i}.norm1"), n2=ln(f"{tp}layers.{i}.norm2"),
                           n3=ln(f"{tp}layers.{i}.norm3"), n4=ln(f"{tp}layers.{i}.norm4")) for i in range(depth)]
            self.tw.append(dict(layers=layers, final=attn(tp + "final_attn_token_to_image"), nf=ln(tp + "norm_final_attn")))
        # ---- hyper networks, up-scaling, class heads
        self.hyper = [lin(f"{md}hyper_mlp.layers.{i}") for i in range(3)]
        ho = []
        for i in range(3):
            w = g(f"{md}hyper_mlp_out.layers.{i}.weight").float().reshape(g(f"{md}hyper_mlp_out.layers.{i}.weight").shape[0], -1)
            if i == 0:  # K = 300 queries, padded to 304 zero columns
                w = torch.cat([w, torch.zeros(w.shape[0], NQ_PAD - NQ)], 1)
            ho.append(_Lin(_bf(w, dev), _f32(g(f"{md}hyper_mlp_out.layers.{i}.bias"), dev)))
        self.hyper_out = ho

        def convT(name):  # ConvTranspose2d weight [Cin, Cout, 2, 2] -> GEMM rows (dy, dx, co), bias replicated per tap
            w = g(name + ".weight").float()
            cin, cout = w.shape[0], w.shape[1]
            return _Lin(_bf(w.permute(2, 3, 1, 0).reshape(4 * cout, cin), dev), _f32(g(name + ".bias").float().repeat(4), dev)), cout

        self.up_out, self.up_out_c = convT(md + "output_upscaling.0")
        self.up_out_ln = ln(md + "output_upscaling.1")
        self.up2x, self.up2x_c = convT(md + "upsample_2x.0")
        self.up2x_ln = ln(md + "upsample_2x.1")

        def head(name):  # class heads: rows padded to a multiple of 8
            w = g(name).float().reshape(g(name).shape[0], -1)
            n = (w.shape[0] + 7) // 8 * 8
            wp = torch.zeros(n, KMIN)
            wp[: w.shape[0], : w.shape[1]] = w
            return _Lin(_bf(wp, dev)), w.shape[0]

        self.head_s4, self.n_s4 = head(md + "ms3_s4_classfier.weight")
        self.head_avss, self.n_avss = head(md + "avss_classifier.weight")
        self.shuffle = {grid: _shuffle_index(grid, grid, dev), 2 * grid: _shuffle_index(2 * grid, 2 * grid, dev)}

    # ---- building blocks -------------------------------------------------------------------------------------------------
    def _attention(self, A, q_in, k_in, v_in, heads=HEADS):
        """Attention.forward: projections (GEMM), softmax(q k^T / sqrt(hd)) v, out_proj."""
        q = ops.gemm(q_in, A["q"].w, bias=A["q"].b)
        k = ops.gemm(k_in, A["k"].w, bias=A["k"].b)
        v = ops.gemm(v_in, A["v"].w, bias=A["v"].b)
        o = torch.empty((q.shape[0], q.shape[1]), device=q.device, dtype=torch.bfloat16)
        ops.small_attn(q, k, v, o, heads, q.shape[1] // heads)
        return o, A["o"]

    def _two_way(self, T, queries, keys, key_pe):
        """TwoWayTransformer.forward on token-major maps: queries [300, E] (also the query PE), keys [hw, E]."""
        qpe = queries
        for i, L in enumerate(T["layers"]):
            if i == 0:   # skip_first_layer_pe: the self-attention output REPLACES the queries
                o, po = self._attention(L["self_attn"], queries, queries, queries)
                y = ops.gemm(o, po.w, bias=po.b)
            else:
                qq = ops.elementwise(queries, ops.EW_ADD, b=qpe)
                o, po = self._attention(L["self_attn"], qq, qq, queries)
                y = ops.gemm(o, po.w, bias=po.b, residual=queries)
            queries = ops.layernorm(y, *L["n1"], 1e-5)
            qq = ops.elementwise(queries, ops.EW_ADD, b=qpe)
            kk = ops.elementwise(keys, ops.EW_ADD, b=key_pe)
            o, po = self._attention(L["t2i"], qq, kk, keys)
            queries = ops.layernorm(ops.gemm(o, po.w, bias=po.b, residual=queries), *L["n2"], 1e-5)
            hmid = ops.elementwise(ops.gemm(queries, L["lin1"].w, bias=L["lin1"].b), ops.EW_RELU)
            queries = ops.layernorm(ops.gemm(hmid, L["lin2"].w, bias=L["lin2"].b, residual=queries), *L["n3"], 1e-5)
            qq = ops.elementwise(queries, ops.EW_ADD, b=qpe)
            o, po = self._attention(L["i2t"], kk, qq, queries)
            keys = ops.layernorm(ops.gemm(o, po.w, bias=po.b, residual=keys), *L["n4"], 1e-5)
        qq = ops.elementwise(queries, ops.EW_ADD, b=qpe)
        kk = ops.elementwise(keys, ops.EW_ADD, b=key_pe)
        o, po = self._attention(T["final"], qq, kk, keys)
        queries = ops.layernorm(ops.gemm(o, po.w, bias=po.b, residual=queries), *T["nf"], 1e-5)
        return queries, keys

    def _conv_transpose(self, x, h, w, lin_, cout, ln_):
        """ConvTranspose2d(k=2, s=2) -> LayerNorm2d -> GELU on a token-major [h*w, Cin] map -> [(2h)*(2w), max(cout, 64)]
        (columns beyond cout are zero: every GEMM on this path keeps K >= 64, one full k-block)."""
        gm = ops.gemm(x, lin_.w, bias=lin_.b)                                   # [hw, 4*cout], columns (dy, dx, c)
        out = torch.empty((4 * h * w, cout), device=x.device, dtype=torch.bfloat16)
        ops.gather_rows(gm.view(4 * h * w, cout), out, 4 * h * w, cout, src_rows=self.shuffle[h])
        normed = ops.layernorm(out, *ln_, 1e-6)
        if cout >= KMIN:
            return ops.elementwise(normed, ops.EW_GELU)
        wide = torch.zeros((4 * h * w, KMIN), device=x.device, dtype=torch.bfloat16)
        ops.elementwise(normed, ops.EW_GELU, out=wide[:, :cout])
        return wide

    def _query_tokens(self, sparse_row, level):
        """QueryGenerator (last layer) + level embedding for one (sample, level): sparse_row bf16 [8, E], row 0 valid."""
        c = ops.gemm(sparse_row, self.qg_cross.w, bias=self.qg_cross.b)         # out_proj(v_proj(s)): the single-key attention
        q2 = ops.layernorm(ops.elementwise(self.q1, ops.EW_ADD, b=c[:1]), *self.qg_ln2, 1e-5)
        f = ops.gemm(q2, self.qg_ffn0.w, bias=self.qg_ffn0.b, act=ops.ACT_GELU)
        q3 = ops.layernorm(ops.gemm(f, self.qg_ffn2.w, bias=self.qg_ffn2.b, residual=q2), *self.qg_ln3, 1e-5)
        return ops.elementwise(q3, ops.EW_ADD, b=self.level_embed[level:level + 1])

    def _predict(self, img, sparse_row, level, prev_masks, prev_classes, task):
        """MaskDecoderMultiScale.predict_masks for one sample and level -> fp32 token-major masks [(2h)^2, classes_padded]."""
        g0 = self.grid
        tokens = self._query_tokens(sparse_row, level)
        if level == 0:
            src, h = img, g0
        else:
            src = self._conv_transpose(img, g0, g0, self.up2x, self.up2x_c, self.up2x_ln)      # [1024, E] at 2*grid
            h = 2 * g0
            gate = ops.row_mean_f32(prev_masks, prev_classes)                                   # class mean of the previous masks
            src = ops.elementwise(src, ops.EW_GATE, gate=gate)
        src = ops.elementwise(src, ops.EW_ADD, b=self.no_mask)     # dense prompt = no_mask_embed everywhere (its bilinear resize too)
        hs, keys = self._two_way(self.tw[level], tokens, src, self.pe[h])
        q = hs
        for i, L in enumerate(self.hyper):
            q = ops.gemm(q, L.w, bias=L.b)
            if i < 2:
                q = ops.elementwise(q, ops.EW_RELU)
        qpad = torch.zeros((NQ_PAD, KMIN), device=q.device, dtype=torch.bfloat16)            # 300 -> 304 rows (GEMM N % 8), K 32 -> 64
        ops.gather_rows(q, qpad, NQ, q.shape[1])
        up = self._conv_transpose(keys, h, h, self.up_out, self.up_out_c, self.up_out_ln)      # [(2h)^2, 64] (32 live columns)
        m = ops.gemm(up, qpad)                                                                  # [(2h)^2, 304]: masks per query
        for i, L in enumerate(self.hyper_out):
            if i < 2:
                m = ops.elementwise(ops.gemm(m, L.w, bias=L.b), ops.EW_RELU)
            else:
                wide = torch.zeros((m.shape[0], KMIN), device=m.device, dtype=torch.bfloat16)
                ops.gemm(m, L.w, bias=L.b, out=wide[:, : L.w.shape[0]])
                m = wide
        hd_ = self.head_avss if task == "avss" else self.head_s4
        return ops.gemm(m, hd_.w, out_dtype=torch.float32)

    # ---- SegModule.forward (inference branch) ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, pred_embeddings: torch.Tensor, multi_scale_feats: Sequence[torch.Tensor], task_names: List[str]):
        """pred_embeddings bf16 [bs, scales*tokens_per_scale, d_model]; multi_scale_feats: `scales` bf16 tensors
        [bs, n_img*grid^2, 1024] -> list of fp32 [num_classes, image, image] masks (one object per sample, as quick_start).
        On a GPU the ~190 small launches of one object are captured in a CUDA graph per task kind and replayed (the head is
        launch-bound: 3.9 ms eager); CRAB_SEG_GRAPH=0 keeps the eager path."""
        bs, n, D = pred_embeddings.shape
        assert n == self.scales * self.tps and D == self.d_model, "one object per sample: scales * tokens_per_scale hidden states"
        g0, out = self.grid, []
        use_graph = pred_embeddings.is_cuda and os.environ.get("CRAB_SEG_GRAPH", "1") != "0"
        for i in range(bs):
            feats_i = [multi_scale_feats[l][i][: g0 * g0] for l in range(self.scales)]                        # first image's grid
            if use_graph:
                out.append(self._forward_graph(pred_embeddings[i], feats_i, task_names[i]))
            else:
                out.append(self._forward_one(pred_embeddings[i], feats_i, task_names[i]))
        return out

    def _forward_graph(self, pred_i: torch.Tensor, feats_i: List[torch.Tensor], task: str) -> torch.Tensor:
        key = "avss" if task == "avss" else "s4"
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        if key not in self._graphs:
            p_in = torch.empty_like(pred_i, memory_format=torch.contiguous_format)
            f_in = [torch.empty_like(f, memory_format=torch.contiguous_format) for f in feats_i]
            p_in.copy_(pred_i)
            for d_, s_ in zip(f_in, feats_i):
                d_.copy_(s_)
            side = torch.cuda.Stream(device=pred_i.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._forward_one(p_in, f_in, key)            # warm-up outside capture (kernel attributes, allocator)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(gr):
                o = self._forward_one(p_in, f_in, key)
            kernels = ops.launch_count() - n0
            ops.count_launches(-kernels)                       # capture launches nothing
            self._graphs[key] = (gr, p_in, f_in, o, kernels)
        gr, p_in, f_in, o, kernels = self._graphs[key]
        p_in.copy_(pred_i)
        for d_, s_ in zip(f_in, feats_i):
            d_.copy_(s_)
        gr.replay()
        ops.count_launches(kernels)
        return o.clone()

    def _forward_one(self, pred_i: torch.Tensor, feats_i: List[torch.Tensor], task: str) -> torch.Tensor:
        D, g0 = self.d_model, self.grid
        hid = ops.elementwise(ops.gemm(pred_i, self.fc0.w, bias=self.fc0.b), ops.EW_RELU)                 # [6, D]
        abuf = torch.zeros((8, self.tps * D), device=hid.device, dtype=torch.bfloat16)
        ops.gather_rows(hid.view(self.scales, self.tps * D), abuf, self.scales, self.tps * D)              # [scales, 3D] view
        sparse = ops.gemm(abuf, self.fc2.w, bias=self.fc2.b)                                               # rows 0..scales-1
        classes = self.n_avss if task == "avss" else self.n_s4
        low = torch.zeros((self.low_res * self.low_res, classes), device=hid.device, dtype=torch.float32)
        prev = None
        for l in range(self.scales):
            f = feats_i[l]
            x = ops.layernorm(ops.gemm(f, self.neck0.w), *self.neck_ln1, 1e-6)
            x = ops.layernorm(ops.gemm(ops.im2col3x3(x, g0, g0), self.neck2.w), *self.neck_ln3, 1e-6)
            srow = torch.zeros((8, E), device=hid.device, dtype=torch.bfloat16)
            ops.gather_rows(sparse[l:l + 1], srow, 1, E)
            prev = self._predict(x, srow, l, prev, classes, task)
            side = 2 * g0 * (l + 1)                                                                         # 32, 64
            ops.bilinear_f32(prev, side, side, self.low_res, self.low_res, classes, out=low, alpha=1.0 / self.scales, beta=1.0)
        return ops.bilinear_f32(low, self.low_res, self.low_res, self.image, self.image, classes, nchw_out=True)
