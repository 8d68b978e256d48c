/*
 * crab_b200.h — C ABI of the B200-native kernels behind Crab's AV-prompt -> prefill -> decode hot path.
 *
 * The reference (GeWu-Lab/Crab) has no native code and no FFI: its hot path is a sequence of PyTorch/HF library
 * ops.  Each entry point below replaces one such op sequence; the "replaces:" line cites it (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer; the library keeps no global state besides a TMA-descriptor cache;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - activations and weights are bf16 (uint16 storage), row-major, fp32 accumulation everywhere;
 *   - return value: 0 = ok, negative = error (CRAB_ERR_*); crab_last_error() gives the message (thread-local);
 *   - nothing here throws, exits or falls back to a CPU path.
 */
#ifndef CRAB_B200_H_
#define CRAB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRAB_OK 0
#define CRAB_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define CRAB_ERR_CUDA (-2)    /* CUDA runtime / driver error */
#define CRAB_ERR_ARCH (-3)    /* device is not sm_100 */

const char* crab_last_error(void);
/* Library / device probe: returns CRAB_OK when device `dev` is compute capability 10.x. */
int crab_init(int dev);
int crab_version(void);
/* Programmatic dependent launch for the decode-chain kernels: bit mask over kernel classes (env CRAB_PDL gives the
 * initial value, default 0): 1 = weight-streaming GEMM, 2 = row norm/LoRA pre-pass, 4 = RoPE + KV append,
 * 8 = decode attention, 16 = the other light kernels (norm, gather, arg-max, counters); modifier 32 = the decode
 * attention releases its dependents after its streaming loop instead of at its top.  Read at launch time. */
int crab_set_pdl(int mask);   /* launch policy of the CALLING THREAD (thread-local, like crab_set_gemm_2cta) */

/* ------------------------------------------------------------------------------------------------------------------
 * Dense linear:  C[M,N] = epilogue( A[M,K] . B[N,K]^T )       (tcgen05 + TMEM + TMA, persistent, warp-specialised)
 * replaces: every nn.Linear / F.linear on the path — HF LlamaAttention/LlamaMLP projections
 *           (models/modeling_llama.py:239-271, 286-464), CLIPAttention/CLIPMLP (transformers clip), BEATs
 *           q/k/v/out/fc1/fc2 (models/beats/backbone.py:249-273, 495-500), Q-Former dense layers
 *           (models/Qformer.py:171-377), build_mlp (models/multimodal_encoder.py:25-30), the patch-embed convs as
 *           GEMMs (models/beats/BEATs.py:148-158), lm_head (models/unified_llama.py:35), and — through K-extension
 *           columns — the hyper-LoRA side path (peft_hyper/tuners/lora.py:338-350).
 * ---------------------------------------------------------------------------------------------------------------- */
enum crab_act {
  CRAB_ACT_NONE = 0,
  CRAB_ACT_QUICK_GELU = 1, /* x * sigmoid(1.702 x)           (CLIP MLP) */
  CRAB_ACT_GELU = 2,       /* 0.5 x (1 + erf(x / sqrt 2))    (BEATs, Q-Former, build_mlp) */
  CRAB_ACT_SWIGLU = 3,     /* packed gate|up columns in 64-wide groups -> silu(gate) * up, N_out = N / 2 */
  CRAB_ACT_LORA_Z = 4      /* 11-wide groups (3 router logits, 8 lora_A outputs) -> 24-wide r_i * u_j groups */
};
enum crab_dtype { CRAB_BF16 = 0, CRAB_F32 = 1 };

typedef struct crab_gemm_args {
  const void* A;        /* bf16 [M, lda]                                                         */
  const void* B;        /* bf16 [N, ldb]  (nn.Linear weight layout: out_features x in_features)   */
  void* C;              /* out  [M, ldc]  bf16 or fp32                                            */
  const float* bias;    /* fp32 [N] or NULL                                                       */
  const void* residual; /* bf16 [M, ldr] or NULL: C = act(acc + bias) + res_scale * residual      */
  int32_t M, N, K;
  int32_t lda, ldb, ldc, ldr; /* row strides in elements; lda, ldb multiples of 8                 */
  float res_scale;
  float out_scale;      /* multiplies act(acc + bias) before the residual add (1.0 normally)      */
  int32_t act;          /* enum crab_act                                                          */
  int32_t out_dtype;    /* enum crab_dtype                                                        */
  int32_t block_n;      /* 0 = auto; else 64 / 128 / 256                                          */
  int32_t max_ctas;     /* 0 = one persistent CTA per SM                                          */
} crab_gemm_args;

int crab_gemm_bf16(const crab_gemm_args* args, void* stream);
/* CTA-pair (tcgen05 cta_group::2, 256 x 256 tile on two SMs) policy for crab_gemm_bf16: 0 never, 1 auto (default: K >= 2048,
 * M >= 16384, >= 4 waves of pairs), 2 whenever block_n resolves to 256 (env CRAB_GEMM_2CTA gives the initial value). */
int crab_set_gemm_2cta(int mode);

/* ------------------------------------------------------------------------------------------------------------------
 * Norms
 * replaces: nn.LayerNorm in CLIP / BEATs / Q-Former / projectors (models/multimodal_encoder.py:99,128;
 *           models/beats/backbone.py:118,262,273; models/Qformer.py:66,284,371) and LlamaRMSNorm
 *           (models/modeling_llama.py:103-117: fp32 statistics, weight * x.to(bf16)).
 * x, y: bf16 [rows, ld]; gamma/beta fp32 [cols]; cols % 8 == 0.
 * ---------------------------------------------------------------------------------------------------------------- */
int crab_layernorm(const void* x, int ldx, const float* gamma, const float* beta, void* y, int ldy, int rows,
                   int cols, float eps, void* stream);
int crab_rmsnorm(const void* x, int ldx, const float* gamma, void* y, int ldy, int rows, int cols, float eps,
                 void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Rotary embedding + KV-cache append
 * replaces: LlamaRotaryEmbedding + apply_rotary_pos_emb (models/modeling_llama.py:123-156, 204-236) and HF
 *           DynamicCache.update (torch.cat growth) with a static cache [B, KVH, ctx_max, head_dim].
 * crab_rope_table: cos_sin fp32 [max_pos, head_dim] = [cos(head_dim/2) | sin(head_dim/2)] per position.
 * crab_rope_kv_append: qkv bf16 [B*S, ldq] rows = [q (H*hd) | k (KVH*hd) | v (KVH*hd)]; q rotated in place, rotated k
 *           and v written to the caches at position past+s; `past` is read from *past_dev when non-NULL (CUDA graphs).
 * ---------------------------------------------------------------------------------------------------------------- */
int crab_rope_table(float* cos_sin, int max_pos, int head_dim, double theta, void* stream);
int crab_rope_kv_append(void* qkv, int ldq, const float* cos_sin, void* k_cache, void* v_cache, int B, int S, int H,
                        int KVH, int head_dim, int ctx_max, const int* past_dev, int past_host, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Attention
 * crab_flash_attn replaces: eager softmax(QK^T)V in CLIPAttention (transformers clip), BEATs MultiheadAttention
 *           incl. the gated relative-position bias (models/beats/backbone.py:623-671), BertSelfAttention self/cross
 *           (models/Qformer.py:207-266) and causal LlamaAttention / Qwen2Attention prefill
 *           (models/modeling_llama.py:405-450, models/qwen/modeling_qwen2.py:190-199 repeat_kv).
 *           All strides are in elements; head_dim in {64,128}; bias = gate[B,H,Sq] * bias_table[H,Sq,Sk] (fp32) or NULL.
 *           head_dim-128 problems with >= 128 queries, no bias and TMA-describable strides run on the tcgen05 / TMEM kernel,
 *           which fetches whole 64-key tiles: rows up to the next multiple of 64 past Sk are read but never contribute (their
 *           scores are masked before the exponential and the kernel zeroes the V rows past Sk in shared memory, so even NaN bit
 *           patterns in an uninitialised cache are harmless); rows past the end of the tensor are zero-filled by TMA.
 *           With sk_dev the number of keys is read on the device: the GQA decode step calls this with the G query heads of
 *           a kv group as the Sq = G "rows" of one problem (q_rs = head_dim), so grouped-query decode runs on tensor cores.
 * crab_attn_decode replaces: the same attention at q_len == 1 over the KV cache (decode step), split over the
 *           context into `nsplit` partitions (workspace from crab_attn_decode_workspace_bytes when nsplit > 1).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct crab_attn_args {
  const void* q; const void* k; const void* v; void* o; /* bf16 */
  int64_t q_bs, q_rs, q_hs;                              /* batch / row / head strides */
  int64_t k_bs, k_rs, k_hs;
  int64_t v_bs, v_rs, v_hs;
  int64_t o_bs, o_rs, o_hs;
  int32_t B, H, KVH, Sq, Sk, head_dim;
  float scale;
  int32_t causal;                                        /* key j visible to query i iff j <= i + (Sk - Sq) */
  const float* gate;
  const float* bias_table;
  const int* sk_dev;                                     /* NULL, or device int overriding Sk (decode inside a CUDA graph) */
} crab_attn_args;
int crab_flash_attn(const crab_attn_args* args, void* stream);
int crab_attn_decode_workspace_bytes(int B, int H, int head_dim, int nsplit, int64_t* bytes);
int crab_attn_decode(const void* q, int ldq, const void* k_cache, const void* v_cache, void* o, int ldo,
                     float* workspace, int B, int H, int KVH, int head_dim, int ctx_max, int nsplit,
                     const int* len_dev, int len_host, float scale, void* stream);

/* crab_attn_decode_fused replaces, in ONE launch per layer of the decode step: apply_rotary_pos_emb on the new q / k
 *           (models/modeling_llama.py:204-236), the DynamicCache append, the q_len == 1 attention, and — when lora_ra is
 *           given — the o_proj hyper-LoRA router / A projections with the fp32 router softmax
 *           (peft_hyper/tuners/lora.py:344-350), whose 24 z values land in lora_z (the o_proj GEMM's K-extension).
 *           qkv: RAW [B, ldq] rows [q | k | v] from the qkv projection; cos_sin from crab_rope_table; *past_dev = position
 *           of the new token (valid keys = past + 1).  lora_ws: B*KVH*11 floats (B*H*11 when nsplit > 1: the pre-pass then runs in
 *           the split-KV combine launch, one block per head); lora_counters: B ints, zero on entry (left zero). */
typedef struct crab_decode_fused_args {
  const void* qkv; int32_t ldq;
  const float* cos_sin;
  void* k_cache; void* v_cache;           /* bf16 [B, KVH, ctx_max, head_dim] */
  void* o; int32_t ldo;                   /* bf16 [B, ldo] */
  float* workspace;                       /* split-KV partials (nsplit > 1) or NULL */
  int32_t B, H, KVH, head_dim, ctx_max, nsplit;
  const int* past_dev;
  float scale;
  const void* lora_ra; int32_t ld_ra;     /* bf16 [11, ld_ra] or NULL */
  void* lora_z; int32_t ld_z;             /* bf16 [B, ld_z] */
  float lora_scale;
  float* lora_ws; int* lora_counters;
  /* != 0 (head_dim 128, grouped-query models): RoPE + append as one launch, then the G query heads of a kv group as the G rows of a
   * tensor-core flash-attention problem, split over nsplit key ranges, then the combine launch (which also does the o_proj pre-pass:
   * lora_ra needs nsplit > 1 here).  workspace: B * H * nsplit * (head_dim + 2) floats. */
  int32_t gqa_tensor_cores; int32_t reserved0;
} crab_decode_fused_args;
int crab_attn_decode_fused(const crab_decode_fused_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Gathers, casts, patchify, CLIP embeddings, BEATs helpers, arg-max
 * crab_gather_rows replaces: embed_tokens lookups and the torch.cat / left-pad splice of
 *           prepare_multimodal_inputs (models/unified_arch.py:262-373): dst[dst_rows[i]] = src[src_rows[i]]
 *           (NULL index array = identity).
 * crab_patchify replaces: the im2col half of the stride==kernel patch-embed convs (CLIP 14x14, BEATs 16x16,
 *           models/beats/BEATs.py:148-151); rows are patches in (h', w') order, columns in (c, kh, kw) order.
 * crab_clip_embed_ln replaces: CLIPVisionEmbeddings cls/pos add + pre_layrnorm.
 * crab_beats_gate replaces: the grep_linear gate (models/beats/backbone.py:650-662).
 * crab_beats_group_pack / crab_beats_posconv_finish bracket the grouped pos-conv (backbone.py:33-46,114-116), which
 *           runs as one Toeplitz GEMM per conv group through crab_gemm_bf16.
 * crab_argmax replaces: HF greedy argmax over logits[:, -1].
 * ---------------------------------------------------------------------------------------------------------------- */
int crab_gather_rows(const void* src, int lds, const int64_t* src_rows, void* dst, int ldd, const int64_t* dst_rows,
                     int n, int cols, void* stream);
int crab_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
int crab_patchify(const float* images, void* out, int ld_out, int n_img, int C, int H, int W, int patch, void* stream);
int crab_clip_embed_ln(const void* patch_emb, const float* cls, const float* pos, const float* gamma,
                       const float* beta, void* out, int n_img, int tokens, int D, float eps, void* stream);
int crab_beats_gate(const void* q, int ldq, const float* grep_w, const float* grep_b, const float* grep_a, float* gate,
                    int B, int T, int H, void* stream);
int crab_beats_group_pack(const void* x, void* xg, int B, int T, int C, int G, void* stream);
int crab_beats_posconv_finish(const void* x, const void* conv_g, const float* bias, void* y, int B, int T, int C,
                              int G, void* stream);
int crab_argmax(const float* logits, int ld, int rows, int V, int64_t* out, void* stream);
int crab_add_scalar_i32(int* p, int v, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Token choice and loss on fp32 logits
 * crab_sample_top_k_top_p replaces: HF generate(do_sample=True)'s logits processors + multinomial draw for the checkpoint's
 *           generation_config (LLaMA-2-chat: temperature 0.6, top_p 0.9, top_k 50; SURVEY.md 8c shim 7): logits / temperature ->
 *           TopKLogitsWarper -> TopPLogitsWarper (keep the smallest set of most probable tokens whose mass reaches top_p) -> draw by
 *           inverse CDF over the kept tokens in index order with the caller's uniform u[row] in [0, 1).  top_k 0 = off.
 * crab_cross_entropy replaces: the CrossEntropyLoss of UnifiedForCausalLM.forward(labels=...) (models/unified_llama.py:150-160,
 *           HF LlamaForCausalLM loss): loss[row] = logsumexp(logits[row]) - logits[row, labels[row]], 0 for labels < 0 (-100).
 * ---------------------------------------------------------------------------------------------------------------- */
int crab_sample_top_k_top_p(const float* logits, int ld, int rows, int V, float temperature, int top_k, float top_p,
                            const float* u, int64_t* out, void* stream);
int crab_cross_entropy(const float* logits, int ld, int rows, int V, const int64_t* labels, float* loss, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Decode-step (M <= 32 rows) kernels
 * crab_gemm_skinny_bf16 replaces: the same nn.Linear / hyper-LoRA linears at q_len == 1 (HF generate's per-token
 *           forward, models/unified_llama.py:125-127): swap-AB tcgen05 weight streaming; K is split across the CTAs
 *           of a thread-block cluster and reduced through distributed shared memory (deterministic, no workspace).
 *           crab_gemm_skinny_plan reports the K-split the library would pick for (N, K).
 * crab_pack_skinny_weight: streaming layout for decode weights — block (tile, kb) = the 128x64 tile, pre-swizzled for
 *           the tensor core, 16 KB contiguous; `out` needs crab_skinny_packed_bytes(N, K) bytes, 128-byte aligned.
 *           swiglu_interleave=1 converts the prefill layout ([64 gate | 64 up] row groups) to interleaved rows
 *           (2i = gate_i, 2i+1 = up_i), which CRAB_ACT_SWIGLU requires here.
 * crab_row_norm_loraz replaces: LlamaRMSNorm + lora_route / lora_A projections + fp32 router softmax for up to three
 *           linears sharing one input row (peft_hyper/tuners/lora.py:344-350).  gamma/y NULL = no norm (LoRA only).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct crab_skinny_args {
  const void* X;        /* bf16 [M, ldx], M <= 32 */
  const void* W;        /* bf16 [N, ldw] row-major (read through TMA), or NULL when W_packed is given */
  const void* W_packed; /* same weight pre-packed by crab_pack_skinny_weight, or NULL */
  void* C;              /* [M, ldc] bf16 or fp32 */
  const float* bias;    /* fp32 [N] or NULL */
  const void* residual; /* bf16 [M, ldr] or NULL (may alias C) */
  int32_t M, N, K, ldx, ldw, ldc, ldr;
  int32_t act;          /* CRAB_ACT_NONE or CRAB_ACT_SWIGLU (needs W_packed with swiglu_interleave) */
  int32_t out_dtype;
  int32_t splits;       /* K-split = cluster size, 1..8; 0 = auto */
  /* ---- optional (all zero = plain GEMM): the decode step's fused form of LlamaRMSNorm + hyper-LoRA Linear.forward ----
   * Z / Kext: K-extension activations (z' columns, bf16 [M, ldz]) multiplied against columns K .. K+Kext of the packed weight
   *      ([W' | B_0 B_1 B_2] per wrapped linear); K must be a multiple of 64.
   * norm: RMSNorm as an epilogue scale — the packed W' already carries gamma (W * diag(gamma)), C = rstd[b] * acc (+bias, ...).
   * stats_linears (0..3) + stats_packed (crab_decode_chain_stats_bytes layout): ONE extra cluster of the launch computes
   *      rstd[b] = rsqrt(mean(X[b]^2) + eps) and, per wrapped linear, t = X . [R;A]^T, z' = lora_scale * softmax(rstd * t[0:3])_i *
   *      t[3+j] -> Z[b, 24*linear + 8*i + j]; the K-extension k-blocks and the epilogues of the other clusters wait for it.
   *      replaces the separate crab_row_norm_loraz launch (peft_hyper/tuners/lora.py:344-350, models/modeling_llama.py:103-117).
   * rstd: 32 floats of scratch; flags: 64 ints, 128-byte aligned, zero on entry (left zero).
   * stats_clusters / stats_scratch: the statistics item is latency-bound (x tiles from L2 behind everybody's weight stream), so
   *      its K range may be shared by up to 8 clusters whose partial sums meet in stats_scratch (8 x 36 x 32 floats; any launch
   *      may reuse the same buffer) and are added in cluster order by the last one to arrive (deterministic).  0 = the library's
   *      choice (several when stats_scratch is given, else one).
   * flags_clear: optional.  NULL: the launch zeroes its own `flags` before it exits (one atomic per CTA, ~1 us at the tail).
   *      Otherwise: the `flags` slot (first 2 ints) of the launch that ran BEFORE this one on the stream — it is zeroed here and
   *      this launch leaves its own slot set for the next launch to zero: give consecutive launches distinct slots in a ring. */
  const void* Z; int32_t ldz; int32_t Kext;
  const void* stats_packed; int32_t stats_linears; int32_t norm;
  float eps; float lora_scale;
  float* rstd; int* flags;
  /* prefetch / prefetch_bytes: optional — a global span (the NEXT launch's packed weights) that the CTAs pull into L2 once their own
   * loads are issued, so HBM does not idle between two dependent launches; no effect on results. */
  const void* prefetch; int64_t prefetch_bytes;
  float* stats_scratch; int* flags_clear; int32_t stats_clusters; int32_t reserved0;
} crab_skinny_args;
int crab_gemm_skinny_plan(int N, int K, int* splits, int64_t* workspace_bytes, int* n_counters);
int crab_skinny_packed_bytes(int N, int K, int64_t* bytes);
int crab_pack_skinny_weight(const void* W, int N, int K, int ldw, void* out, int swiglu_interleave, void* stream);
int crab_gemm_skinny_bf16(const crab_skinny_args* args, void* stream);
int crab_row_norm_loraz(const void* x, int ldx, const float* gamma, float eps, void* y, int ldy, const void* ra,
                        int ldra, int groups, void* z, int ldz, float scale, int rows, int cols, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Decode-step GEMM chain (csrc/decode_chain.cu): up to four DEPENDENT M <= 32 linears of one decoder layer in one
 * persistent launch — o_proj(+residual) -> [RMSNorm] gate/up + SwiGLU -> down_proj(+residual) -> [RMSNorm] qkv of the next
 * layer (or the final norm + lm_head) — with the weight stream running across the phase boundaries.
 * replaces: per decode step and layer, LlamaDecoderLayer.forward minus the attention core (models/modeling_llama.py:765-837:
 *           two LlamaRMSNorm :103-117, o_proj :448-460, LlamaMLP :239-271, next layer's q/k/v projections :368-392; Qwen2:
 *           models/qwen/modeling_qwen2.py:175-187, 234-237 with q/k/v bias), each linear being the hyper-LoRA Linear.forward of
 *           peft_hyper/tuners/lora.py:338-369 (router softmax in fp32, r = 8 x 3 experts), and for the last layer
 *           LlamaModel.norm + lm_head (models/modeling_llama.py:1119, 1254-1261).
 * Per phase: C[M, N] = epilogue( s[b] * ( X[M, K] . W'[N, K]^T + Z[M, Kext] . Bcat[N, Kext]^T ) ) with
 *   - W_packed: crab_pack_skinny_weight layout of the row-major [N, K + Kext] matrix [W' | Bcat] (SwiGLU: swiglu_interleave);
 *     for a normalised phase W' = W * diag(gamma) (RMSNorm scale folded at load time) and s[b] = rstd[b], else s = 1;
 *   - norm / stats_linears: the phase starts with a statistics item that computes rstd[b] = rsqrt(mean(X[b]^2) + eps) and, for
 *     each of the `stats_linears` linears sharing X, t = X . [R; A]^T (11 dots), z' = lora_scale * softmax(s * t[0:3])_i *
 *     t[3 + j] -> Z[b, 24 * linear + 8 * i + j] (bf16; with norm the rows of stats_packed carry gamma as well).  stats_packed:
 *     [K / 64] blocks of 40 x 64 bf16, rows 11 * linear + j, pre-swizzled like the weights (crab_decode_chain_stats_bytes);
 *     stats_linears == 0 with Kext > 0 means Z was filled by an earlier kernel (the fused decode attention writes o_proj's z);
 *   - act CRAB_ACT_SWIGLU (N_out = N / 2), bias fp32 [N], residual bf16 [M, ldr] (may alias C), bf16 or fp32 output.
 * Phase i + 1 may read what phase i wrote (X / residual); phases of one launch must not write a buffer an earlier phase still
 * reads.  counters: 288 ints (nine 128-byte slots), 128-byte aligned, zero on entry (the kernel leaves them zero).  cluster: K-split = thread-block-cluster size
 * (1, 2, 4, 8; 0 = 4).  The launch is cooperative in effect: grid = (max co-resident clusters) x cluster, so it must not be
 * launched concurrently with other work on the same device.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct crab_chain_phase {
  const void* X; int32_t ldx; int32_t K;          /* bf16 [M, ldx], K % 64 == 0 */
  const void* Z; int32_t ldz; int32_t Kext;       /* bf16 [M, ldz] z' columns (written by the stats item when stats_linears > 0) */
  const void* W_packed;
  const void* stats_packed; int32_t stats_linears; int32_t norm;
  float eps; float lora_scale;
  float* rstd;                                    /* 32 floats of scratch (norm phases) */
  void* C; int32_t ldc; int32_t N; int32_t out_dtype; int32_t act;
  const float* bias;
  const void* residual; int32_t ldr;
} crab_chain_phase;
typedef struct crab_chain_args {
  crab_chain_phase phase[4];
  int32_t n_phases, M, cluster, max_clusters;      /* max_clusters 0 = all that fit */
  int* counters;
} crab_chain_args;
int crab_decode_chain(const crab_chain_args* args, void* stream);
int crab_decode_chain_stats_bytes(int K, int64_t* bytes);
int crab_decode_chain_max_clusters(int cluster, int* n);

/* ------------------------------------------------------------------------------------------------------------------
 * Front-end preprocessing (the step immediately before the path)
 * crab_kaldi_fbank replaces: dataset/audio_processor.py:29-41 `preprocess` = torchaudio.compliance.kaldi.fbank(
 *           waveform * 2**15, num_mel_bins=128, sample_frequency=16000, frame_length=25, frame_shift=10) [defaults:
 *           dither 0, remove_dc_offset, preemphasis 0.97, povey window, round_to_power_of_two, snip_edges, use_power,
 *           use_log_fbank] followed by (fbank - mean) / (2 * std).  wave: fp32 [n_seg, wave_stride] (n_samples valid);
 *           window [400]; twiddle [256][2] = (cos, -sin)(2 pi k / 512); the mel filters are passed as sparse rows (mel_start[n_mel], mel_off[n_mel+1], mel_w[nnz]) built
 *           by the host from the Kaldi formula; out: fp32 [n_seg, n_frames, n_mel], n_frames from crab_fbank_num_frames.
 *           `std2` is the full divisor (2 * 6.55582 in the reference).
 * crab_patchify_u8 replaces: CLIPImageProcessor rescale + normalize (dataset/quick_start_dataset.py:303-315 calls
 *           `video_processor.preprocess(frames)` on 224x224 decoded frames, for which resize / centre-crop are
 *           identities) fused with crab_patchify: uint8 [n, H, W, 3] -> bf16 patch rows.
 * crab_normalize_u8: same arithmetic to the reference's fp32 `pixel_values` [n, 3, H, W].
 * ---------------------------------------------------------------------------------------------------------------- */
int crab_kaldi_fbank(const float* wave, int64_t wave_stride, int n_seg, int n_samples, const float* window,
                     const float* twiddle, const int* mel_start, const int* mel_off, const float* mel_w, int n_mel, float in_scale,
                     float mean, float std2, float* out, void* stream);
int crab_fbank_num_frames(int n_samples, int* n_frames);
int crab_patchify_u8(const void* images_hwc, void* out, int ld_out, int n_img, int H, int W, int patch,
                     const float* mean3, const float* std3, float rescale, void* stream);
int crab_normalize_u8(const void* images_hwc, float* out_nchw, int n_img, int H, int W, const float* mean3,
                      const float* std3, float rescale, void* stream);
/* crab_resample_u8 replaces: the processor's shortest-edge resize, i.e. PIL.Image.resize(size, BICUBIC) of the pinned
 *           transformers (Pillow's ImagingResample{Horizontal,Vertical}_8bpc), one separable pass per call on uint8
 *           [n, H, W, 3]: axis 1 resamples columns (in_h == out_h), axis 0 rows (in_w == out_w).  bounds[2*i] / bounds[2*i+1]
 *           = first tap / tap count and kk[i*ksize ..] = 22-bit fixed-point coefficients of output index i (built on the host
 *           from Pillow's formula); output element j reads coefficient row out0 + j, which folds the centre crop in. */
int crab_resample_u8(const void* in_hwc, void* out_hwc, int n_img, int in_h, int in_w, int out_h, int out_w, int axis,
                     const int* bounds, const int* kk, int ksize, int out0, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Segmentation head helpers (SegModule / MaskDecoderMultiScale, models/multimodal_encoder.py:268-543,
 * 891-1445).  Feature maps are token-major (one row per pixel); convolutions and MLPs go through crab_gemm_bf16.
 * crab_small_attn replaces: Attention.forward (:1368-1393) and nn.MultiheadAttention of the query generator: heads of 16 or
 *           32 channels at column h * head_dim of q / k / v / o (bf16), softmax(q k^T * scale) v in fp32.
 * crab_elementwise: op 0 out = a + b (b_rows 1 = row broadcast), 1 ReLU, 2 GELU (erf), 3 (sigmoid(gate[r]) + 1) * a
 *           (predict_masks :1110-1112); bf16 in / out.
 * crab_row_mean_f32: mean over the first `cols` fp32 columns of every row (class mean of the previous masks).
 * crab_im2col3x3: bf16 [h*w, C] -> [h*w, 9*C], columns (ky, kx, c), zero padding (image_feature_neck's 3x3 conv).
 * crab_bilinear_f32: F.interpolate(mode="bilinear", align_corners=False) on token-major fp32 maps;
 *           out = beta * out + alpha * interp; nchw_out = 1 writes [C, hout, wout] (the final masks).
 * ---------------------------------------------------------------------------------------------------------------- */
int crab_small_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int Nq, int Nk,
                    int heads, int head_dim, float scale, void* stream);
int crab_elementwise(const void* a, int lda, const void* b, int ldb, int b_rows, const float* gate, void* out, int ldo, int rows,
                     int cols, int op, void* stream);
int crab_row_mean_f32(const float* x, int ldx, int rows, int cols, float* out, void* stream);
int crab_im2col3x3(const void* in, int ldi, void* out, int h, int w, int C, void* stream);
int crab_bilinear_f32(const float* in, int ldi, int hin, int win, float* out, int ldo, int hout, int wout, int C, float alpha,
                      float beta, int nchw_out, void* stream);

/* Diagnostics only (tools/trace_skinny.py): arm (buf != NULL) or disarm a device buffer of nslots x ctas_per_slot x 16 uint64
 * globaltimer stamps; each following crab_gemm_skinny_bf16 / crab_attn_decode_fused launch takes the next slot and its CTAs
 * record when they started, passed the dependency wait, saw their first operand, finished their MMAs / stream and left.
 * Process-wide, not thread-safe, costs ~1 us per CTA while armed. */
int crab_debug_trace(void* buf, int nslots, int ctas_per_slot);

#ifdef __cplusplus
}
#endif
#endif /* CRAB_B200_H_ */
