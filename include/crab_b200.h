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

#ifdef __cplusplus
}
#endif
#endif /* CRAB_B200_H_ */
