// Weight-streaming GEMM for the decode step (M <= 32 token rows):  C[M,N] = epilogue(X[M,K] . W[N,K]^T)
//
// At M <= 32 the op is HBM-bound (2 FLOP per weight byte): the job is to stream W once at full bandwidth.  The
// operands are swapped so the weights take the 128-row A slot of the tensor core and the batch takes the N slot:
//     D[128 weight rows, 32 batch columns] += W_tile[128, 64] . X_tile[32, 64]^T      (tcgen05.mma M=128 N=32 K=16)
// so no weight byte is fetched twice and the X tile (4 KB / k-block) is the only redundant traffic.
// N/128 tiles cannot fill 148 SMs for the 4096-wide projections, so K is split across CTAs (split-K); every CTA parks
// its fp32 partial tile in a workspace and the LAST CTA to finish a tile (atomic ticket) reduces the partials in
// fixed split order — deterministic, unlike fp32 atomics — and applies the epilogue (bias, residual, SwiGLU, cast).
// Two CTAs are resident per SM (5 x 20 KB stages each) so one CTA's prologue / fix-up overlaps the other's stream.
//
//   warp 0 : TMA producer (W 128x64 + X 32x64 per stage, SWIZZLE_128B)     warp 1 : MMA issuer + TMEM owner
//   warps 2-5 : TMEM -> workspace, ticket, fix-up epilogue
#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int SK_BM = 128;   // weight rows per tile
static constexpr int SK_MB = 32;    // batch columns (UMMA N)
static constexpr int SK_BK = 64;
static constexpr int SK_STAGES = 5;
static constexpr int SK_W_BYTES = SK_BM * SK_BK * 2;
static constexpr int SK_X_BYTES = SK_MB * SK_BK * 2;
static constexpr int SK_STAGE_BYTES = SK_W_BYTES + SK_X_BYTES;
static constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + 1024 + 128;
static constexpr int SK_THREADS = 192;

struct SkinnyParams {
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  float* ws;
  int* counters;
  int M, N, K, ldc, ldr;
  int act, out_dtype, splits;
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(SK_THREADS, 2)
gemm_skinny_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                           const SkinnyParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_last;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + SK_STAGES * SK_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SK_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * SK_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * SK_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int kb_total = (p.K + SK_BK - 1) / SK_BK;
  const int kb0 = (int)((long long)split * kb_total / p.splits);
  const int kb1 = (int)((long long)(split + 1) * kb_total / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < SK_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  pdl_trigger();  // the next kernel of the chain may become resident now (it waits for us before touching data)
  if (warp == 0) {
    if (lane == 0) {
      // Weights never depend on an earlier kernel: fill the whole ring with W tiles BEFORE waiting for the producer
      // of X, so the weight stream is already in flight while the previous kernel drains.
      const int npre = min(kb1 - kb0, SK_STAGES);
      for (int i = 0; i < npre; ++i) {
        mbar_arrive_expect_tx(full_bar(i), SK_STAGE_BYTES);
        tma_load_2d_hint(smem_base + i * SK_STAGE_BYTES, &tmap_w, full_bar(i), (kb0 + i) * SK_BK, tile * SK_BM, kEvictFirst);
      }
      pdl_wait();
      for (int i = 0; i < npre; ++i)
        tma_load_2d_hint(smem_base + i * SK_STAGE_BYTES + SK_W_BYTES, &tmap_x, full_bar(i), (kb0 + i) * SK_BK, 0, kEvictLast);
      uint32_t stage = 0, phase = 1;  // ring position after the prefill above (npre == SK_STAGES wraps to stage 0)
      if (npre < SK_STAGES) { stage = npre; phase = 0; }
      for (int kb = kb0 + npre; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(stage), SK_STAGE_BYTES);
        const uint32_t sw = smem_base + stage * SK_STAGE_BYTES;
        tma_load_2d_hint(sw, &tmap_w, full_bar(stage), kb * SK_BK, tile * SK_BM, kEvictFirst);  // weights: read once
        tma_load_2d_hint(sw + SK_W_BYTES, &tmap_x, full_bar(stage), kb * SK_BK, 0, kEvictLast);  // X: shared by all CTAs
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(SK_BM, SK_MB);
      uint32_t stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sw = smem_base + stage * SK_STAGE_BYTES;
        const uint64_t da = make_sdesc_sw128(sw);
        const uint64_t db = make_sdesc_sw128(sw + SK_W_BYTES);
#pragma unroll
        for (int k = 0; k < SK_BK / 16; ++k) umma_bf16_ss(tmem_base, da + 2u * k, db + 2u * k, idesc, (kb > kb0) | (k > 0));
        umma_commit(empty_bar(stage));
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    // ---- park the partial tile: ws[(tile*splits + split)][row 0..127][b 0..31] ----
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    pdl_wait();  // before the first write to the shared workspace / read of residual
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16), r);
    tmem_ld_wait();
    float* wrow = p.ws + (((size_t)tile * p.splits + split) * SK_BM + row) * SK_MB;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      __stcg(reinterpret_cast<float4*>(wrow) + g, make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                                              __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])));
    __threadfence();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 2 && lane == 0) {
      const int old = atomicAdd(p.counters + tile, 1);
      s_last = (old == p.splits - 1);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (s_last) {
      __threadfence();
      const int tt = (warp - 2) * 32 + lane;
      const float* wt = p.ws + (size_t)tile * p.splits * SK_BM * SK_MB;
      if (p.act == CRAB_ACT_SWIGLU) {
        // tile rows = [64 gate | 64 up]  ->  64 output columns
        const int n = tile * 64 + tt;
        if (tt < 64 && n < (p.N >> 1)) {
          float g[32], u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { g[j] = 0.f; u[j] = 0.f; }
          for (int s = 0; s < p.splits; ++s) {
            const float* pg = wt + ((size_t)s * SK_BM + tt) * SK_MB;
            const float* pu = wt + ((size_t)s * SK_BM + tt + 64) * SK_MB;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a = ldcg4(pg + 4 * q), b = ldcg4(pu + 4 * q);
              g[4 * q] += a.x; g[4 * q + 1] += a.y; g[4 * q + 2] += a.z; g[4 * q + 3] += a.w;
              u[4 * q] += b.x; u[4 * q + 1] += b.y; u[4 * q + 2] += b.z; u[4 * q + 3] += b.w;
            }
          }
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C);
#pragma unroll
          for (int b = 0; b < 32; ++b)
            if (b < p.M) c[(size_t)b * p.ldc + n] = __float2bfloat16_rn(g[b] / (1.0f + __expf(-g[b])) * u[b]);
        }
      } else {
        const int n = tile * SK_BM + tt;
        if (n < p.N) {
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = 0.f;
          for (int s = 0; s < p.splits; ++s) {
            const float* pa = wt + ((size_t)s * SK_BM + tt) * SK_MB;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 v = ldcg4(pa + 4 * q);
              a[4 * q] += v.x; a[4 * q + 1] += v.y; a[4 * q + 2] += v.z; a[4 * q + 3] += v.w;
            }
          }
          const float bias = p.bias ? p.bias[n] : 0.f;
#pragma unroll
          for (int b = 0; b < 32; ++b) {
            if (b < p.M) {
              float v = a[b] + bias;
              if (p.residual) v += __bfloat162float(p.residual[(size_t)b * p.ldr + n]);
              if (p.out_dtype == CRAB_BF16) reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)b * p.ldc + n] = __float2bfloat16_rn(v);
              else reinterpret_cast<float*>(p.C)[(size_t)b * p.ldc + n] = v;
            }
          }
        }
      }
      if (warp == 2 && lane == 0) p.counters[tile] = 0;  // ready for the next launch (stream order)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 32); }
}

int choose_splits(int N, int K) {
  const int tiles = (N + SK_BM - 1) / SK_BM;
  const int kb = (K + SK_BK - 1) / SK_BK;
  const int target = 2 * sm_count();
  int s = (target + tiles / 2) / tiles;
  if (s < 1) s = 1;
  if (s > kb / 4) s = kb / 4 > 0 ? kb / 4 : 1;
  if (s > 16) s = 16;
  return s;
}

}  // namespace crab

using namespace crab;

extern "C" int crab_gemm_skinny_plan(int N, int K, int* splits, int64_t* workspace_bytes, int* n_counters) {
  CRAB_REQUIRE(N > 0 && K > 0 && splits && workspace_bytes && n_counters, "crab_gemm_skinny_plan: bad args");
  const int tiles = (N + SK_BM - 1) / SK_BM;
  *splits = choose_splits(N, K);
  *workspace_bytes = (int64_t)tiles * (*splits) * SK_BM * SK_MB * 4;
  *n_counters = tiles;
  return CRAB_OK;
}

extern "C" int crab_gemm_skinny_bf16(const crab_skinny_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CRAB_REQUIRE(a && a->X && a->W && a->C && a->workspace && a->counters, "crab_gemm_skinny_bf16: null pointer");
  CRAB_REQUIRE(a->M > 0 && a->M <= SK_MB, "crab_gemm_skinny_bf16: M must be in 1..32 (got %d)", a->M);
  CRAB_REQUIRE(a->N > 0 && a->K > 0 && a->ldx % 8 == 0 && a->ldw % 8 == 0 && a->ldx >= a->K && a->ldw >= a->K,
               "crab_gemm_skinny_bf16: bad shape/strides N=%d K=%d ldx=%d ldw=%d", a->N, a->K, a->ldx, a->ldw);
  CRAB_REQUIRE(((uintptr_t)a->X % 16 == 0) && ((uintptr_t)a->W % 16 == 0), "crab_gemm_skinny_bf16: X/W must be 16-byte aligned");
  CRAB_REQUIRE(a->act == CRAB_ACT_NONE || a->act == CRAB_ACT_SWIGLU, "crab_gemm_skinny_bf16: act must be NONE or SWIGLU");
  if (a->act == CRAB_ACT_SWIGLU)
    CRAB_REQUIRE(a->N % 128 == 0 && !a->bias && !a->residual && a->out_dtype == CRAB_BF16, "crab_gemm_skinny_bf16: SWIGLU constraints");
  const int tiles = (a->N + SK_BM - 1) / SK_BM;
  const int kb = (a->K + SK_BK - 1) / SK_BK;
  int splits = a->splits > 0 ? a->splits : choose_splits(a->N, a->K);
  if (splits > kb) splits = kb;
  if (splits > 16) splits = 16;
  CRAB_REQUIRE(a->n_counters >= tiles, "crab_gemm_skinny_bf16: need %d counters (got %d)", tiles, a->n_counters);
  CRAB_REQUIRE(a->workspace_bytes >= (int64_t)tiles * splits * SK_BM * SK_MB * 4, "crab_gemm_skinny_bf16: workspace too small");
  static bool attr_set = false;
  if (!attr_set) {
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
    attr_set = true;
  }
  CUtensorMap tw, tx;
  int rc = encode_tmap_bf16_2d(&tw, a->W, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldw, SK_BM, SK_BK);
  if (rc != 0) return rc;
  rc = encode_tmap_bf16_2d(&tx, a->X, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->ldx, SK_MB, SK_BK);
  if (rc != 0) return rc;
  SkinnyParams p;
  p.C = a->C; p.bias = a->bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.ws = a->workspace; p.counters = a->counters;
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldc = a->ldc; p.ldr = a->ldr;
  p.act = a->act; p.out_dtype = a->out_dtype; p.splits = splits;
  CRAB_CHECK_CUDA(launch_pdl(gemm_skinny_tcgen05_kernel, dim3(tiles * splits), dim3(SK_THREADS), SK_SMEM, stream, tw, tx, p));
  return CRAB_OK;
}
