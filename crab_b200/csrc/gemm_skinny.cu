// Weight-streaming GEMM for the decode step (M <= 32 token rows):  C[M,N] = epilogue(X[M,K] . W[N,K]^T)
//
// At M <= 32 the op is HBM-bound (2 FLOP per weight byte): the job is to stream W once at full bandwidth.  The
// operands are swapped so the weights take the 128-row A slot of the tensor core and the batch takes the N slot:
//     D[128 weight rows, 32 batch columns] += W_tile[128, 64] . X_tile[32, 64]^T      (tcgen05.mma M=128 N=32 K=16)
// so no weight byte is fetched twice and the X tile (4 KB / k-block) is the only redundant traffic.
//
// Scheduling is persistent stream-K: the (tile, k-block) space is flattened and cut into one contiguous range per
// CTA (one CTA per SM), so every SM streams the same number of bytes whatever N and K are.  A CTA's range crosses
// at most a few tile boundaries; each piece ("segment") accumulates in TMEM (double-buffered so the epilogue of one
// segment overlaps the stream of the next), is parked as an fp32 partial in an L2-resident workspace, and the LAST
// contributor of a tile (atomic ticket) sums the partials in fixed order — deterministic, unlike fp32 atomics — and
// applies the epilogue (bias, residual, SwiGLU, cast).  The TMA ring (5 x 20 KB stages) never drains between segments.
// One CTA takes half an SM's shared memory on purpose: with PDL the NEXT kernel of the decode chain becomes resident
// on the same SM while this one is still streaming, fills its own ring with weight tiles, and only then waits for its
// producer — so launch latency, prologue and first-byte latency of every GEMM hide behind the previous kernel.
//
//   warp 0 : TMA producer (W 128x64 + X 32x64 per stage, SWIZZLE_128B)     warp 1 : MMA issuer + TMEM owner
//   warps 2-5 : TMEM -> workspace, ticket, fix-up epilogue
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int SK_BM = 128;   // weight rows per tile
static constexpr int SK_MB = 32;    // batch columns (UMMA N)
static constexpr int SK_BK = 64;
static constexpr int SK_STAGES = 5;   // 100 KB: leaves room for the NEXT kernel's CTA on the same SM (PDL overlap)
static constexpr int SK_W_BYTES = SK_BM * SK_BK * 2;
static constexpr int SK_X_BYTES = SK_MB * SK_BK * 2;
static constexpr int SK_STAGE_BYTES = SK_W_BYTES + SK_X_BYTES;
static constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + 1024 + 256;
static constexpr int SK_THREADS = 192;
static constexpr int SK_MAX_SLOTS = 16;  // max contributors (CTAs) per tile (bounds the fix-up loop)

struct SkinnyParams {
  void* C;
  const float* bias;
  const __nv_bfloat16* residual;
  float* ws;
  int* counters;
  int M, N, K, ldc, ldr;
  int act, out_dtype;
  const __nv_bfloat16* w_tiled;  // non-null: weights pre-packed as contiguous, pre-swizzled 16 KB (tile, k-block) blocks
  unsigned long long* trace;  // diagnostic: per-CTA timestamps of the last segment's epilogue (env CRAB_SK_TRACE=1)
  int debug;  // diagnostic bit mask (env CRAB_SK_DEBUG): 1 = no X loads, 2 = no MMA, 4 = no epilogue/fix-up (wrong results!)
  int tiles, kb_per_tile, max_segs;  // max_segs: workspace slots per CTA (segments a CTA range can touch)
  int total_kb;
  int q, rem;  // balanced partition of total_kb over the grid: CTA c owns [c*q + min(c,rem), ...) — no 64-bit divisions on device
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// flattened k-block range of CTA c: the first `rem` CTAs own q+1 blocks, the rest q  (q = T / G, rem = T % G)
__device__ __forceinline__ int sk_lo(int c, int q, int rem) { return c * q + min(c, rem); }
// the CTA whose range contains flattened k-block s
__device__ __forceinline__ int sk_owner(int s, int q, int rem) {
  const int big = rem * (q + 1);
  return s < big ? (int)((unsigned)s / (unsigned)(q + 1)) : rem + (int)((unsigned)(s - big) / (unsigned)q);
}

__global__ void __launch_bounds__(SK_THREADS, 2)
gemm_skinny_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                           const SkinnyParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_last[2];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + SK_STAGES * SK_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SK_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SK_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SK_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SK_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb_per_tile;
  const int lo = sk_lo(blockIdx.x, p.q, p.rem), hi = sk_lo(blockIdx.x + 1, p.q, p.rem);

  if (warp == 0 && lane == 0) {
    if (!p.w_tiled) tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < SK_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
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
      const int npre = min(SK_STAGES, hi - lo);
      int tile = lo / KB, kb = lo - tile * KB;  // advanced incrementally: no division in the issue loop
      const int tile0 = tile, kb0 = kb;
      for (int i = 0; i < npre; ++i) {
        const int f = lo + i;
        mbar_arrive_expect_tx(full_bar(i), (p.debug & 1) ? SK_W_BYTES : SK_STAGE_BYTES);
        if (p.w_tiled) bulk_load_1d_hint(smem_base + i * SK_STAGE_BYTES, p.w_tiled + (size_t)f * (SK_BM * SK_BK), SK_W_BYTES, full_bar(i), kEvictFirst);
        else tma_load_2d_hint(smem_base + i * SK_STAGE_BYTES, &tmap_w, full_bar(i), kb * SK_BK, tile * SK_BM, kEvictFirst);
        if (++kb == KB) { kb = 0; ++tile; }
      }
      pdl_wait();
      {
        int kx = kb0, tx_ = tile0;
        for (int i = 0; i < npre && !(p.debug & 1); ++i) {
          tma_load_2d_hint(smem_base + i * SK_STAGE_BYTES + SK_W_BYTES, &tmap_x, full_bar(i), kx * SK_BK, 0, kEvictLast);
          if (++kx == KB) { kx = 0; ++tx_; }
        }
      }
      uint32_t stage = 0, phase = 1;  // ring position after the prefill above (npre == SK_STAGES wraps to stage 0)
      for (int f = lo + npre; f < hi; ++f) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_arrive_expect_tx(full_bar(stage), (p.debug & 1) ? SK_W_BYTES : SK_STAGE_BYTES);
        const uint32_t sw = smem_base + stage * SK_STAGE_BYTES;
        if (p.w_tiled) bulk_load_1d_hint(sw, p.w_tiled + (size_t)f * (SK_BM * SK_BK), SK_W_BYTES, full_bar(stage), kEvictFirst);
        else tma_load_2d_hint(sw, &tmap_w, full_bar(stage), kb * SK_BK, tile * SK_BM, kEvictFirst);   // weights: read once
        if (!(p.debug & 1)) tma_load_2d_hint(sw + SK_W_BYTES, &tmap_x, full_bar(stage), kb * SK_BK, 0, kEvictLast);  // X: shared by all CTAs
        if (++kb == KB) { kb = 0; ++tile; }
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(SK_BM, SK_MB);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      int f = lo;
      int tile_end = (lo / KB + 1) * KB;
      while (f < hi) {
        const int seg_end = min(hi, tile_end);
        tile_end += KB;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * SK_MB;
        for (int g = f; g < seg_end; ++g) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sw = smem_base + stage * SK_STAGE_BYTES;
          const uint64_t da = make_sdesc_sw128(sw);
          const uint64_t db = make_sdesc_sw128(sw + SK_W_BYTES);
          if (p.debug & 2) {
            mbar_arrive(empty_bar(stage));
          } else {
#pragma unroll
            for (int k = 0; k < SK_BK / 16; ++k) umma_bf16_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, (g > f) | (k > 0));
            umma_commit(empty_bar(stage));
          }
          if (++stage == SK_STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.debug & 2) mbar_arrive(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        f = seg_end;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // Per segment: TMEM -> registers; a tile this CTA owns alone is finished straight from registers; a shared tile is
    // parked in the workspace and the last contributor (ticket) reduces it.  Publication follows the semaphore
    // pattern (stores; bar; ONE thread: fence + atomic; bar) instead of a fence in every thread.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int tt = (warp - 2) * 32 + lane;
    pdl_wait();  // before the first write to the shared workspace / read of residual
    uint32_t acc = 0, acc_phase = 0;
    int seg_parity = 0;
    int f = lo;
    const int first_tile = lo / KB;
    int tile = first_tile;
    while (f < hi) {
      const int t0 = tile * KB;
      const int seg_end = min(hi, t0 + KB);
      const int c_first = sk_owner(t0, p.q, p.rem), c_last = sk_owner(t0 + KB - 1, p.q, p.rem);
      const int n_contrib = c_last - c_first + 1;
      const bool tr = (p.trace != nullptr) && tt == 0 && seg_end == hi;
      unsigned long long* trp = p.trace ? p.trace + (size_t)blockIdx.x * 8 : nullptr;
      if (tr) trp[0] = globaltimer_ns();
      mbar_wait(tfull_bar(acc), acc_phase);
      if (tr) trp[1] = globaltimer_ns();
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * SK_MB, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));  // accumulator free: the MMA warp may start the next segment
      if (p.debug & 4) { acc ^= 1; if (acc == 0) acc_phase ^= 1; f = seg_end; ++tile; continue; }
      const bool swiglu = p.act == CRAB_ACT_SWIGLU;
      bool reduce_from_ws = false;
      if (n_contrib == 1 && !swiglu) {
        // ---- sole owner: finish from registers (thread = weight row n, 32 batch values) ----
        const int n = tile * SK_BM + row;
        if (n < p.N) {
          const float bias = p.bias ? p.bias[n] : 0.f;
          float resv[32];
#pragma unroll
          for (int b = 0; b < 32; ++b)  // all residual loads first: C may alias the residual (in-place x += ...)
            resv[b] = (p.residual != nullptr && b < p.M) ? __bfloat162float(p.residual[(size_t)b * p.ldr + n]) : 0.f;
#pragma unroll
          for (int b = 0; b < 32; ++b) {
            if (b < p.M) {
              const float v = __uint_as_float(r[b]) + bias + resv[b];
              if (p.out_dtype == CRAB_BF16) reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)b * p.ldc + n] = __float2bfloat16_rn(v);
              else reinterpret_cast<float*>(p.C)[(size_t)b * p.ldc + n] = v;
            }
          }
        }
      } else {
        // ---- park the partial tile: ws[cta][segment index within the CTA][row 0..127][b 0..31] ----
        const int my_seg = tile - first_tile;
        float* wrow = p.ws + (((size_t)blockIdx.x * p.max_segs + my_seg) * SK_BM + row) * SK_MB;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (!(p.debug & 32)) __stcg(reinterpret_cast<float4*>(wrow) + g, make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                                                  __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])));
        if (tr) trp[2] = globaltimer_ns();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tr) trp[3] = globaltimer_ns();
        if (tt == 0) {
          int last = 1;
          if (p.debug & 16) last = ((int)blockIdx.x == c_last);
          else if (n_contrib > 1) {
            __threadfence();  // release: cumulative over the CTA's partial stores ordered by the barrier above
            last = (atomicAdd(p.counters + tile, 1) == n_contrib - 1);
            if (last) { __threadfence(); p.counters[tile] = 0; }  // acquire; counter ready for the next launch
          }
          s_last[seg_parity] = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        reduce_from_ws = s_last[seg_parity] != 0;
        seg_parity ^= 1;
        if (tr) trp[4] = globaltimer_ns();
      }
      if (reduce_from_ws && !(p.debug & 8)) {
        // contributor s is CTA c_first + s; its partial for this tile sits at segment index tile - first_tile(cta)
        const int n_valid = swiglu ? ((tt < 64 && tile * 64 + tt < (p.N >> 1)) ? 1 : 0) : (tile * SK_BM + tt < p.N ? 1 : 0);
        if (n_valid) {
          float a[32], u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = 0.f; u[j] = 0.f; }
          for (int s = 0; s < n_contrib; ++s) {
            const int cc = c_first + s;
            // only the first contributor can start before this tile; every later one starts inside it (segment 0)
            const int seg_idx = (s == 0) ? tile - (int)((unsigned)sk_lo(cc, p.q, p.rem) / (unsigned)KB) : 0;
            const float* wt = p.ws + ((size_t)cc * p.max_segs + seg_idx) * SK_BM * SK_MB;
            const float* pa = wt + (size_t)tt * SK_MB;
            float4 va[8], vu[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) va[q] = ldcg4(pa + 4 * q);
            if (swiglu) {
#pragma unroll
              for (int q = 0; q < 8; ++q) vu[q] = ldcg4(pa + 64 * SK_MB + 4 * q);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              a[4 * q] += va[q].x; a[4 * q + 1] += va[q].y; a[4 * q + 2] += va[q].z; a[4 * q + 3] += va[q].w;
              if (swiglu) { u[4 * q] += vu[q].x; u[4 * q + 1] += vu[q].y; u[4 * q + 2] += vu[q].z; u[4 * q + 3] += vu[q].w; }
            }
          }
          if (tr) trp[5] = globaltimer_ns();
          if (swiglu) {
            const int n = tile * 64 + tt;
            __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C);
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (b < p.M) c[(size_t)b * p.ldc + n] = __float2bfloat16_rn(a[b] / (1.0f + __expf(-a[b])) * u[b]);
          } else {
            const int n = tile * SK_BM + tt;
            const float bias = p.bias ? p.bias[n] : 0.f;
#pragma unroll
            for (int b = 0; b < 32; ++b)  // residual loads first (C may alias the residual); u[] is free here
              u[b] = (p.residual != nullptr && b < p.M) ? __bfloat162float(p.residual[(size_t)b * p.ldr + n]) : 0.f;
#pragma unroll
            for (int b = 0; b < 32; ++b) {
              if (b < p.M) {
                const float v = a[b] + bias + u[b];
                if (p.out_dtype == CRAB_BF16) reinterpret_cast<__nv_bfloat16*>(p.C)[(size_t)b * p.ldc + n] = __float2bfloat16_rn(v);
                else reinterpret_cast<float*>(p.C)[(size_t)b * p.ldc + n] = v;
              }
            }
          }
        }
      }
      if (tr) trp[6] = globaltimer_ns();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
      f = seg_end;
      ++tile;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// Pre-pack a row-major weight [N, ldw] into the streaming layout: block f = tile * KB + kb holds the 128 x 64 tile
// exactly as the UMMA SWIZZLE_128B smem layout wants it (row r: 128 bytes, 16-byte chunk c stored at c ^ (r & 7)),
// zero-padded past N / K.  One CTA's stream-K range is then ONE contiguous span of HBM.
__global__ void pack_skinny_weight_kernel(const __nv_bfloat16* __restrict__ w, int N, int K, int ldw,
                                          __nv_bfloat16* __restrict__ out, int kb_per_tile, long long total_chunks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk per thread
  if (i >= total_chunks) return;
  const int cs = (int)(i & 7);                 // stored chunk position
  const int r = (int)((i >> 3) & 127);         // row within tile
  const long long f = i >> 10;                 // block index
  const int kb = (int)(f % kb_per_tile), tile = (int)(f / kb_per_tile);
  const int c = cs ^ (r & 7);                  // source chunk
  const int row = tile * SK_BM + r, col = kb * SK_BK + c * 8;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (row < N) {
    if (col + 8 <= K) v = *reinterpret_cast<const uint4*>(w + (size_t)row * ldw + col);
    else if (col < K) {
      __nv_bfloat16 t[8];
      for (int e = 0; e < 8; ++e) t[e] = (col + e < K) ? w[(size_t)row * ldw + col + e] : __float2bfloat16_rn(0.f);
      v = *reinterpret_cast<uint4*>(t);
    }
  }
  *reinterpret_cast<uint4*>(out + i * 8) = v;
}

static unsigned long long* g_trace = nullptr;  // diagnostic timestamps (CRAB_SK_TRACE=1)

// CTAs for (N, K): one per SM, but never so many that a CTA streams fewer than 8 k-blocks or a tile gets more than
// SK_MAX_SLOTS contributors.
int choose_ctas(int N, int K) {
  const long long tiles = (N + SK_BM - 1) / SK_BM;
  const long long kb = (K + SK_BK - 1) / SK_BK;
  const long long T = tiles * kb;
  long long g = sm_count();
  if (g > T / 8) g = T / 8 > 0 ? T / 8 : 1;
  return (int)g;
}

}  // namespace crab

using namespace crab;

extern "C" int crab_gemm_skinny_plan(int N, int K, int* ctas, int64_t* workspace_bytes, int* n_counters) {
  CRAB_REQUIRE(N > 0 && K > 0 && ctas && workspace_bytes && n_counters, "crab_gemm_skinny_plan: bad args");
  const int tiles = (N + SK_BM - 1) / SK_BM;
  *ctas = choose_ctas(N, K);
  *workspace_bytes = (int64_t)(tiles + 2 * (*ctas)) * SK_BM * SK_MB * 4;
  *n_counters = tiles;
  return CRAB_OK;
}

extern "C" int crab_debug_skinny_trace(unsigned long long* host_out, int n_ctas) {
  // diagnostic: copy the per-CTA epilogue timestamps (8 per CTA) of the latest launch made with CRAB_SK_TRACE=1
  using namespace crab;
  CRAB_REQUIRE(g_trace != nullptr && host_out != nullptr && n_ctas > 0 && n_ctas <= 1024, "crab_debug_skinny_trace: tracing is off");
  CRAB_CHECK_CUDA(cudaDeviceSynchronize());
  CRAB_CHECK_CUDA(cudaMemcpy(host_out, g_trace, (size_t)n_ctas * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return CRAB_OK;
}

extern "C" int crab_skinny_packed_bytes(int N, int K, int64_t* bytes) {
  CRAB_REQUIRE(N > 0 && K > 0 && bytes, "crab_skinny_packed_bytes: bad args");
  *bytes = (int64_t)((N + SK_BM - 1) / SK_BM) * ((K + SK_BK - 1) / SK_BK) * SK_BM * SK_BK * 2;
  return CRAB_OK;
}

extern "C" int crab_pack_skinny_weight(const void* W, int N, int K, int ldw, void* out, void* stream) {
  CRAB_REQUIRE(W && out && N > 0 && K > 0 && ldw >= K && ldw % 8 == 0, "crab_pack_skinny_weight: bad args");
  CRAB_REQUIRE(((uintptr_t)W % 16 == 0) && ((uintptr_t)out % 128 == 0), "crab_pack_skinny_weight: alignment (W 16 B, out 128 B)");
  const int kb = (K + SK_BK - 1) / SK_BK;
  const long long chunks = (long long)((N + SK_BM - 1) / SK_BM) * kb * SK_BM * 8;
  pack_skinny_weight_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(W), N, K, ldw, reinterpret_cast<__nv_bfloat16*>(out), kb, chunks);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_gemm_skinny_bf16(const crab_skinny_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CRAB_REQUIRE(a && a->X && (a->W || a->W_packed) && a->C && a->workspace && a->counters, "crab_gemm_skinny_bf16: null pointer");
  CRAB_REQUIRE(a->M > 0 && a->M <= SK_MB, "crab_gemm_skinny_bf16: M must be in 1..32 (got %d)", a->M);
  CRAB_REQUIRE(a->N > 0 && a->K > 0 && a->ldx % 8 == 0 && a->ldx >= a->K, "crab_gemm_skinny_bf16: bad shape/strides N=%d K=%d ldx=%d",
               a->N, a->K, a->ldx);
  if (a->W_packed) CRAB_REQUIRE((uintptr_t)a->W_packed % 128 == 0, "crab_gemm_skinny_bf16: W_packed must be 128-byte aligned");
  else CRAB_REQUIRE(a->ldw % 8 == 0 && a->ldw >= a->K && ((uintptr_t)a->W % 16 == 0), "crab_gemm_skinny_bf16: W alignment / ldw=%d", a->ldw);
  CRAB_REQUIRE((uintptr_t)a->X % 16 == 0, "crab_gemm_skinny_bf16: X must be 16-byte aligned");
  CRAB_REQUIRE(a->act == CRAB_ACT_NONE || a->act == CRAB_ACT_SWIGLU, "crab_gemm_skinny_bf16: act must be NONE or SWIGLU");
  if (a->act == CRAB_ACT_SWIGLU)
    CRAB_REQUIRE(a->N % 128 == 0 && !a->bias && !a->residual && a->out_dtype == CRAB_BF16, "crab_gemm_skinny_bf16: SWIGLU constraints");
  const int tiles = (a->N + SK_BM - 1) / SK_BM;
  const int kb = (a->K + SK_BK - 1) / SK_BK;
  const long long T = (long long)tiles * kb;
  long long ctas = a->splits > 0 ? a->splits : choose_ctas(a->N, a->K);  // `splits` = explicit CTA count (testing)
  if (ctas > T) ctas = T;
  // a tile of kb k-blocks may be shared by at most SK_MAX_SLOTS CTAs
  const long long min_range = (kb + SK_MAX_SLOTS - 2) / (SK_MAX_SLOTS - 1);
  if (T / ctas < min_range) ctas = T / min_range > 0 ? T / min_range : 1;
  CRAB_REQUIRE(a->n_counters >= tiles, "crab_gemm_skinny_bf16: need %d counters (got %d)", tiles, a->n_counters);
  const long long range_max = (T + ctas - 1) / ctas;
  const int max_segs = (int)((range_max + kb - 1) / kb) + 1;
  const long long ws_need = ctas * max_segs * (long long)SK_BM * SK_MB * 4;
  CRAB_REQUIRE(a->workspace_bytes >= ws_need, "crab_gemm_skinny_bf16: workspace too small (need %lld bytes)", ws_need);
  static bool attr_set = false;
  if (!attr_set) {
    CRAB_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM));
    attr_set = true;
  }
  CUtensorMap tw, tx;
  int rc = 0;
  if (a->W_packed) memset(&tw, 0, sizeof(tw));
  else rc = encode_tmap_bf16_2d(&tw, a->W, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldw, SK_BM, SK_BK);
  if (rc != 0) return rc;
  rc = encode_tmap_bf16_2d(&tx, a->X, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->ldx, SK_MB, SK_BK);
  if (rc != 0) return rc;
  SkinnyParams p;
  p.C = a->C; p.bias = a->bias; p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.ws = a->workspace; p.counters = a->counters;
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldc = a->ldc; p.ldr = a->ldr;
  p.act = a->act; p.out_dtype = a->out_dtype;
  p.w_tiled = reinterpret_cast<const __nv_bfloat16*>(a->W_packed);
  {
    static int want = -1;
    if (want < 0) { const char* e = getenv("CRAB_SK_TRACE"); want = (e && e[0] == '1') ? 1 : 0; }
    if (want && !g_trace) { cudaMalloc(&g_trace, 1024 * 8 * sizeof(unsigned long long)); cudaMemset(g_trace, 0, 1024 * 8 * sizeof(unsigned long long)); }
    p.trace = want ? g_trace : nullptr;
  }
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("CRAB_SK_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  p.tiles = tiles; p.kb_per_tile = kb; p.total_kb = (int)T; p.max_segs = max_segs;
  p.q = (int)(T / ctas); p.rem = (int)(T % ctas);
  CRAB_CHECK_CUDA(launch_pdl(gemm_skinny_tcgen05_kernel, dim3((unsigned)ctas), dim3(SK_THREADS), SK_SMEM, stream, tw, tx, p));
  return CRAB_OK;
}
