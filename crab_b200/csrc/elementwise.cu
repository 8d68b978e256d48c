// HBM-bound glue kernels of the hot path: norms, RoPE + KV-cache append, gathers/splices, patchify, CLIP embedding
// assembly, BEATs gate / pos-conv packing, arg-max.  All use 16-byte vector loads, fp32 math and warp-shuffle
// reductions; one row per warp or per block, grids sized by rows.
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = (l < nw) ? sh[l] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  f[0] = bf16lo(q.x); f[1] = bf16hi(q.x); f[2] = bf16lo(q.y); f[3] = bf16hi(q.y);
  f[4] = bf16lo(q.z); f[5] = bf16hi(q.z); f[6] = bf16lo(q.w); f[7] = bf16hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]); q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]); q.w = pack_bf16x2(f[6], f[7]);
  return q;
}

// ----------------------------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm: one block per row, row cached in registers (cols <= 8 * 8 * blockDim.x)
// ----------------------------------------------------------------------------------------------------------------
template <bool RMS, int VPT /* 8-element vectors per thread */>
__global__ void __launch_bounds__(256) norm_kernel(const __nv_bfloat16* x, int ldx,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   __nv_bfloat16* y, int ldy, int rows, int cols, float eps) {
  // x is produced by the previous kernel of a (possibly PDL) chain: ordered loads, no __restrict__ / __ldg (ptx.cuh: ld_dep_u4)
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[32];
  const int row = blockIdx.x;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  float v[VPT][8];
  const int nvec = cols >> 3;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      unpack8(ld_dep_u4(reinterpret_cast<const uint4*>(xr) + vi), v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += RMS ? v[i][j] * v[i][j] : v[i][j];
    }
  }
  s = block_sum(s, sh);
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(s / cols + eps);
  } else {
    mean = s / cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
      }
    }
    q = block_sum(q, sh);
    rstd = rsqrtf(q / cols + eps);
  }
  __nv_bfloat16* yr = y + (size_t)row * ldy;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      float o[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      if (RMS) {
        // HF LlamaRMSNorm: weight * (x_fp32 * rstd).to(bf16)   (models/modeling_llama.py:112-117)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = g[j] * __bfloat162float(__float2bfloat16_rn(v[i][j] * rstd));
      } else {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi + 1);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
      }
      *(reinterpret_cast<uint4*>(yr) + vi) = pack8(o);
    }
  }
}

template <bool RMS>
static int launch_norm(const void* x, int ldx, const float* gamma, const float* beta, void* y, int ldy, int rows,
                       int cols, float eps, cudaStream_t st) {
  CRAB_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "norm: cols/ld must be multiples of 8 (cols=%d)", cols);
  CRAB_REQUIRE(cols <= 16384, "norm: cols=%d too large", cols);
  if (rows <= 0) return CRAB_OK;
  const int nvec = cols / 8;
  auto xx = reinterpret_cast<const __nv_bfloat16*>(x);
  auto yy = reinterpret_cast<__nv_bfloat16*>(y);
  cudaError_t e;
  if (nvec <= 128) e = launch_pdl(PDL_LIGHT, norm_kernel<RMS, 1>, dim3(rows), dim3(128), 0, st, xx, ldx, gamma, beta, yy, ldy, rows, cols, eps);
  else if (nvec <= 256) e = launch_pdl(PDL_LIGHT, norm_kernel<RMS, 1>, dim3(rows), dim3(256), 0, st, xx, ldx, gamma, beta, yy, ldy, rows, cols, eps);
  else if (nvec <= 512) e = launch_pdl(PDL_LIGHT, norm_kernel<RMS, 2>, dim3(rows), dim3(256), 0, st, xx, ldx, gamma, beta, yy, ldy, rows, cols, eps);
  else if (nvec <= 1024) e = launch_pdl(PDL_LIGHT, norm_kernel<RMS, 4>, dim3(rows), dim3(256), 0, st, xx, ldx, gamma, beta, yy, ldy, rows, cols, eps);
  else e = launch_pdl(PDL_LIGHT, norm_kernel<RMS, 8>, dim3(rows), dim3(256), 0, st, xx, ldx, gamma, beta, yy, ldy, rows, cols, eps);
  CRAB_CHECK_CUDA(e);
  return CRAB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// RoPE table (accurate, one-time) and RoPE + KV-cache append
// ----------------------------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(float* cs, int max_pos, int half, double theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * half) return;
  const int pos = i / half, j = i % half;
  // fp32 inv_freq and fp32 product, as LlamaRotaryEmbedding does (models/modeling_llama.py:123-156)
  const float inv = 1.0f / (float)pow(theta, (double)(2 * j) / (double)(2 * half));
  const float ang = (float)pos * inv;
  double s, c;
  sincos((double)ang, &s, &c);
  cs[(size_t)pos * 2 * half + j] = (float)c;
  cs[(size_t)pos * 2 * half + half + j] = (float)s;
}

// qkv: [B*S, ldq] rows hold [q (H*hd) | k (KV*hd) | v (KV*hd)].  q is rotated in place; rotated k and v go to the
// cache [B, KV, ctx_max, hd] at position past + s.  One thread per (row, head, 8-element chunk of the low half): it
// owns the pair of 16-byte vectors (j..j+7, j+hd/2..j+hd/2+7), so every access is a 128-bit load/store.
template <int HD>
__global__ void __launch_bounds__(256) rope_kv_kernel(__nv_bfloat16* __restrict__ qkv, int ldq, const float* __restrict__ cs,
                                                      __nv_bfloat16* __restrict__ kc, __nv_bfloat16* __restrict__ vc, int B, int S,
                                                      int H, int KV, int ctx_max, const int* __restrict__ past_dev,
                                                      int past_host) {
  pdl_trigger();
  pdl_wait();
  constexpr int half = HD / 2;
  constexpr int CPH = half / 8;  // chunks per head
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int heads_total = H + 2 * KV;
  if (tid >= (long long)B * S * heads_total * CPH) return;
  const int ch = (int)(tid % CPH);
  const int hh = (int)((tid / CPH) % heads_total);
  const int row = (int)(tid / ((long long)CPH * heads_total));
  const int b = row / S, s = row % S;
  const int pos = (past_dev ? *past_dev : past_host) + s;
  __nv_bfloat16* src = qkv + (size_t)row * ldq + (size_t)hh * HD + ch * 8;
  const uint4 qlo = *reinterpret_cast<const uint4*>(src);
  const uint4 qhi = *reinterpret_cast<const uint4*>(src + half);
  if (hh < H + KV) {
    float lo[8], hi[8], o_lo[8], o_hi[8];
    unpack8(qlo, lo);
    unpack8(qhi, hi);
    const float* cp = cs + (size_t)pos * HD + ch * 8;
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(cp)), c1 = __ldg(reinterpret_cast<const float4*>(cp) + 1);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(cp + half)), s1 = __ldg(reinterpret_cast<const float4*>(cp + half) + 1);
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // out[j] = x[j] c[j] - x[j+h] s[j];  out[j+h] = x[j+h] c[j] + x[j] s[j]   (rotate_half, modeling_llama.py:204-236)
      o_lo[j] = lo[j] * c[j] - hi[j] * sn[j];
      o_hi[j] = hi[j] * c[j] + lo[j] * sn[j];
    }
    __nv_bfloat16* dst = src;
    if (hh >= H) dst = kc + (((size_t)b * KV + (hh - H)) * ctx_max + pos) * HD + ch * 8;
    *reinterpret_cast<uint4*>(dst) = pack8(o_lo);
    *reinterpret_cast<uint4*>(dst + half) = pack8(o_hi);
  } else {
    __nv_bfloat16* dst = vc + (((size_t)b * KV + (hh - H - KV)) * ctx_max + pos) * HD + ch * 8;
    *reinterpret_cast<uint4*>(dst) = qlo;
    *reinterpret_cast<uint4*>(dst + half) = qhi;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Row gather / scatter (embedding lookup, splice of modality embeddings, left padding)
// dst[dst_rows[i] or i] = src[src_rows[i] or i]   (cols bf16, multiples of 8)
// ----------------------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const __nv_bfloat16* src, int lds, const int64_t* src_rows,
                                   __nv_bfloat16* dst, int ldd, const int64_t* dst_rows, int n, int cols) {
  // the row indices (and, for the splice, the source rows) may come from the previous kernel of a PDL chain: plain loads
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x;
  const int64_t sr = src_rows ? *reinterpret_cast<const volatile int64_t*>(src_rows + i) : i;
  const int64_t dr = dst_rows ? *reinterpret_cast<const volatile int64_t*>(dst_rows + i) : i;
  const uint4* s = reinterpret_cast<const uint4*>(src + sr * lds);
  uint4* d = reinterpret_cast<uint4*>(dst + dr * ldd);
  for (int v = threadIdx.x; v < cols / 8; v += blockDim.x) d[v] = ld_dep_u4(s + v);
}

// fp32 -> bf16 rows (encoder inputs / fp32 features entering the bf16 path)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + i) = o;
  } else {
    for (; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Patchify: fp32 NCHW images -> bf16 patch rows [n_img * gh * gw, ld] with column order (c, kh, kw), i.e. the
// flattened Conv2d weight order, so the stride==kernel patch-embed conv is a GEMM.
// ----------------------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int ld, int C, int Hh,
                                int Ww, int p, int gh, int gw) {
  const int prow = blockIdx.x;  // patch row index
  const int n = prow / (gh * gw), r = prow % (gh * gw);
  const int py = r / gw, px = r % gw;
  const int kcols = C * p * p;
  __nv_bfloat16* o = out + (size_t)prow * ld;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) {
    float v = 0.f;
    if (k < kcols) {
      const int c = k / (p * p), rem = k % (p * p);
      const int ky = rem / p, kx = rem % p;
      v = img[(((size_t)n * C + c) * Hh + (py * p + ky)) * Ww + (px * p + kx)];
    }
    o[k] = __float2bfloat16_rn(v);
  }
}

// CLIP embeddings: x[n, 0] = cls + pos[0]; x[n, 1+t] = patch[n, t] + pos[1+t]; then pre_layrnorm.  One block per
// token row; D <= 2048.
__global__ void __launch_bounds__(256)
clip_embed_ln_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                     const float* __restrict__ pos, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ out, int tokens /*incl. cls*/, int D, float eps) {
  __shared__ float sh[32];
  const int row = blockIdx.x;
  const int n = row / tokens, t = row % tokens;
  float v[8];
  const int c0 = threadIdx.x * 8;
  float s = 0.f;
  if (c0 < D) {
    if (t == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = cls[c0 + j];
    } else {
      unpack8(__ldg(reinterpret_cast<const uint4*>(patch + ((size_t)n * (tokens - 1) + (t - 1)) * D + c0)), v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] += pos[(size_t)t * D + c0 + j]; s += v[j]; }
  }
  const float mean = block_sum(s, sh) / D;
  float q = 0.f;
  if (c0 < D) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q += d * d; }
  }
  const float rstd = rsqrtf(block_sum(q, sh) / D + eps);
  if (c0 < D) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * gamma[c0 + j] + beta[c0 + j];
    *reinterpret_cast<uint4*>(out + (size_t)row * D + c0) = pack8(o);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// BEATs helpers
// ----------------------------------------------------------------------------------------------------------------
// gate[b, h, t] = ga * (gb * grep_a[h] - 1) + 2 with (ga, gb) = sigmoid(sum of 4 outputs each of grep_linear(q))
// (models/beats/backbone.py:650-662).  q: [B*T, ldq] unscaled, head h at column h*64.  One warp per (row, head).
__global__ void beats_gate_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const float* __restrict__ gw /*[8,64]*/,
                                  const float* __restrict__ gb /*[8]*/, const float* __restrict__ grep_a /*[H]*/,
                                  float* __restrict__ gate /*[B,H,T]*/, int B, int T, int H) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * T * H) return;
  const int row = warp / H, h = warp % H;
  const int b = row / T, t = row % T;
  const __nv_bfloat16* qp = q + (size_t)row * ldq + h * 64;
  const float x0 = __bfloat162float(qp[lane]), x1 = __bfloat162float(qp[lane + 32]);
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = warp_sum(x0 * gw[o * 64 + lane] + x1 * gw[o * 64 + 32 + lane]);
  if (lane == 0) {
    float a = 0.f, bb = 0.f;
#pragma unroll
    for (int o = 0; o < 4; ++o) { a += acc[o] + gb[o]; bb += acc[4 + o] + gb[4 + o]; }
    const float ga = 1.f / (1.f + __expf(-a)), gbv = 1.f / (1.f + __expf(-bb));
    gate[((size_t)b * H + h) * T + t] = ga * (gbv * grep_a[h] - 1.f) + 2.f;
  }
}

// x [B, T, C] -> xg [G][B][T * cg]   (cg = C / G channels per conv group), for the Toeplitz-GEMM pos-conv.
__global__ void beats_group_pack_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ xg, int B,
                                        int T, int C, int G) {
  const int cg = C / G;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T * C) return;
  const int c = i % C;
  const int t = (i / C) % T;
  const int b = i / ((size_t)C * T);
  const int g = c / cg, ci = c % cg;
  xg[((size_t)g * B + b) * (T * cg) + t * cg + ci] = x[i];
}

// y[b,t,c] = x[b,t,c] + GELU(conv[g][b][t*cg+co] + bias[c])   (backbone.py:114-116), bf16 out (LayerNorm follows)
__global__ void beats_posconv_finish_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ cg_out,
                                            const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int B, int T,
                                            int C, int G) {
  const int cg = C / G;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T * C) return;
  const int c = i % C;
  const int t = (i / C) % T;
  const int b = i / ((size_t)C * T);
  const int g = c / cg, co = c % cg;
  const float v = __bfloat162float(cg_out[((size_t)g * B + b) * (T * cg) + t * cg + co]) + bias[c];
  const float ge = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  y[i] = __float2bfloat16_rn(__bfloat162float(x[i]) + ge);
}

// ----------------------------------------------------------------------------------------------------------------
// arg-max over the first V columns of fp32 logits rows (first index wins ties, like torch.argmax on CPU/CUDA)
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) argmax_kernel(const float* logits, int ld, int V, int64_t* out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sv[8];
  __shared__ int si[8];
  const float* r = logits + (size_t)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = r[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (w == 0) {
    best = (l < 8) ? sv[l] : -INFINITY;
    bi = (l < 8) ? si[l] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (l == 0) out[blockIdx.x] = bi;
  }
}


// Large vocabularies (Qwen2: 152 k logits per row): the row is split over the 8 CTAs of a cluster, the (max, index) partials
// meet in rank 0's shared memory (DSMEM) and rank 0 picks the winner — smallest index among equal maxima, like the
// single-block kernel.  133 us -> a few us per decode step at bs 32.
__global__ void __launch_bounds__(256) argmax_cluster_kernel(const float* logits, int ld, int V, int64_t* out) {
  // no __restrict__ / __ldg on the logits: they come from the previous kernel of the (possibly PDL) chain, and invariant loads
  // may be hoisted above griddepcontrol.wait
  pdl_trigger();
  pdl_wait();
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ float pv[8];
  __shared__ int pi[8];
  uint32_t rank, nranks;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nranks));
  const float* r = logits + (size_t)blockIdx.x * ld;
  const int per = ((V + (int)nranks - 1) / (int)nranks + 3) & ~3;  // multiple of 4: float4 loads (ld % 4 == 0 on the host)
  const int lo = (int)rank * per, hi = min(V, lo + per);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lo + 4 * (int)threadIdx.x; i < hi; i += 4 * (int)blockDim.x) {
    if (i + 3 < hi) {
      const float4 q = *reinterpret_cast<const float4*>(r + i);
      const float vv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (vv[e] > best || (vv[e] == best && i + e < bi)) { best = vv[e]; bi = i + e; }
    } else {
      for (int e = 0; i + e < hi; ++e) {
        const float v = r[i + e];
        if (v > best || (v == best && i + e < bi)) { best = v; bi = i + e; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (w == 0) {
    best = (l < 8) ? sv[l] : -INFINITY;
    bi = (l < 8) ? si[l] : 0x7fffffff;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (l == 0) {
      uint32_t a0 = smem_u32(&pv[rank]), a1 = smem_u32(&pi[rank]), r0, r1;
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r0) : "r"(a0));
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r1) : "r"(a1));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(r0), "f"(best) : "memory");
      asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(r1), "r"(bi) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0 && threadIdx.x == 0) {
    float b = -INFINITY;
    int ix = 0x7fffffff;
    for (uint32_t k = 0; k < nranks; ++k)
      if (pv[k] > b || (pv[k] == b && pi[k] < ix)) { b = pv[k]; ix = pi[k]; }
    out[blockIdx.x] = ix;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// Small-M (decode) row kernel: optional RMSNorm, then the hyper-LoRA pre-pass for up to 3 linears sharing the row:
//   t = y . [R_g ; A_g]^T  (11 outputs per linear),  z[g*24 + i*8 + j] = scale * softmax(t[g,0:3])_i * t[g,3+j]
// One block per row; the row stays in registers.  Replaces a norm launch + a 1-CTA skinny GEMM per linear group
// (peft_hyper/tuners/lora.py:344-350; models/modeling_llama.py:103-117).
// ----------------------------------------------------------------------------------------------------------------
template <bool NORM, int VPT>
__global__ void __launch_bounds__(256) row_loraz_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                        const float* __restrict__ gamma, float eps,
                                                        __nv_bfloat16* __restrict__ y, int ldy,
                                                        const __nv_bfloat16* __restrict__ ra, int ldra, int groups,
                                                        __nv_bfloat16* __restrict__ z, int ldz, float scale, int cols) {
  // grid = (rows, max(groups, 1)): block (row, g) owns the 11 outputs of linear g; the norm is recomputed per block
  // (8 KB row read) and written by the g == 0 block only.
  pdl_trigger();
  __shared__ float sh[32];
  __shared__ float red[8][11];
  __shared__ float tot[11];
  const int row = blockIdx.x, grp = blockIdx.y;
  const int nvec = cols >> 3;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const __nv_bfloat16* xr = x + (size_t)row * ldx;
  // The router/A rows are weights (no dependence on the previous kernel): for short rows fetch them BEFORE the PDL wait
  // so the loads fly while the producer of x drains, and during the norm reduction.
  constexpr bool PRE = (VPT <= 2);
  const __nv_bfloat16* rg = (groups > 0) ? ra + (size_t)grp * 11 * ldra : nullptr;
  uint4 rq[PRE ? VPT : 1][11];
  if (PRE && groups > 0) {
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = threadIdx.x + i * 256;
#pragma unroll
      for (int o = 0; o < 11; ++o)
        rq[PRE ? i : 0][o] = (vi < nvec) ? __ldg(reinterpret_cast<const uint4*>(rg + (size_t)o * ldra) + vi) : make_uint4(0, 0, 0, 0);
    }
  }
  pdl_wait();
  float v[VPT][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      unpack8(ld_dep_u4(reinterpret_cast<const uint4*>(xr) + vi), v[i]);  // x comes from the previous kernel: ordered load
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  if (NORM) {
    const float rstd = rsqrtf(block_sum(ss, sh) / cols + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = threadIdx.x + i * 256;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi);
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)  // same two roundings as norm_kernel<RMS>; the GEMM then reads these bf16 values
          v[i][j] = __bfloat162float(__float2bfloat16_rn(g[j] * __bfloat162float(__float2bfloat16_rn(v[i][j] * rstd))));
        if (grp == 0) *(reinterpret_cast<uint4*>(y + (size_t)row * ldy) + vi) = pack8(v[i]);
      }
    }
  }
  if (groups == 0) return;
  float acc[11];
#pragma unroll
  for (int o = 0; o < 11; ++o) acc[o] = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 256;
    if (vi < nvec) {
      uint4 q[11];
#pragma unroll
      for (int o = 0; o < 11; ++o)
        q[o] = PRE ? rq[PRE ? i : 0][o] : __ldg(reinterpret_cast<const uint4*>(rg + (size_t)o * ldra) + vi);  // 11 loads in flight
#pragma unroll
      for (int o = 0; o < 11; ++o) {
        float rf[8];
        unpack8(q[o], rf);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o] += v[i][j] * rf[j];
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 11; ++o) {
    acc[o] = warp_sum(acc[o]);
    if (l == 0) red[w][o] = acc[o];
  }
  __syncthreads();
  if (threadIdx.x < 11) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x < 24) {
    const int i = threadIdx.x / 8, j = threadIdx.x % 8;
    const float l0 = tot[0], l1 = tot[1], l2 = tot[2];
    const float mx = fmaxf(l0, fmaxf(l1, l2));
    const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
    const float ri = (i == 0 ? e0 : (i == 1 ? e1 : e2)) / (e0 + e1 + e2);
    z[(size_t)row * ldz + grp * 24 + threadIdx.x] = __float2bfloat16_rn(scale * ri * tot[3 + j]);
  }
}

// Cluster version for the decode step: the columns of one (row, linear) are split over the P CTAs of a thread-block
// cluster, so a 32-row batch fills 256-384 CTAs instead of 32-96 and every CTA's dependent-load chain is P times shorter.
// Two exchanges through distributed shared memory: (1) the partial sums of squares go to every peer (each CTA needs rstd
// to normalise its slice), (2) the 11 partial dots go to rank 0, which applies the router softmax and writes z.  Partials
// are always summed in rank order: deterministic.  128 threads, VPT x 8 columns per thread.
__device__ __forceinline__ uint32_t rl_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void rl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void rl_st_remote(const void* local_smem, uint32_t rank, float v) {
  uint32_t a = smem_u32(local_smem), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(r), "f"(v) : "memory");
}

template <bool NORM, int VPT>
__global__ void __launch_bounds__(128) row_loraz_cluster_kernel(const __nv_bfloat16* x, int ldx, const float* __restrict__ gamma,
                                                                float eps, __nv_bfloat16* y, int ldy,
                                                                const __nv_bfloat16* __restrict__ ra, int ldra, int groups,
                                                                __nv_bfloat16* z, int ldz, float scale, int cols, int P) {
  pdl_trigger();
  __shared__ float sh[32];
  __shared__ float ss_part[8];        // [rank] partial sums of squares (written by every peer)
  __shared__ float dot_part[8][11];   // [rank][11] partial dots (rank 0 only)
  __shared__ float red[4][11];
  __shared__ float tot[11];
  const int row = blockIdx.x, grp = blockIdx.y;
  const int rank = (int)rl_cluster_rank();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int nvec_slice = (cols >> 3) / P;          // host guarantees divisibility
  const int v0 = rank * nvec_slice;
  const __nv_bfloat16* rg = (groups > 0) ? ra + (size_t)grp * 11 * ldra : nullptr;
  // weights first (they do not depend on the previous kernel): 11 router/A vectors and gamma for this thread's columns
  constexpr bool PRE = (VPT <= 2);  // wider slices fetch the router/A vectors inside the dot loop (register budget)
  uint4 rq[PRE ? VPT : 1][11];
  float4 gq[VPT][2];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 128;
    const bool ok = vi < nvec_slice;
    if (PRE && groups > 0) {
#pragma unroll
      for (int o = 0; o < 11; ++o)
        rq[PRE ? i : 0][o] = ok ? __ldg(reinterpret_cast<const uint4*>(rg + (size_t)o * ldra) + v0 + vi) : make_uint4(0, 0, 0, 0);
    }
    if (NORM) {
      gq[i][0] = ok ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * (v0 + vi)) : make_float4(0, 0, 0, 0);
      gq[i][1] = ok ? __ldg(reinterpret_cast<const float4*>(gamma) + 2 * (v0 + vi) + 1) : make_float4(0, 0, 0, 0);
    }
  }
  pdl_wait();
  float v[VPT][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 128;
    if (vi < nvec_slice) {
      unpack8(ld_dep_u4(reinterpret_cast<const uint4*>(x + (size_t)row * ldx) + v0 + vi), v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  if (NORM) {
    const float part = block_sum(ss, sh);
    if (threadIdx.x < P) rl_st_remote(&ss_part[rank], (uint32_t)threadIdx.x, part);
    rl_cluster_sync();
    float total = 0.f;
    for (int r = 0; r < P; ++r) total += ss_part[r];
    const float rstd = rsqrtf(total / cols + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int vi = threadIdx.x + i * 128;
      if (vi < nvec_slice) {
        const float g[8] = {gq[i][0].x, gq[i][0].y, gq[i][0].z, gq[i][0].w, gq[i][1].x, gq[i][1].y, gq[i][1].z, gq[i][1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j)  // same two roundings as norm_kernel<RMS>
          v[i][j] = __bfloat162float(__float2bfloat16_rn(g[j] * __bfloat162float(__float2bfloat16_rn(v[i][j] * rstd))));
        if (grp == 0) *(reinterpret_cast<uint4*>(y + (size_t)row * ldy) + v0 + vi) = pack8(v[i]);
      }
    }
  }
  if (groups == 0) return;  // uniform over the cluster: nobody waits on a barrier below
  float acc[11];
#pragma unroll
  for (int o = 0; o < 11; ++o) acc[o] = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int vi = threadIdx.x + i * 128;
    uint4 q11[11];
#pragma unroll
    for (int o = 0; o < 11; ++o)
      q11[o] = PRE ? rq[PRE ? i : 0][o]
                   : (vi < nvec_slice ? __ldg(reinterpret_cast<const uint4*>(rg + (size_t)o * ldra) + v0 + vi) : make_uint4(0, 0, 0, 0));
#pragma unroll
    for (int o = 0; o < 11; ++o) {
      float rf[8];
      unpack8(q11[o], rf);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[o] += v[i][j] * rf[j];
    }
  }
#pragma unroll
  for (int o = 0; o < 11; ++o) {
    acc[o] = warp_sum(acc[o]);
    if (l == 0) red[w][o] = acc[o];
  }
  __syncthreads();
  if (threadIdx.x < 11) rl_st_remote(&dot_part[rank][threadIdx.x], 0u, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
  rl_cluster_sync();
  if (rank != 0) return;
  if (threadIdx.x < 11) {
    float t = 0.f;
    for (int r = 0; r < P; ++r) t += dot_part[r][threadIdx.x];
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x < 24) {
    const int i = threadIdx.x / 8, j = threadIdx.x % 8;
    const float l0 = tot[0], l1 = tot[1], l2 = tot[2];
    const float mx = fmaxf(l0, fmaxf(l1, l2));
    const float e0 = __expf(l0 - mx), e1 = __expf(l1 - mx), e2 = __expf(l2 - mx);
    const float ri = (i == 0 ? e0 : (i == 1 ? e1 : e2)) / (e0 + e1 + e2);
    z[(size_t)row * ldz + grp * 24 + threadIdx.x] = __float2bfloat16_rn(scale * ri * tot[3 + j]);
  }
}

// device-side scalar increment (decode position counter inside a CUDA graph)
__global__ void add_scalar_kernel(int* p, int v) {
  pdl_trigger();
  pdl_wait();
  *p += v;
}

}  // namespace crab

using namespace crab;

extern "C" int crab_layernorm(const void* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                              int rows, int cols, float eps, void* stream) {
  CRAB_REQUIRE(x && y && gamma && beta, "crab_layernorm: null pointer");
  return launch_norm<false>(x, ldx, gamma, beta, y, ldy, rows, cols, eps, (cudaStream_t)stream);
}

extern "C" int crab_rmsnorm(const void* x, int ldx, const float* gamma, void* y, int ldy, int rows, int cols,
                            float eps, void* stream) {
  CRAB_REQUIRE(x && y && gamma, "crab_rmsnorm: null pointer");
  return launch_norm<true>(x, ldx, gamma, nullptr, y, ldy, rows, cols, eps, (cudaStream_t)stream);
}

extern "C" int crab_rope_table(float* cos_sin, int max_pos, int head_dim, double theta, void* stream) {
  CRAB_REQUIRE(cos_sin && max_pos > 0 && head_dim % 2 == 0, "crab_rope_table: bad args");
  const int n = max_pos * (head_dim / 2);
  rope_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cos_sin, max_pos, head_dim / 2, theta);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_rope_kv_append(void* qkv, int ldq, const float* cos_sin, void* k_cache, void* v_cache, int B, int S,
                                   int H, int KV, int head_dim, int ctx_max, const int* past_dev, int past_host,
                                   void* stream) {
  CRAB_REQUIRE(qkv && cos_sin && k_cache && v_cache, "crab_rope_kv_append: null pointer");
  CRAB_REQUIRE(head_dim == 64 || head_dim == 128, "crab_rope_kv_append: head_dim must be 64 or 128 (got %d)", head_dim);
  CRAB_REQUIRE(past_dev != nullptr || past_host + S <= ctx_max, "crab_rope_kv_append: past+S exceeds ctx_max");
  CRAB_REQUIRE(ldq % 8 == 0 && ((uintptr_t)qkv % 16 == 0), "crab_rope_kv_append: qkv must be 16-byte aligned with ldq %% 8 == 0");
  const long long threads = (long long)B * S * (H + 2 * KV) * (head_dim / 16);
  if (threads == 0) return CRAB_OK;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  auto q = reinterpret_cast<__nv_bfloat16*>(qkv);
  auto kc = reinterpret_cast<__nv_bfloat16*>(k_cache);
  auto vc = reinterpret_cast<__nv_bfloat16*>(v_cache);
  if (head_dim == 128)
    CRAB_CHECK_CUDA(launch_pdl(PDL_ROPE, rope_kv_kernel<128>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, q, ldq, cos_sin, kc, vc, B, S, H, KV, ctx_max, past_dev, past_host));
  else
    CRAB_CHECK_CUDA(launch_pdl(PDL_ROPE, rope_kv_kernel<64>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, q, ldq, cos_sin, kc, vc, B, S, H, KV, ctx_max, past_dev, past_host));
  return CRAB_OK;
}

extern "C" int crab_gather_rows(const void* src, int lds, const int64_t* src_rows, void* dst, int ldd,
                                const int64_t* dst_rows, int n, int cols, void* stream) {
  CRAB_REQUIRE(src && dst, "crab_gather_rows: null pointer");
  CRAB_REQUIRE(cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "crab_gather_rows: cols/ld must be multiples of 8");
  if (n <= 0) return CRAB_OK;
  CRAB_CHECK_CUDA(launch_pdl(PDL_LIGHT, gather_rows_kernel, dim3(n), dim3(128), 0, (cudaStream_t)stream,
                             reinterpret_cast<const __nv_bfloat16*>(src), lds, src_rows,
                             reinterpret_cast<__nv_bfloat16*>(dst), ldd, dst_rows, n, cols));
  return CRAB_OK;
}

extern "C" int crab_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  CRAB_REQUIRE(src && dst && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0), "crab_cast_f32_bf16: bad pointer");
  if (n <= 0) return CRAB_OK;
  const int64_t th = (n + 3) / 4;
  cast_f32_bf16_kernel<<<(unsigned)((th + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), (size_t)n);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_patchify(const float* images, void* out, int ld_out, int n_img, int C, int H, int W, int patch,
                             void* stream) {
  CRAB_REQUIRE(images && out && patch > 0, "crab_patchify: bad args");
  const int gh = H / patch, gw = W / patch;
  CRAB_REQUIRE(ld_out >= C * patch * patch && ld_out % 8 == 0, "crab_patchify: ld_out too small / unaligned");
  const int rows = n_img * gh * gw;
  if (rows <= 0) return CRAB_OK;
  patchify_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(images, reinterpret_cast<__nv_bfloat16*>(out), ld_out, C, H, W, patch, gh, gw);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_clip_embed_ln(const void* patch_emb, const float* cls, const float* pos, const float* gamma,
                                  const float* beta, void* out, int n_img, int tokens, int D, float eps, void* stream) {
  CRAB_REQUIRE(patch_emb && cls && pos && gamma && beta && out, "crab_clip_embed_ln: null pointer");
  CRAB_REQUIRE(D % 8 == 0 && D <= 2048, "crab_clip_embed_ln: D must be <= 2048 and a multiple of 8");
  if (n_img <= 0) return CRAB_OK;
  clip_embed_ln_kernel<<<n_img * tokens, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(patch_emb), cls, pos, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), tokens, D, eps);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_beats_gate(const void* q, int ldq, const float* grep_w, const float* grep_b, const float* grep_a,
                               float* gate, int B, int T, int H, void* stream) {
  CRAB_REQUIRE(q && grep_w && grep_b && grep_a && gate, "crab_beats_gate: null pointer");
  const long warps = (long)B * T * H;
  if (warps == 0) return CRAB_OK;
  beats_gate_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), ldq, grep_w, grep_b, grep_a, gate, B, T, H);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_beats_group_pack(const void* x, void* xg, int B, int T, int C, int G, void* stream) {
  CRAB_REQUIRE(x && xg && G > 0 && C % G == 0, "crab_beats_group_pack: bad args");
  const size_t n = (size_t)B * T * C;
  if (n == 0) return CRAB_OK;
  beats_group_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(xg), B, T, C, G);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_beats_posconv_finish(const void* x, const void* conv_g, const float* bias, void* y, int B, int T,
                                         int C, int G, void* stream) {
  CRAB_REQUIRE(x && conv_g && bias && y && G > 0 && C % G == 0, "crab_beats_posconv_finish: bad args");
  const size_t n = (size_t)B * T * C;
  if (n == 0) return CRAB_OK;
  beats_posconv_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(conv_g), bias,
      reinterpret_cast<__nv_bfloat16*>(y), B, T, C, G);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_argmax(const float* logits, int ld, int rows, int V, int64_t* out, void* stream) {
  CRAB_REQUIRE(logits && out && V > 0, "crab_argmax: bad args");
  if (rows <= 0) return CRAB_OK;
  if (V >= 16384 && ld % 4 == 0 && ((uintptr_t)logits % 16 == 0)) {  // one cluster of 8 CTAs per row
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)rows, 8);
    cfg.blockDim = dim3(256);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 8; at[0].val.clusterDim.z = 1;
    int na = 1;
    if (pdl_mask() & PDL_LIGHT) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, argmax_cluster_kernel, logits, ld, V, out));
    return CRAB_OK;
  }
  CRAB_CHECK_CUDA(launch_pdl(PDL_LIGHT, argmax_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, logits, ld, V, out));
  return CRAB_OK;
}


extern "C" int crab_row_norm_loraz(const void* x, int ldx, const float* gamma, float eps, void* y, int ldy,
                                   const void* ra, int ldra, int groups, void* z, int ldz, float scale, int rows,
                                   int cols, void* stream) {
  CRAB_REQUIRE(x != nullptr, "crab_row_norm_loraz: null x");
  CRAB_REQUIRE(cols % 8 == 0 && ldx % 8 == 0 && cols <= 8 * 512 * 8, "crab_row_norm_loraz: cols=%d unsupported", cols);
  CRAB_REQUIRE(groups >= 0 && groups <= 3, "crab_row_norm_loraz: groups must be 0..3");
  CRAB_REQUIRE(groups == 0 || (ra && z && ldra % 8 == 0), "crab_row_norm_loraz: ra/z needed when groups > 0");
  CRAB_REQUIRE((gamma == nullptr) == (y == nullptr), "crab_row_norm_loraz: gamma and y go together");
  if (gamma) CRAB_REQUIRE(ldy % 8 == 0, "crab_row_norm_loraz: ldy alignment");
  if (rows <= 0) return CRAB_OK;
  const int nvec = cols / 8;
  const int vpt = (nvec + 255) / 256;
  {
    // decode batches: split the columns of each row over a cluster of P CTAs (largest P <= 8 that divides the row into
    // slices of at most 256 vectors, preferring >= 64 vectors per CTA)
    static int use_cluster = -1;
    if (use_cluster < 0) { const char* e = getenv("CRAB_ROW_CLUSTER"); use_cluster = (e && e[0] == '0') ? 0 : 1; }
    int P = 0, V = 2;
    if (use_cluster && rows <= 64 && nvec >= 256) {
      for (int c = 8; c >= 2 && P == 0; c >>= 1)
        if (nvec % c == 0 && nvec / c <= 256 && nvec / c >= 64) P = c;
      if (P == 8 && nvec % 4 == 0 && nvec / 4 <= 128) P = 4;  // 128 vectors per CTA keeps every thread busy with one vector
      if (P == 0 && nvec % 8 == 0 && nvec / 8 <= 512) { P = 8; V = 4; }  // very wide rows (Qwen2-7B's 18944-column MLP)
    }
    if (P > 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)rows, (unsigned)(groups > 0 ? groups : 1), (unsigned)P);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)P;
      int na = 1;
      if (pdl_mask() & PDL_ROW) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
      cfg.attrs = at;
      cfg.numAttrs = na;
      auto xx2 = reinterpret_cast<const __nv_bfloat16*>(x);
      auto yy2 = reinterpret_cast<__nv_bfloat16*>(y);
      auto rr2 = reinterpret_cast<const __nv_bfloat16*>(ra);
      auto zz2 = reinterpret_cast<__nv_bfloat16*>(z);
      if (gamma && V == 2) CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, row_loraz_cluster_kernel<true, 2>, xx2, ldx, gamma, eps, yy2, ldy, rr2, ldra, groups, zz2, ldz, scale, cols, P));
      else if (gamma) CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, row_loraz_cluster_kernel<true, 4>, xx2, ldx, gamma, eps, yy2, ldy, rr2, ldra, groups, zz2, ldz, scale, cols, P));
      else if (V == 2) CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, row_loraz_cluster_kernel<false, 2>, xx2, ldx, gamma, eps, yy2, ldy, rr2, ldra, groups, zz2, ldz, scale, cols, P));
      else CRAB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, row_loraz_cluster_kernel<false, 4>, xx2, ldx, gamma, eps, yy2, ldy, rr2, ldra, groups, zz2, ldz, scale, cols, P));
      return CRAB_OK;
    }
  }
  auto xx = reinterpret_cast<const __nv_bfloat16*>(x);
  auto yy = reinterpret_cast<__nv_bfloat16*>(y);
  auto rr = reinterpret_cast<const __nv_bfloat16*>(ra);
  auto zz = reinterpret_cast<__nv_bfloat16*>(z);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define CRAB_ROW_CASE(V)                                                                                              \
  if (vpt <= V) {                                                                                                     \
    const dim3 grid(rows, groups > 0 ? groups : 1);                                                                   \
    if (gamma) e = launch_pdl(PDL_ROW, row_loraz_kernel<true, V>, grid, dim3(256), 0, st, xx, ldx, gamma, eps, yy, ldy, rr, ldra, groups, zz, ldz, scale, cols); \
    else e = launch_pdl(PDL_ROW, row_loraz_kernel<false, V>, grid, dim3(256), 0, st, xx, ldx, gamma, eps, yy, ldy, rr, ldra, groups, zz, ldz, scale, cols);      \
  } else
  CRAB_ROW_CASE(2) CRAB_ROW_CASE(4) CRAB_ROW_CASE(8) { return set_error(CRAB_ERR_INVALID, "crab_row_norm_loraz: cols too large"); }
#undef CRAB_ROW_CASE
  CRAB_CHECK_CUDA(e);
  return CRAB_OK;
}

extern "C" int crab_add_scalar_i32(int* p, int v, void* stream) {
  CRAB_REQUIRE(p != nullptr, "crab_add_scalar_i32: null pointer");
  CRAB_CHECK_CUDA(launch_pdl(PDL_LIGHT, add_scalar_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, p, v));
  return CRAB_OK;
}
