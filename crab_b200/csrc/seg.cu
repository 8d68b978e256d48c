// Small kernels of the segmentation head (SegModule / MaskDecoderMultiScale, models/multimodal_encoder.py:
// 268-543, 891-1445).  Feature maps are kept token-major ("NHWC": one row per pixel, channels contiguous), so every 1x1
// conv, hyper-MLP, conv-transpose(k=2,s=2) and the 3x3 conv (after im2col) is a crab_gemm_bf16 call and LayerNorm2d is
// the row LayerNorm.  What is left are these latency-class helpers:
//
//   small_attn_kernel     — softmax(q k^T / sqrt(hd)) v for the SAM-style attention layers: 8 heads of 16 or 32 channels,
//                           <= 1024 keys (Attention.forward :1368-1393); one thread per (query, head), keys staged
//                           through shared memory 256 at a time, online softmax in fp32.
//   ew_kernel             — add with row broadcast (pos-embeddings, level embedding, no-mask embedding), ReLU, GELU,
//                           and the mask gate (sigmoid(m) + 1) * x of predict_masks (:1110-1112).
//   row_mean_kernel       — mean over the class channels of the previous level's masks (:1110).
//   im2col3x3_kernel      — 3x3 / pad 1 neighbourhoods for image_feature_neck's second conv (:318-325).
//   bilinear_kernel       — F.interpolate(mode="bilinear", align_corners=False) on token-major fp32 maps, with
//                           out = beta * out + alpha * interp (the multi-scale accumulation, :455) and an NCHW output
//                           option for the final (num_classes, 224, 224) masks (:501-543).
#include "host_common.h"
#include "ptx.cuh"

namespace crab {

template <int HD>
__global__ void __launch_bounds__(128) small_attn_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                                         const __nv_bfloat16* __restrict__ k, int ldk,
                                                         const __nv_bfloat16* __restrict__ v, int ldv,
                                                         __nv_bfloat16* __restrict__ o, int ldo, int Nq, int Nk, float scale) {
  constexpr int KT = 256;
  __shared__ __nv_bfloat16 sk[KT][HD];
  __shared__ __nv_bfloat16 sv[KT][HD];
  const int h = blockIdx.y;
  const int row = blockIdx.x * 128 + threadIdx.x;
  const bool active = row < Nq;
  float qf[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    qf[d] = active ? __bfloat162float(q[(size_t)row * ldq + h * HD + d]) * scale : 0.f;
    acc[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < Nk; k0 += KT) {
    const int nk = min(KT, Nk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * HD; i += 128) {
      const int kk = i / HD, d = i % HD;
      sk[kk][d] = k[(size_t)(k0 + kk) * ldk + h * HD + d];
      sv[kk][d] = v[(size_t)(k0 + kk) * ldv + h * HD + d];
    }
    __syncthreads();
    if (active) {
      for (int kk = 0; kk < nk; ++kk) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) s += qf[d] * __bfloat162float(sk[kk][d]);
        const float mn = fmaxf(m, s);
        const float c = __expf(m - mn);   // 0 on the first key (m = -inf)
        const float p = __expf(s - mn);
        m = mn;
        l = l * c + p;
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = acc[d] * c + p * __bfloat162float(sv[kk][d]);
      }
    }
  }
  if (active) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[(size_t)row * ldo + h * HD + d] = __float2bfloat16_rn(acc[d] * inv);
  }
}

// op: 0 out = a + b (b has b_rows rows: 1 = broadcast over rows), 1 out = relu(a), 2 out = gelu(a), 3 out = (sigmoid(g[r]) + 1) * a
__global__ void ew_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb, int b_rows,
                          const float* __restrict__ g, __nv_bfloat16* __restrict__ out, int ldo, int rows, int cols, int op) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  const float x = __bfloat162float(a[(size_t)r * lda + c]);
  float y;
  if (op == 0) y = x + __bfloat162float(b[(size_t)(b_rows == 1 ? 0 : r) * ldb + c]);
  else if (op == 1) y = fmaxf(x, 0.f);
  else if (op == 2) y = 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
  else y = (1.0f / (1.0f + __expf(-g[r])) + 1.0f) * x;
  out[(size_t)r * ldo + c] = __float2bfloat16_rn(y);
}

__global__ void row_mean_kernel(const float* __restrict__ x, int ldx, int rows, int cols, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = 0; c < cols; ++c) s += x[(size_t)r * ldx + c];
  out[r] = s / (float)cols;
}

// in [h*w, C] (row stride ldi) -> out [h*w, 9*C], column order (ky, kx, c), zero padding outside the map
__global__ void im2col3x3_kernel(const __nv_bfloat16* __restrict__ in, int ldi, __nv_bfloat16* __restrict__ out, int h, int w,
                                 int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)h * w * 9 * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int tap = (int)((i / C) % 9);
  const int t = (int)(i / (9LL * C));
  const int y = t / w + tap / 3 - 1, x = t % w + tap % 3 - 1;
  __nv_bfloat16 val = __float2bfloat16_rn(0.f);
  if (y >= 0 && y < h && x >= 0 && x < w) val = in[(size_t)(y * w + x) * ldi + c];
  out[i] = val;
}

// token-major fp32 [hin*win, ldi] (C channels) -> [hout*wout, ldo] (nchw = 0) or [C, hout, wout] (nchw = 1)
__global__ void bilinear_kernel(const float* __restrict__ in, int ldi, int hin, int win, float* __restrict__ out, int ldo, int hout,
                                int wout, int C, float alpha, float beta, int nchw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)hout * wout * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int ox = (int)((i / C) % wout);
  const int oy = (int)(i / ((long long)C * wout));
  // PyTorch area_pixel_compute_source_index, align_corners = False: src = max(0, scale * (dst + 0.5) - 0.5), scale = in / out
  const float sy = fmaxf(0.f, ((float)hin / (float)hout) * ((float)oy + 0.5f) - 0.5f);
  const float sx = fmaxf(0.f, ((float)win / (float)wout) * ((float)ox + 0.5f) - 0.5f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, hin - 1), x1 = min(x0 + 1, win - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float v00 = in[(size_t)(y0 * win + x0) * ldi + c], v01 = in[(size_t)(y0 * win + x1) * ldi + c];
  const float v10 = in[(size_t)(y1 * win + x0) * ldi + c], v11 = in[(size_t)(y1 * win + x1) * ldi + c];
  const float val = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  const size_t o = nchw ? ((size_t)c * hout + oy) * wout + ox : (size_t)(oy * wout + ox) * ldo + c;
  out[o] = (beta != 0.f ? beta * out[o] : 0.f) + alpha * val;
}

}  // namespace crab

using namespace crab;

extern "C" int crab_small_attn(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int Nq,
                               int Nk, int heads, int head_dim, float scale, void* stream) {
  CRAB_REQUIRE(q && k && v && o, "crab_small_attn: null pointer");
  CRAB_REQUIRE(head_dim == 16 || head_dim == 32, "crab_small_attn: head_dim must be 16 or 32 (got %d)", head_dim);
  CRAB_REQUIRE(Nq > 0 && Nk > 0 && heads > 0, "crab_small_attn: bad shape");
  CRAB_REQUIRE(ldq >= heads * head_dim && ldk >= heads * head_dim && ldv >= heads * head_dim && ldo >= heads * head_dim,
               "crab_small_attn: row strides smaller than heads * head_dim");
  const dim3 grid((unsigned)((Nq + 127) / 128), (unsigned)heads);
  auto qq = reinterpret_cast<const __nv_bfloat16*>(q);
  auto kk = reinterpret_cast<const __nv_bfloat16*>(k);
  auto vv = reinterpret_cast<const __nv_bfloat16*>(v);
  auto oo = reinterpret_cast<__nv_bfloat16*>(o);
  if (head_dim == 16) small_attn_kernel<16><<<grid, 128, 0, (cudaStream_t)stream>>>(qq, ldq, kk, ldk, vv, ldv, oo, ldo, Nq, Nk, scale);
  else small_attn_kernel<32><<<grid, 128, 0, (cudaStream_t)stream>>>(qq, ldq, kk, ldk, vv, ldv, oo, ldo, Nq, Nk, scale);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_elementwise(const void* a, int lda, const void* b, int ldb, int b_rows, const float* gate, void* out, int ldo,
                                int rows, int cols, int op, void* stream) {
  CRAB_REQUIRE(a && out && rows > 0 && cols > 0, "crab_elementwise: bad args");
  CRAB_REQUIRE(op >= 0 && op <= 3, "crab_elementwise: op must be 0 (add) / 1 (relu) / 2 (gelu) / 3 (mask gate)");
  CRAB_REQUIRE(op != 0 || (b != nullptr && (b_rows == 1 || b_rows == rows)), "crab_elementwise: add needs b with 1 or `rows` rows");
  CRAB_REQUIRE(op != 3 || gate != nullptr, "crab_elementwise: the mask gate needs the per-row fp32 gate");
  const long long n = (long long)rows * cols;
  ew_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), lda, reinterpret_cast<const __nv_bfloat16*>(b), ldb, b_rows, gate,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, cols, op);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_row_mean_f32(const float* x, int ldx, int rows, int cols, float* out, void* stream) {
  CRAB_REQUIRE(x && out && rows > 0 && cols > 0 && ldx >= cols, "crab_row_mean_f32: bad args");
  row_mean_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, ldx, rows, cols, out);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_im2col3x3(const void* in, int ldi, void* out, int h, int w, int C, void* stream) {
  CRAB_REQUIRE(in && out && h > 0 && w > 0 && C > 0 && ldi >= C, "crab_im2col3x3: bad args");
  const long long n = (long long)h * w * 9 * C;
  im2col3x3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), ldi, reinterpret_cast<__nv_bfloat16*>(out), h, w, C);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_bilinear_f32(const float* in, int ldi, int hin, int win, float* out, int ldo, int hout, int wout, int C,
                                 float alpha, float beta, int nchw_out, void* stream) {
  CRAB_REQUIRE(in && out && hin > 0 && win > 0 && hout > 0 && wout > 0 && C > 0 && ldi >= C, "crab_bilinear_f32: bad args");
  CRAB_REQUIRE(nchw_out || ldo >= C, "crab_bilinear_f32: ldo smaller than C");
  const long long n = (long long)hout * wout * C;
  bilinear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, hin, win, out, ldo, hout, wout, C, alpha,
                                                                             beta, nchw_out);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}
