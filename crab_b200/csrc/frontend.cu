// Front-end preprocessing on the GPU: the step immediately before the hot path.
//
//  fbank_kernel        — Kaldi-compatible log-mel filterbank, i.e. dataset/audio_processor.py:29-41
//                        (`ta_kaldi.fbank(waveform * 2**15, num_mel_bins=128, sample_frequency=16000, frame_length=25,
//                        frame_shift=10)` followed by `(fbank - 15.41663) / (2 * 6.55582)`): per frame DC removal,
//                        pre-emphasis 0.97, Povey window, zero-pad 400 -> 512, 512-point FFT in shared memory, power
//                        spectrum, 128 triangular mel filters (sparse rows, built on the host from the Kaldi formula),
//                        log, normalisation.  One block per frame; fp32 throughout.
//  patchify_u8_kernel  — CLIPImageProcessor's rescale (1/255) + normalise ((x - mean) / std) fused with the im2col of
//                        the patch-embed conv: uint8 frames [n, H, W, 3] as the video decoder delivers them
//                        (dataset/quick_start_dataset.py:303-315) -> bf16 patch rows, skipping the fp32 NCHW tensor.
//  normalize_u8_kernel — the same arithmetic to a fp32 NCHW tensor (`pixel_values`), for callers that want the
//                        reference's intermediate.
//  resample_u8_kernel  — Pillow's 8-bit bicubic resampler (the processor's shortest-edge resize), one separable pass,
//                        integer arithmetic, bit-exact; the centre crop is folded in.
#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int FB_WIN = 400;     // 25 ms @ 16 kHz
static constexpr int FB_SHIFT = 160;   // 10 ms
static constexpr int FB_N = 512;       // round_to_power_of_two
static constexpr int FB_LOGN = 9;
static constexpr int FB_THREADS = 256;

struct FbankParams {
  const float* wave;      // [n_seg, wave_stride] fp32 in [-1, 1]
  long long wave_stride;
  int n_seg, n_frames, n_mel;
  const float* window;    // [FB_WIN]
  const float* twiddle;   // [FB_N / 2][2] = (cos, -sin)(2 pi k / 512), rounded from float64 on the host
  const int* mel_start;   // [n_mel]  first FFT bin with a non-zero weight
  const int* mel_off;     // [n_mel + 1] offsets into mel_w
  const float* mel_w;     // packed non-zero weights, ascending bin order
  float* out;             // [n_seg, n_frames, n_mel]
  float in_scale, mean, inv_scale;
};

__global__ void __launch_bounds__(FB_THREADS) fbank_kernel(const FbankParams p) {
  __shared__ float s_re[FB_N], s_im[FB_N], s_x[FB_N];
  __shared__ float tw_re[FB_N / 2], tw_im[FB_N / 2];
  __shared__ float red[FB_THREADS / 32];
  const int t = threadIdx.x;
  const int seg = blockIdx.x / p.n_frames, frame = blockIdx.x % p.n_frames;
  const float* w = p.wave + (size_t)seg * p.wave_stride + (size_t)frame * FB_SHIFT;

  // twiddles exp(-2 pi i k / 512), k < 256 (host table: correctly rounded, independent of --use_fast_math)
  {
    const float2 tw = reinterpret_cast<const float2*>(p.twiddle)[t];
    tw_re[t] = tw.x;
    tw_im[t] = tw.y;
  }
  // frame (scaled to 16-bit range), mean over the 400 samples
  float v0 = (t < FB_WIN) ? w[t] * p.in_scale : 0.f;
  float v1 = (t + FB_THREADS < FB_WIN) ? w[t + FB_THREADS] * p.in_scale : 0.f;
  float s = warp_sum(v0 + v1);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < FB_THREADS / 32; ++i) tot += red[i];
  const float mean = __fdiv_rn(tot, (float)FB_WIN);
  s_x[t] = v0 - mean;                       // remove_dc_offset
  s_x[t + FB_THREADS] = (t + FB_THREADS < FB_WIN) ? v1 - mean : 0.f;
  __syncthreads();
  // pre-emphasis (x[j] - 0.97 x[max(j-1,0)]), Povey window, zero pad, bit-reversed placement for the DIT FFT
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = t + h * FB_THREADS;
    float y = 0.f;
    if (j < FB_WIN) y = (s_x[j] - 0.97f * s_x[j > 0 ? j - 1 : 0]) * p.window[j];
    const int r = (int)(__brev((unsigned)j) >> (32 - FB_LOGN));
    s_re[r] = y;
    s_im[r] = 0.f;
  }
  __syncthreads();
  // 9 radix-2 stages, one butterfly per thread per stage
#pragma unroll
  for (int st = 0; st < FB_LOGN; ++st) {
    const int half = 1 << st;
    const int pos = t & (half - 1);
    const int i = ((t >> st) << (st + 1)) + pos;
    const int j = i + half;
    const int k = pos << (FB_LOGN - 1 - st);
    const float wr = tw_re[k], wi = tw_im[k];
    const float xr = s_re[j], xi = s_im[j];
    const float tr = xr * wr - xi * wi, ti = xr * wi + xi * wr;
    const float ur = s_re[i], ui = s_im[i];
    s_re[i] = ur + tr; s_im[i] = ui + ti;
    s_re[j] = ur - tr; s_im[j] = ui - ti;
    __syncthreads();
  }
  // power spectrum as the reference forms it: |X| then squared (bins 0..256)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = t + h * FB_THREADS;
    if (k <= FB_N / 2) {
      const float a = __fsqrt_rn(s_re[k] * s_re[k] + s_im[k] * s_im[k]);
      s_x[k] = a * a;
    }
  }
  __syncthreads();
  if (t < p.n_mel) {
    const int b0 = p.mel_start[t], o0 = p.mel_off[t], n = p.mel_off[t + 1] - o0;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += s_x[b0 + i] * p.mel_w[o0 + i];
    // torch.finfo(float32).eps floor; the log goes through fp64 because --use_fast_math turns logf into lg2.approx * ln2
    const float e = (float)log((double)fmaxf(acc, 1.1920928955078125e-07f));
    p.out[((size_t)seg * p.n_frames + frame) * p.n_mel + t] = (e - p.mean) * p.inv_scale;
  }
}

// uint8 [n, H, W, C=3] -> bf16 patch rows [n * gh * gw, ld]; column order (c, ky, kx) as crab_patchify.
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ out, int ld, int Hh,
                                   int Ww, int p, int gh, int gw, float m0, float m1, float m2, float is0, float is1,
                                   float is2, float rescale) {
  const int prow = blockIdx.x;
  const int n = prow / (gh * gw), r = prow % (gh * gw);
  const int py = r / gw, px = r % gw;
  const int kcols = 3 * p * p;
  __nv_bfloat16* o = out + (size_t)prow * ld;
  for (int k = threadIdx.x; k < ld; k += blockDim.x) {
    float v = 0.f;
    if (k < kcols) {
      const int c = k / (p * p), rem = k % (p * p);
      const int ky = rem / p, kx = rem % p;
      const float u = (float)img[(((size_t)n * Hh + (py * p + ky)) * Ww + (px * p + kx)) * 3 + c];
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
      const float istd = c == 0 ? is0 : (c == 1 ? is1 : is2);
      v = (u * rescale - mean) * istd;
    }
    o[k] = __float2bfloat16_rn(v);
  }
}

// uint8 [n, H, W, 3] -> fp32 [n, 3, H, W]  ((u / 255 - mean) / std)
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, size_t n_pix, int HW,
                                    float m0, float m1, float m2, float is0, float is1, float is2, float rescale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index over n * H * W
  if (i >= n_pix) return;
  const size_t n = i / HW, r = i % HW;
  const uint8_t* s = img + i * 3;
  float* o = out + n * 3 * HW + r;
  o[0] = ((float)s[0] * rescale - m0) * is0;
  o[HW] = ((float)s[1] * rescale - m1) * is1;
  o[2 * (size_t)HW] = ((float)s[2] * rescale - m2) * is2;
}

// One separable pass of Pillow's 8-bit resampler (src/libImaging/Resample.c: ImagingResampleHorizontal_8bpc /
// ImagingResampleVertical_8bpc — the routine HF's CLIPImageProcessor of the reference's pinned transformers calls through
// PIL.Image.resize(..., BICUBIC)): out = clip8((2^21 + sum_k in[xmin + k] * kk[k]) >> 22) with integer coefficients built on
// the host.  uint8 [n, H, W, 3] interleaved; `axis` 1 = along x, 0 = along y; output index i uses coefficient row out0 + i, so the
// centre crop that follows the resize costs nothing (only the kept window is computed).  Integer arithmetic: bit-exact.
__global__ void resample_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int in_h, int in_w,
                                   int out_h, int out_w, int axis, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, int ksize, int out0) {
  const long long total = (long long)n * out_h * out_w * 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % 3);
  const int x = (int)((i / 3) % out_w);
  const int y = (int)((i / (3LL * out_w)) % out_h);
  const int img = (int)(i / (3LL * out_w * out_h));
  const int o = out0 + (axis ? x : y);
  const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
  const int* k = kk + (size_t)o * ksize;
  int ss = 1 << 21;
  if (axis) {
    const uint8_t* p = in + (((size_t)img * in_h + y) * in_w + lo) * 3 + c;
    for (int t = 0; t < cnt; ++t) ss += (int)p[(size_t)t * 3] * k[t];
  } else {
    const uint8_t* p = in + (((size_t)img * in_h + lo) * in_w + x) * 3 + c;
    for (int t = 0; t < cnt; ++t) ss += (int)p[(size_t)t * in_w * 3] * k[t];
  }
  ss >>= 22;
  out[i] = (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

}  // namespace crab

using namespace crab;

extern "C" int crab_kaldi_fbank(const float* wave, int64_t wave_stride, int n_seg, int n_samples, const float* window,
                                const float* twiddle, const int* mel_start, const int* mel_off, const float* mel_w, int n_mel, float in_scale,
                                float mean, float std2, float* out, void* stream) {
  CRAB_REQUIRE(wave && window && twiddle && mel_start && mel_off && mel_w && out, "crab_kaldi_fbank: null pointer");
  CRAB_REQUIRE((uintptr_t)twiddle % 8 == 0, "crab_kaldi_fbank: twiddle table must be 8-byte aligned");
  CRAB_REQUIRE(n_mel > 0 && n_mel <= FB_THREADS, "crab_kaldi_fbank: n_mel must be in 1..%d", FB_THREADS);
  CRAB_REQUIRE(n_samples >= FB_WIN, "crab_kaldi_fbank: need at least %d samples per segment (got %d)", FB_WIN, n_samples);
  CRAB_REQUIRE(std2 != 0.f, "crab_kaldi_fbank: zero scale");
  if (n_seg <= 0) return CRAB_OK;
  FbankParams p;
  p.wave = wave; p.wave_stride = wave_stride; p.n_seg = n_seg;
  p.n_frames = 1 + (n_samples - FB_WIN) / FB_SHIFT;  // snip_edges = True
  p.n_mel = n_mel; p.window = window; p.twiddle = twiddle; p.mel_start = mel_start; p.mel_off = mel_off; p.mel_w = mel_w; p.out = out;
  p.in_scale = in_scale; p.mean = mean; p.inv_scale = 1.0f / std2;
  fbank_kernel<<<(unsigned)(n_seg * p.n_frames), FB_THREADS, 0, (cudaStream_t)stream>>>(p);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_fbank_num_frames(int n_samples, int* n_frames) {
  CRAB_REQUIRE(n_frames != nullptr, "crab_fbank_num_frames: null");
  *n_frames = n_samples < FB_WIN ? 0 : 1 + (n_samples - FB_WIN) / FB_SHIFT;
  return CRAB_OK;
}

extern "C" int crab_patchify_u8(const void* images_hwc, void* out, int ld_out, int n_img, int H, int W, int patch,
                                const float* mean3, const float* std3, float rescale, void* stream) {
  CRAB_REQUIRE(images_hwc && out && mean3 && std3 && patch > 0, "crab_patchify_u8: bad args");
  CRAB_REQUIRE(ld_out >= 3 * patch * patch && ld_out % 8 == 0, "crab_patchify_u8: ld_out too small / unaligned");
  CRAB_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "crab_patchify_u8: zero std");
  const int gh = H / patch, gw = W / patch;
  const int rows = n_img * gh * gw;
  if (rows <= 0) return CRAB_OK;
  patchify_u8_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(images_hwc),
                                                             reinterpret_cast<__nv_bfloat16*>(out), ld_out, H, W, patch, gh, gw,
                                                             mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1],
                                                             1.f / std3[2], rescale);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_normalize_u8(const void* images_hwc, float* out_nchw, int n_img, int H, int W, const float* mean3,
                                 const float* std3, float rescale, void* stream) {
  CRAB_REQUIRE(images_hwc && out_nchw && mean3 && std3, "crab_normalize_u8: bad args");
  CRAB_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "crab_normalize_u8: zero std");
  const size_t n_pix = (size_t)n_img * H * W;
  if (n_pix == 0) return CRAB_OK;
  normalize_u8_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(images_hwc), out_nchw, n_pix, H * W, mean3[0], mean3[1], mean3[2], 1.f / std3[0],
      1.f / std3[1], 1.f / std3[2], rescale);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_resample_u8(const void* in_hwc, void* out_hwc, int n_img, int in_h, int in_w, int out_h, int out_w, int axis,
                                const int* bounds, const int* kk, int ksize, int out0, void* stream) {
  CRAB_REQUIRE(in_hwc && out_hwc && bounds && kk && ksize > 0 && out0 >= 0, "crab_resample_u8: bad args");
  CRAB_REQUIRE(axis == 0 || axis == 1, "crab_resample_u8: axis must be 0 (rows) or 1 (columns)");
  CRAB_REQUIRE(axis == 1 ? in_h == out_h : in_w == out_w, "crab_resample_u8: the other axis must keep its size");
  const long long total = (long long)n_img * out_h * out_w * 3;
  if (total <= 0) return CRAB_OK;
  resample_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(in_hwc), reinterpret_cast<uint8_t*>(out_hwc), n_img, in_h, in_w, out_h, out_w, axis, bounds,
      kk, ksize, out0);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}
