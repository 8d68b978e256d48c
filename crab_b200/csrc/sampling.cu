// Token choice and loss on fp32 logits rows: temperature / top-k / top-p sampling (what HF `generate(do_sample=True)` does
// with LLaMA-2-chat's generation_config: T 0.6, top-p 0.9, top-k 50) and the shifted-label cross-entropy of
// UnifiedForCausalLM.forward(labels=...).  One block per row, the row is streamed from global memory (it is L2-resident:
// 128 KB .. 600 KB), fp32 throughout, every reduction in a fixed order (deterministic).
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace crab {

static constexpr int SP_THREADS = 1024;

__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : -INFINITY;
  if (w == 0) {
    r = warp_max(r);
    if (l == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (l == 0) sh[0] = r;
  }
  __syncthreads();
  r = sh[0];
  __syncthreads();
  return r;
}
// order-preserving map float -> uint32 (larger float = larger key)
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Radix select over the keys of one row: returns the largest threshold key t such that  measure({key >= t}) >= target,
// where measure is a COUNT (top-k) or a probability MASS (top-p), restricted to keys >= floor_key.  Four 8-bit passes; each pass
// histograms the next digit of the keys that match the prefix found so far.  Masses are accumulated as 2^-40 fixed point in
// 64-bit integers, so the shared-memory atomics commute and the result does not depend on the order of arrival.
static constexpr float SP_FIX = 1099511627776.0f;   // 2^40
__device__ __forceinline__ unsigned long long mass_fix(float s, float mx) { return (unsigned long long)(__expf(s - mx) * SP_FIX); }

template <bool MASS>
__device__ uint32_t radix_select(const float* row, int V, float inv_t, float mx, uint32_t floor_key, unsigned long long target,
                                 unsigned long long* hist /*256*/, unsigned long long* sh2) {
  uint32_t prefix = 0, mask = 0;
  unsigned long long acc_above = 0;   // measure of keys strictly above the current prefix range
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0ull;
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float s = row[i] * inv_t;
      const uint32_t k = fkey(s);
      if (k >= floor_key && (k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], MASS ? mass_fix(s, mx) : 1ull);
    }
    __syncthreads();
    // walk the digits from the top; stop at the digit with which the running measure reaches the target
    if (threadIdx.x == 0) {
      unsigned long long run = acc_above;
      int d = 255;
      for (; d > 0; --d) {
        if (run + hist[d] >= target) break;
        run += hist[d];
      }
      sh2[0] = run;
      sh2[1] = (unsigned long long)d;
    }
    __syncthreads();
    acc_above = sh2[0];
    const uint32_t d = (uint32_t)sh2[1];
    __syncthreads();
    prefix |= d << shift;
    mask |= 255u << shift;
  }
  return prefix;
}

// next[row] ~ softmax(filter(logits[row] / T)) by inverse CDF over the kept tokens in index order, driven by u[row] in [0, 1).
// filter = HF's TopKLogitsWarper then TopPLogitsWarper (keep the smallest set of most probable tokens whose mass reaches top_p,
// at least one token); tokens tied with the threshold are all kept.
__global__ void __launch_bounds__(SP_THREADS) sample_kernel(const float* __restrict__ logits, int ld, int V, float inv_t, int top_k,
                                                            float top_p, const float* __restrict__ u, int64_t* __restrict__ out) {
  __shared__ float sh[32];
  __shared__ unsigned long long hist[256];
  __shared__ unsigned long long sh2[2];
  __shared__ unsigned long long zsum;
  __shared__ float chunk_mass[SP_THREADS];
  const float* row = logits + (size_t)blockIdx.x * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, row[i] * inv_t);
  mx = block_reduce_max(mx, sh);
  uint32_t floor_key = 0;
  if (top_k > 0 && top_k < V) floor_key = radix_select<false>(row, V, inv_t, mx, 0u, (unsigned long long)top_k, hist, sh2);
  uint32_t thr = floor_key;
  if (top_p < 1.0f) {
    if (threadIdx.x == 0) zsum = 0ull;
    __syncthreads();
    unsigned long long z = 0ull;
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
      const float s = row[i] * inv_t;
      if (fkey(s) >= floor_key) z += mass_fix(s, mx);
    }
    atomicAdd(&zsum, z);
    __syncthreads();
    const unsigned long long target = (unsigned long long)((double)top_p * (double)zsum);
    thr = radix_select<true>(row, V, inv_t, mx, floor_key, target, hist, sh2);
  }
  // inverse CDF in index order over {key >= thr}: contiguous chunk per thread, block scan of the chunk masses
  const int per = (V + blockDim.x - 1) / blockDim.x;
  const int i0 = threadIdx.x * per, i1 = min(V, i0 + per);
  float m = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float s = row[i] * inv_t;
    if (fkey(s) >= thr) m += __expf(s - mx);
  }
  chunk_mass[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int t = 0; t < (int)blockDim.x; ++t) total += chunk_mass[t];
    const float r = fminf(u[blockIdx.x], 0.99999994f) * total;
    float run = 0.f;
    int t = 0;
    for (; t < (int)blockDim.x - 1; ++t) {
      if (run + chunk_mass[t] > r) break;
      run += chunk_mass[t];
    }
    sh[0] = run;
    sh[1] = __int_as_float(t);
    sh[2] = r;
  }
  __syncthreads();
  if ((int)threadIdx.x == __float_as_int(sh[1])) {
    float run = sh[0];
    const float r = sh[2];
    int pick = -1, last_kept = -1;
    for (int i = i0; i < i1; ++i) {
      const float s = row[i] * inv_t;
      if (fkey(s) >= thr) {
        last_kept = i;
        run += __expf(s - mx);
        if (run > r) { pick = i; break; }
      }
    }
    if (pick < 0) pick = last_kept;       // rounding at the very end of the chunk
    if (pick < 0) {                       // chunk without kept tokens (only possible for the last chunk by rounding): arg-max fallback
      float best = -INFINITY;
      for (int i = 0; i < V; ++i) if (row[i] > best) { best = row[i]; pick = i; }
    }
    out[blockIdx.x] = pick;
  }
}

// loss[row] = logsumexp(logits[row, :V]) - logits[row, label[row]]   (0 and not counted when label < 0, i.e. ignore_index -100)
__global__ void __launch_bounds__(256) xent_kernel(const float* __restrict__ logits, int ld, int V, const int64_t* __restrict__ labels,
                                                   float* __restrict__ loss) {
  __shared__ float sh[32];
  const int64_t lab = labels[blockIdx.x];
  if (lab < 0 || lab >= V) {
    if (threadIdx.x == 0) loss[blockIdx.x] = 0.f;
    return;
  }
  const float* row = logits + (size_t)blockIdx.x * ld;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) mx = fmaxf(mx, row[i]);
  mx = block_reduce_max(mx, sh);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(row[i] - mx);
  s = block_reduce_sum(s, sh);
  if (threadIdx.x == 0) loss[blockIdx.x] = logf(s) + mx - row[lab];
}

}  // namespace crab

using namespace crab;

extern "C" int crab_sample_top_k_top_p(const float* logits, int ld, int rows, int V, float temperature, int top_k, float top_p,
                                       const float* u, int64_t* out, void* stream) {
  CRAB_REQUIRE(logits && u && out && V > 0 && ld >= V, "crab_sample_top_k_top_p: bad args");
  CRAB_REQUIRE(temperature > 0.f && top_p > 0.f && top_p <= 1.0f && top_k >= 0, "crab_sample_top_k_top_p: temperature > 0, 0 < top_p <= 1, top_k >= 0");
  if (rows <= 0) return CRAB_OK;
  sample_kernel<<<rows, SP_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, V, 1.0f / temperature, top_k, top_p, u, out);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}

extern "C" int crab_cross_entropy(const float* logits, int ld, int rows, int V, const int64_t* labels, float* loss, void* stream) {
  CRAB_REQUIRE(logits && labels && loss && V > 0 && ld >= V, "crab_cross_entropy: bad args");
  if (rows <= 0) return CRAB_OK;
  xent_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(logits, ld, V, labels, loss);
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}
