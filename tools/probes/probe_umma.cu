// Probe (diagnostics, not product): tcgen05.mma SS-mode issue cost per shape and the TMEM lane layout of M=64 accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I crab_b200/csrc tools/probes/probe_umma.cu -o /tmp/probe_umma
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace crab;

__global__ void __launch_bounds__(128, 1) probe(int M, int N, int iters, float* out, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A: 128 rows x 64 K bf16, SW128 K-major; row r holds value (r+1) in column k=0 only.   B: 256 rows x 64, row n holds 1.0 at k=0.
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gen);
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(gen + 16384);
  for (int i = threadIdx.x; i < 128 * 64; i += 128) A[i] = __float2bfloat16_rn(0.f);
  for (int i = threadIdx.x; i < 256 * 64; i += 128) B[i] = __float2bfloat16_rn(0.f);
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += 128) {
    // element (r, k=0): chunk 0 stored at chunk position 0 ^ (r & 7)
    A[r * 64 + ((0 ^ (r & 7)) * 8)] = __float2bfloat16_rn((float)(r + 1));
  }
  for (int n = threadIdx.x; n < 256; n += 128) B[n * 64 + ((0 ^ (n & 7)) * 8)] = __float2bfloat16_rn(1.0f);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tslot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t idesc = make_idesc_bf16_f32(M, N);
  if (threadIdx.x == 0) {
    const uint64_t da = make_sdesc_sw128(base), db = make_sdesc_sw128(base + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, da + 2u * k, db + 2u * k, idesc, (i > 0) | (k > 0));
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    cycles[0] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  // dump column 0 of every TMEM lane: lane l of warp w = TMEM lane 32 w + l
  uint32_t r[32];
  tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  tmem_ld_wait();
  out[threadIdx.x] = __uint_as_float(r[0]) / (float)iters;   // = (row + 1) of the row that lives in this lane, 0 if unused
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 128 * 4); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int shapes[][2] = {{128, 32}, {128, 128}, {128, 256}, {64, 32}, {64, 128}, {64, 256}};
  for (auto& s : shapes) {
    const int iters = 512;
    cudaMemset(out, 0, 128 * 4);
    probe<<<1, 128, 64 * 1024>>>(s[0], s[1], iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", s[0], s[1], cudaGetErrorString(e)); return 1; }
    std::vector<float> h(128); long long c;
    cudaMemcpy(h.data(), out, 128 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d: %.1f cycles per MMA (K=16) over %d x 4 back-to-back MMAs; lane -> row+1:", s[0], s[1], (double)c / (iters * 4), iters);
    for (int l = 0; l < 128; ++l) { if (l % 32 == 0) printf("\n    lanes %3d..: ", l); printf("%g ", h[l]); }
    printf("\n");
  }
  return 0;
}
