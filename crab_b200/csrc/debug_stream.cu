// Diagnostic only (not on the product path): pure HBM -> smem streaming with the same ring the skinny GEMM uses, no
// MMA and no epilogue, to measure what the load path alone sustains for a given (CTAs, stages, chunk) configuration.
#include "host_common.h"
#include "ptx.cuh"

namespace crab {

__global__ void __launch_bounds__(64, 1) debug_stream_kernel(const uint8_t* __restrict__ src, long long chunks, int chunk_bytes,
                                                            int stages, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + stages * chunk_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long G = gridDim.x;
  const long long lo = blockIdx.x * chunks / G, hi = (blockIdx.x + 1) * chunks / G;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (stages + s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {
    uint32_t st = 0, ph = 0;
    for (long long f = lo; f < hi; ++f) {
      mbar_wait(bar + 8 * (stages + st), ph ^ 1);
      mbar_arrive_expect_tx(bar + 8 * st, chunk_bytes);
      bulk_load_1d_hint(base + st * chunk_bytes, src + f * chunk_bytes, chunk_bytes, bar + 8 * st, kEvictFirst);
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t st = 0, ph = 0;
    unsigned long long acc = 0;
    for (long long f = lo; f < hi; ++f) {
      mbar_wait(bar + 8 * st, ph);
      uint32_t v;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(base + st * chunk_bytes));
      acc += v;
      mbar_arrive(bar + 8 * (stages + st));
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1; }
    }
    if (acc == 0x123456789ull) *sink = acc;
  }
}

}  // namespace crab

// declared here, not in include/crab_b200.h: this file is only built into _lib/libcrab_diag.so (crab_b200/build.py::build_diag)
extern "C" int crab_debug_stream(const void* src, int64_t bytes, int chunk_bytes, int stages, int ctas, void* sink, void* stream) {
  using namespace crab;
  CRAB_REQUIRE(src && sink && chunk_bytes % 1024 == 0 && stages >= 1 && ctas >= 1, "crab_debug_stream: bad args");
  const int smem = stages * chunk_bytes + 1024 + 16 * stages + 64;
  CRAB_REQUIRE(smem <= 227 * 1024, "crab_debug_stream: smem");
  CRAB_CHECK_CUDA(cudaFuncSetAttribute(debug_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  debug_stream_kernel<<<ctas, 64, smem, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(src), bytes / chunk_bytes, chunk_bytes,
                                                               stages, reinterpret_cast<unsigned long long*>(sink));
  CRAB_CHECK_CUDA(cudaGetLastError());
  return CRAB_OK;
}
