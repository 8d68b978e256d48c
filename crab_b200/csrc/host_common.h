// Host-side helpers shared by the C-ABI translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/crab_b200.h"

namespace crab {

// Thread-local last error string, returned by crab_last_error().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define CRAB_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      return crab::set_error(CRAB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                                    \
  } while (0)

#define CRAB_REQUIRE(cond, ...)                                         \
  do {                                                                  \
    if (!(cond)) return crab::set_error(CRAB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Encode a 2-D bf16 row-major tensor [rows, cols] (row stride ld elements) with a [box_rows, 64]-element box and
// 128-byte swizzle. Returns 0 or a negative error code.
int encode_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols);

int sm_count();

// cudaFuncSetAttribute is per device: `once_per_device(flags)` is true the first time it is called for the current device with
// this flag array (callers keep one `static DeviceOnce` per kernel), so a process that drives several GPUs sets the attribute on each.
struct DeviceOnce { bool done[64] = {false}; };
inline bool first_on_device(DeviceOnce& f) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (f.done[dev]) return false;
  f.done[dev] = true;
  return true;
}

// Programmatic dependent launch (env CRAB_PDL or crab_set_pdl): a bit mask over kernel classes.  A kernel whose class
// bit is set is launched with cudaLaunchAttributeProgrammaticStreamSerialization, i.e. its CTAs may become resident as
// soon as every CTA of the PREVIOUS kernel in the stream has executed griddepcontrol.launch_dependents (all decode-chain
// kernels do so at their top) and block in griddepcontrol.wait until that kernel has completed.  The mask is read at
// launch time, so the host can change it between two launches (the engine does, see engine._decoder_layers).
enum PdlClass { PDL_GEMM = 1, PDL_ROW = 2, PDL_ROPE = 4, PDL_ATTN = 8, PDL_LIGHT = 16, PDL_ATTN_LATE = 32 /* modifier */ };
int pdl_mask();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Diagnostics (tools/trace_skinny.py): when crab_debug_trace() has armed a buffer, every traced launch gets the next slot of
// [ctas][8] globaltimer stamps; nullptr otherwise (the kernels then skip every stamp).
unsigned long long* next_trace_slot(int ctas);

}  // namespace crab
