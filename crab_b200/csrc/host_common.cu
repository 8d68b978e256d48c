// Error reporting, device probe and TMA tensor-map encoding for the C ABI.
#include "host_common.h"
#include <stdlib.h>

namespace crab {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

static thread_local int g_pdl = -1;   // launch policy of the calling thread: two engines on two threads do not see each other
int pdl_mask() {
  if (g_pdl < 0) {
    const char* e = getenv("CRAB_PDL");
    g_pdl = (e != nullptr) ? (atoi(e) & 63) : 0;
  }
  return g_pdl;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so the library does not link against libcuda directly
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(CRAB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};  // bytes, dimension 1
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(CRAB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) base=%p rows=%llu cols=%llu ld=%llu box=%ux%u",
                     (int)r, base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows,
                     box_cols);
  return 0;
}

static unsigned long long* g_trace_buf = nullptr;
static int g_trace_slots = 0, g_trace_ctas = 0, g_trace_next = 0;
unsigned long long* next_trace_slot(int ctas) {
  if (!g_trace_buf || g_trace_next >= g_trace_slots) return nullptr;
  const int slot = g_trace_next++;
  if (ctas > g_trace_ctas) return nullptr;
  return g_trace_buf + (size_t)slot * g_trace_ctas * 16;
}

}  // namespace crab

extern "C" int crab_debug_trace(void* buf, int nslots, int ctas_per_slot) {
  crab::g_trace_buf = reinterpret_cast<unsigned long long*>(buf);
  crab::g_trace_slots = buf ? nslots : 0;
  crab::g_trace_ctas = ctas_per_slot;
  crab::g_trace_next = 0;
  return CRAB_OK;
}

extern "C" const char* crab_last_error(void) { return crab::last_error_buf(); }

extern "C" int crab_version(void) { return 1; }

extern "C" int crab_set_pdl(int mask) {
  crab::g_pdl = mask & 63;
  return CRAB_OK;
}

extern "C" int crab_init(int dev) {
  using namespace crab;
  int n = 0;
  CRAB_CHECK_CUDA(cudaGetDeviceCount(&n));
  CRAB_REQUIRE(dev >= 0 && dev < n, "crab_init: device %d out of range (%d devices)", dev, n);
  int major = 0, minor = 0;
  CRAB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CRAB_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return set_error(CRAB_ERR_ARCH, "crab_init: device %d is sm_%d%d; this library is sm_100a only", dev, major, minor);
  return CRAB_OK;
}
