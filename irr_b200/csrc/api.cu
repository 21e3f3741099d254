// Error plumbing and device queries for the irr_b200 C ABI (include/irr_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace irr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_arg(const char* fn, const char* what) {
  set_error("%s: invalid argument: %s", fn, what);
  return IRR_E_ARG;
}

// The reference launcher checks cudaGetLastError() after its launches and reports failure
// (correlation_cuda_kernel.cu:383-392); same contract, but the code is returned instead of printf'd.
int check_launch(const char* fn) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", fn, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool make_nchw_map(CUtensorMap* map, const void* base, long long bs, int B, int C, int H, int W, int bw, int bh,
                   int bc, int pitch, int elem_bytes) {
  EncodeTiledFn enc = encode_tiled();
  const int P = pitch > 0 ? pitch : W;   // row pitch in elements; the map's inner dimension stays W: columns >= W read 0
  const int E = elem_bytes, A = 16 / E;  // elements per 16 bytes
  if (E != 4 && E != 2) return false;
  if (!enc || (P % A) != 0 || P < W || (bs % A) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  if (bw > 256 || bh > 256 || bc > 256 || ((bw * E) % 16) != 0) return false;
  memset(map, 0, sizeof(*map));
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)P * E, (cuuint64_t)H * P * E, (cuuint64_t)bs * E};
  cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, E == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
             strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace irr

extern "C" {

int irr_abi_version(void) { return IRR_ABI_VERSION; }

const char* irr_last_error(void) { return irr::g_err; }

int irr_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    irr::set_error("irr_device_info: %s", cudaGetErrorString(e));
    return (int)e;
  }
  int n = 0, ma = 0, mi = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  return 0;
}

}  // extern "C"
