// Shared helpers for the irr_b200 kernels (sm_100a).  Internal header — the public ABI is include/irr_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/irr_b200.h"

namespace irr {

void set_error(const char* fmt, ...);
int fail_arg(const char* fn, const char* what);
int check_launch(const char* fn);
int sm_count();
int current_device();  // cudaGetDevice(), -1 on failure

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE function attribute: remember, per device, how far it
// has been raised (one process may drive several GPUs — ADVICE r1).
struct SmemAttrCache { size_t bytes[64]; };
void set_error(const char* fmt, ...);
template <typename Kern>
static inline int ensure_dyn_smem(Kern kern, size_t smem, SmemAttrCache& cache, const char* fn) {
  const int dev = current_device();
  const bool cacheable = dev >= 0 && dev < 64;
  if (cacheable && smem <= cache.bytes[dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%zu): %s", fn, smem, cudaGetErrorString(e));
    return (int)e;
  }
  if (cacheable) cache.bytes[dev] = smem;
  return 0;
}

#define IRR_REQUIRE(cond, fn, what) \
  do {                              \
    if (!(cond)) return ::irr::fail_arg(fn, what); \
  } while (0)

static inline cudaStream_t as_stream(irr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// cuTensorMapEncodeTiled, fetched at run time through cudaGetDriverEntryPoint (the library does not link libcuda);
// nullptr when the driver does not provide it.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled();
// fp32 NCHW channel-slice view (W, H, C, B; batch stride bs elements) -> tiled tensor map with the given box
// (elements, innermost first).  Out-of-bounds box elements read as zero.  Needs W % 4 == 0, bs % 4 == 0, 16-byte base.
// elem_bytes = 2: the same view over bf16 storage (strides / pitch in bf16 elements; needs pitch % 8 == 0, bs % 8 == 0).
bool make_nchw_map(CUtensorMap* map, const void* base, long long bs, int B, int C, int H, int W, int bw, int bh, int bc,
                   int pitch = 0, int elem_bytes = 4);

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

// ---------------------------------------------------------------------------------------------------------
// mbarrier / shared-address helpers (producer-consumer pipelines inside one CTA)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA: 4-D tiled box load global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// Same, for long waits off the critical path (epilogue / loader roles): back off between probes so the spinning warp
// does not steal issue slots from the producer warps that share its scheduler.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) __nanosleep(128);
  } while (!done);
}

// ---------------------------------------------------------------------------------------------------------
// Sampling-grid arithmetic of WarpingLayer (models/pwc_modules.py:119-127) + grid_sampler_2d's unnormalise
// (ATen/native/cuda/GridSampler.cuh:22-31).  Every op is an explicitly rounded intrinsic so ptxas cannot
// contract or reassociate: the validity mask `sum_w >= 1.0f` (pwc_modules.py:131) is bit-sensitive.
struct GridArgs {
  const float* lin_x;  // W entries (host torch.linspace(-1,1,W)) or nullptr
  const float* lin_y;  // H entries or nullptr
  float den_x, den_y;  // max(W_im-1,1), max(H_im-1,1)
  float div_flow;
  float rcp_x, rcp_y, rcp_div;  // 1.0f/den (host-rounded) for IRR_GRID_RECIP_MUL
  float step_x, step_y;         // 2/(W-1), 2/(H-1) for the in-kernel linspace fallback
  int recip;
};

static inline GridArgs make_grid_args(const float* lin_x, const float* lin_y, int H, int W, int H_im, int W_im,
                                      float div_flow, int flags) {
  GridArgs g;
  g.lin_x = lin_x;
  g.lin_y = lin_y;
  g.den_x = (float)(W_im - 1 > 1 ? W_im - 1 : 1);
  g.den_y = (float)(H_im - 1 > 1 ? H_im - 1 : 1);
  g.div_flow = div_flow;
  volatile float one = 1.0f;  // keep these as true fp32 divisions on the host
  g.rcp_x = one / g.den_x;
  g.rcp_y = one / g.den_y;
  g.rcp_div = one / div_flow;
  g.step_x = W > 1 ? 2.0f / (float)(W - 1) : 0.f;
  g.step_y = H > 1 ? 2.0f / (float)(H - 1) : 0.f;
  g.recip = (flags & IRR_GRID_RECIP_MUL) ? 1 : 0;
  return g;
}

__device__ __forceinline__ float grid_base(const float* lin, float step, int i, int n) {
  if (lin) return __ldg(lin + i);
  if (n <= 1) return -1.0f;  // torch.linspace(-1, 1, 1) == [-1]
  return __fadd_rn(-1.0f, __fmul_rn(step, (float)i));
}

// Correctly rounded a / b for a CONSTANT divisor b whose correctly rounded reciprocal rb = RN(1/b) was computed on the
// host: q = RN(a*rb) is faithful, and each step q += RN(a - q*b) * rb (the residual is exact in one FMA) lands on RN(a/b)
// (Markstein).  Five dependent FMA-pipe instructions instead of __fdiv_rn's ~15 + FCHK + slow-path branch — the grid
// arithmetic of the warp is bit-sensitive (hard mask), so "close" is not good enough: the recipe is checked
// against IEEE division on 56 M random operands (tests/test_oracle.py::test_constant_division_recipe) and bitwise
// against torch through the warp tests.  Inf / NaN inputs give NaN instead of Inf: both mask the pixel out.
__device__ __forceinline__ float div_const_rn(float a, float b, float rb) {
  float q = __fmul_rn(a, rb);
  float r = __fmaf_rn(-q, b, a);
  q = __fmaf_rn(r, rb, q);
  r = __fmaf_rn(-q, b, a);
  return __fmaf_rn(r, rb, q);
}

// Unnormalised source coordinates for output pixel (y, x) with flow (u, v).
__device__ __forceinline__ void sample_coords(const GridArgs& g, float u, float v, int x, int y, int W, int H,
                                              float& ix, float& iy) {
  float fx, fy;
  if (g.recip) {
    fx = __fmul_rn(__fmul_rn(__fmul_rn(u, 2.0f), g.rcp_x), g.rcp_div);
    fy = __fmul_rn(__fmul_rn(__fmul_rn(v, 2.0f), g.rcp_y), g.rcp_div);
  } else {
    fx = div_const_rn(div_const_rn(__fmul_rn(u, 2.0f), g.den_x, g.rcp_x), g.div_flow, g.rcp_div);
    fy = div_const_rn(div_const_rn(__fmul_rn(v, 2.0f), g.den_y, g.rcp_y), g.div_flow, g.rcp_div);
  }
  float gx = __fadd_rn(grid_base(g.lin_x, g.step_x, x, W), fx);
  float gy = __fadd_rn(grid_base(g.lin_y, g.step_y, y, H), fy);
  ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(W - 1));
  iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(H - 1));
}

// Bilinear taps: integer corner (x0,y0), the four weights with out-of-bounds taps zeroed, and the hard mask.
struct Taps {
  int x0, y0;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1), already zero where out of bounds
  float mask;                // 1.0f or 0.0f
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int W, int H) {
  Taps t;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float fx1 = __fadd_rn(fx0, 1.0f), fy1 = __fadd_rn(fy0, 1.0f);
  float ax = __fsub_rn(fx1, ix), bx = __fsub_rn(ix, fx0);
  float ay = __fsub_rn(fy1, iy), by = __fsub_rn(iy, fy0);
  float nw = __fmul_rn(ax, ay), ne = __fmul_rn(bx, ay), sw = __fmul_rn(ax, by), se = __fmul_rn(bx, by);
  // bounds in float first (robust to inf / NaN / huge coordinates), then to int
  bool x0ok = fx0 >= 0.0f && fx0 <= (float)(W - 1);
  bool x1ok = fx1 >= 0.0f && fx1 <= (float)(W - 1);
  bool y0ok = fy0 >= 0.0f && fy0 <= (float)(H - 1);
  bool y1ok = fy1 >= 0.0f && fy1 <= (float)(H - 1);
  t.w00 = (x0ok && y0ok) ? nw : 0.0f;
  t.w01 = (x1ok && y0ok) ? ne : 0.0f;
  t.w10 = (x0ok && y1ok) ? sw : 0.0f;
  t.w11 = (x1ok && y1ok) ? se : 0.0f;
  float m = __fadd_rn(__fadd_rn(__fadd_rn(t.w00, t.w01), t.w10), t.w11);  // grid_sample(ones): 0 + nw + ne + sw + se
  t.mask = (m >= 1.0f) ? 1.0f : 0.0f;
  bool any = (x0ok || x1ok) && (y0ok || y1ok);
  t.x0 = any ? (int)fx0 : 0;
  t.y0 = any ? (int)fy0 : 0;
  if (!any) t.w00 = t.w01 = t.w10 = t.w11 = 0.0f;
  return t;
}

// Gather one channel plane `p` (H x W) with taps t (weights of out-of-bounds taps are zero, so clamp addresses).
// `P` = row pitch of the plane in elements (>= W).
__device__ __forceinline__ float gather_bilinear(const float* __restrict__ p, const Taps& t, int W, int H, int P) {
  int x0 = min(max(t.x0, 0), W - 1), x1 = min(max(t.x0 + 1, 0), W - 1);
  int y0 = min(max(t.y0, 0), H - 1), y1 = min(max(t.y0 + 1, 0), H - 1);
  float acc = __fmul_rn(__ldg(p + (size_t)y0 * P + x0), t.w00);  // same tap order as grid_sampler_2d_kernel
  acc = fmaf(__ldg(p + (size_t)y0 * P + x1), t.w01, acc);
  acc = fmaf(__ldg(p + (size_t)y1 * P + x0), t.w10, acc);
  acc = fmaf(__ldg(p + (size_t)y1 * P + x1), t.w11, acc);
  return acc;
}

}  // namespace irr
