// Standalone WarpingLayer (models/pwc_modules.py:115-133): bilinear gather + rounding-exact hard mask, optionally
// fused with the `a - warp(b)` the refinement nets consume (IRR_PWC.py:132-133,144-145) and a batch rotation so the
// forward and backward directions run as one 2B launch.  One thread per output pixel per 8-channel group: the
// sample coordinates and the four weights are computed once and reused across the group; reads of neighbouring
// pixels hit neighbouring source texels (L1/L2 resident maps), writes are fully coalesced.
#include "common.cuh"

namespace irr {

constexpr int WARP_CG = 8;  // channels per thread

__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ x, long long x_bs,
                                                   const float* __restrict__ flow, long long flow_bs,
                                                   const float* __restrict__ minuend, long long m_bs,
                                                   float* __restrict__ out, long long o_bs,
                                                   float* __restrict__ mask_out, GridArgs g, int B, int C, int H, int W,
                                                   int shift) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;
  int HW = H * W;
  if (pix >= HW) return;
  int b = blockIdx.z;
  int c0 = blockIdx.y * WARP_CG;
  int yy = pix / W, xx = pix - yy * W;
  const float* fp = flow + (size_t)b * flow_bs + pix;
  float u = __ldg(fp), v = __ldg(fp + HW);
  float ix, iy;
  sample_coords(g, u, v, xx, yy, W, H, ix, iy);
  Taps t = make_taps(ix, iy, W, H);
  int bs = b + shift;
  if (bs >= B) bs -= B;
  const float* xp = x + (size_t)bs * x_bs;
  if (mask_out && c0 == 0) mask_out[(size_t)b * HW + pix] = t.mask;
  int cend = min(c0 + WARP_CG, C);
  for (int c = c0; c < cend; ++c) {
    float val = 0.f;
    if (t.mask != 0.f) val = gather_bilinear(xp + (size_t)c * HW, t, W, H);  // x_warp * mask (pwc_modules.py:133)
    if (minuend) val = __fsub_rn(__ldg(minuend + (size_t)b * m_bs + (size_t)c * HW + pix), val);
    out[(size_t)b * o_bs + (size_t)c * HW + pix] = val;
  }
}

}  // namespace irr

using namespace irr;

extern "C" int irr_warp_fwd(const float* x, long long x_bs, const float* flow, long long flow_bs, const float* lin_x,
                            const float* lin_y, const float* minuend, long long minuend_bs, float* out,
                            long long out_bs, float* mask_out, int B, int C, int H, int W, int H_im, int W_im,
                            float div_flow, int x_batch_shift, int grid_flags, irr_stream_t stream) {
  const char* fn = "irr_warp_fwd";
  IRR_REQUIRE(x && flow && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H_im > 0 && W_im > 0, fn, "non-positive size");
  IRR_REQUIRE(x_batch_shift >= 0 && x_batch_shift < B, fn, "x_batch_shift out of range");
  IRR_REQUIRE(B <= 65535, fn, "batch too large");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  dim3 grid((H * W + 255) / 256, (C + WARP_CG - 1) / WARP_CG, B);
  warp_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_bs, flow, flow_bs, minuend, minuend_bs, out, out_bs, mask_out,
                                                   g, B, C, H, W, x_batch_shift);
  return check_launch(fn);
}
