// Standalone WarpingLayer (models/pwc_modules.py:115-133): bilinear gather + rounding-exact hard mask, optionally
// fused with the `a - warp(b)` the refinement nets consume (IRR_PWC.py:132-133,144-145) and a batch rotation so the
// forward and backward directions run as one 2B launch.  One thread per output pixel per 8-channel group: the
// sample coordinates and the four weights are computed once and reused across the group; reads of neighbouring
// pixels hit neighbouring source texels (L1/L2 resident maps), writes are fully coalesced.
#include "common.cuh"

namespace irr {

// WARP_CG channels per thread: 8 for feature maps; 4 for the 2- / 3-channel flow and image warps, whose throughput comes
// from occupancy rather than from loads in flight per thread (measured: 8-wide at C = 3 was 30 % slower than the old loop).
template <int WARP_CG>
__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ x, long long x_bs,
                                                   const float* __restrict__ flow, long long flow_bs,
                                                   const float* __restrict__ minuend, long long m_bs,
                                                   float* __restrict__ out, long long o_bs,
                                                   float* __restrict__ mask_out, GridArgs g, int B, int C, int H, int W,
                                                   int P, int shift) {
  int pix = blockIdx.x * blockDim.x + threadIdx.x;   // dense pixel index; rows are stored with pitch P >= W
  if (pix >= H * W) return;
  int b = blockIdx.z;
  int c0 = blockIdx.y * WARP_CG;
  int yy = pix / W, xx = pix - yy * W;
  const int HW = H * P;            // channel stride
  const int pp = yy * P + xx;      // pitched pixel offset
  const float* fp = flow + (size_t)b * flow_bs + pp;
  float u = __ldg(fp), v = __ldg(fp + HW);
  float ix, iy;
  sample_coords(g, u, v, xx, yy, W, H, ix, iy);
  Taps t = make_taps(ix, iy, W, H);
  int bs = b + shift;
  if (bs >= B) bs -= B;
  const float* xp = x + (size_t)bs * x_bs;
  if (mask_out && c0 == 0) mask_out[(size_t)b * H * W + pix] = t.mask;   // the mask output is dense
  // All tap loads of the thread's channel group first (up to 32 + 8 independent loads in flight), then the arithmetic in
  // grid_sampler's tap order, then the stores: a sequential channel loop left one channel's four loads outstanding per
  // thread and the kernel at ~1.4 TB/s on the 436 x 1024 image warps.
  const int nc = min(WARP_CG, C - c0);
  const bool live = t.mask != 0.f;
  const int x0 = min(max(t.x0, 0), W - 1), x1 = min(max(t.x0 + 1, 0), W - 1);
  const int y0 = min(max(t.y0, 0), H - 1), y1 = min(max(t.y0 + 1, 0), H - 1);
  const size_t o00 = (size_t)y0 * P + x0, o01 = (size_t)y0 * P + x1, o10 = (size_t)y1 * P + x0, o11 = (size_t)y1 * P + x1;
  const float* xc = xp + (size_t)c0 * HW;
  float tp[WARP_CG][4], mn[WARP_CG];
#pragma unroll
  for (int j = 0; j < WARP_CG; ++j) {
    const bool ok = live && j < nc;
    const float* pc = xc + (size_t)j * HW;
    tp[j][0] = ok ? __ldg(pc + o00) : 0.f;
    tp[j][1] = ok ? __ldg(pc + o01) : 0.f;
    tp[j][2] = ok ? __ldg(pc + o10) : 0.f;
    tp[j][3] = ok ? __ldg(pc + o11) : 0.f;
    mn[j] = (minuend && j < nc) ? __ldg(minuend + (size_t)b * m_bs + (size_t)(c0 + j) * HW + pp) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < WARP_CG; ++j) {
    if (j < nc) {
      float val = 0.f;
      if (live) {   // same tap order and roundings as gather_bilinear / grid_sampler_2d_kernel; x_warp * mask (pwc_modules.py:133)
        val = __fmul_rn(tp[j][0], t.w00);
        val = fmaf(tp[j][1], t.w01, val);
        val = fmaf(tp[j][2], t.w10, val);
        val = fmaf(tp[j][3], t.w11, val);
      }
      if (minuend) val = __fsub_rn(mn[j], val);
      out[(size_t)b * o_bs + (size_t)(c0 + j) * HW + pp] = val;
    }
  }
}

// Backward of the WarpingLayer: out = mask * grid_sample(x, grid(flow)) with the mask treated as a
// constant, exactly as autograd treats `(mask >= 1.0).float()` in models/pwc_modules.py:129-133.  One thread per output
// pixel; the channel loop is strided over gridDim.y so wide feature maps spread over more threads.
//   grad_x[b, c, tap]  += mask * w_tap * grad_out[b, c, y, x]                      (atomicAdd: taps of different pixels overlap)
//   grad_flow[b, 0/1]   = mask * sum_c grad_out * d(bilinear)/d(ix, iy) * (W-1)/2 * 2 / max(W_im-1, 1) / div_flow
// with d(bilinear)/d(ix) written as grid_sampler_2d_backward does (ATen/native/cuda/GridSampler.cu): out-of-bounds taps
// contribute zero.  grad_x must be zero-filled by the caller; grad_flow is fully written (partial sums of the channel
// groups are combined with atomicAdd when gridDim.y > 1, so it is zero-filled by the caller as well).
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ x, long long x_bs,
                                                       const float* __restrict__ flow, long long flow_bs,
                                                       const float* __restrict__ go, long long go_bs,
                                                       float* __restrict__ gx, long long gx_bs,
                                                       float* __restrict__ gf, long long gf_bs, GridArgs g, int C, int H,
                                                       int W, float sx, float sy) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int b = blockIdx.z;
  const int yy = pix / W, xx = pix - yy * W;
  const float* fp = flow + (size_t)b * flow_bs + pix;
  float ix, iy;
  sample_coords(g, __ldg(fp), __ldg(fp + HW), xx, yy, W, H, ix, iy);
  const Taps t = make_taps(ix, iy, W, H);
  if (t.mask == 0.f) return;   // masked pixel: no gradient reaches x or flow (grad buffers are zero-filled)
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const float bx = ix - fx0, ax = 1.0f - bx, by = iy - fy0, ay = 1.0f - by;   // distances to the far / near corners
  const bool x0ok = t.x0 >= 0 && t.x0 < W, x1ok = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool y0ok = t.y0 >= 0 && t.y0 < H, y1ok = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  const long long o_nw = (long long)t.y0 * W + t.x0;
  float gix = 0.f, giy = 0.f;
  for (int c = blockIdx.y; c < C; c += gridDim.y) {
    const float gO = __ldg(go + (size_t)b * go_bs + (size_t)c * HW + pix);
    const float* xp = x + (size_t)b * x_bs + (size_t)c * HW;
    float* gp = gx ? gx + (size_t)b * gx_bs + (size_t)c * HW : nullptr;
    const float nw = (x0ok && y0ok) ? __ldg(xp + o_nw) : 0.f;
    const float ne = (x1ok && y0ok) ? __ldg(xp + o_nw + 1) : 0.f;
    const float sw = (x0ok && y1ok) ? __ldg(xp + o_nw + W) : 0.f;
    const float se = (x1ok && y1ok) ? __ldg(xp + o_nw + W + 1) : 0.f;
    if (gp) {
      if (x0ok && y0ok) atomicAdd(gp + o_nw, t.w00 * gO);
      if (x1ok && y0ok) atomicAdd(gp + o_nw + 1, t.w01 * gO);
      if (x0ok && y1ok) atomicAdd(gp + o_nw + W, t.w10 * gO);
      if (x1ok && y1ok) atomicAdd(gp + o_nw + W + 1, t.w11 * gO);
    }
    gix += gO * (ay * (ne - nw) + by * (se - sw));
    giy += gO * (ax * (sw - nw) + bx * (se - ne));
  }
  if (gf) {
    float* q = gf + (size_t)b * gf_bs + pix;
    if (gridDim.y == 1) {
      q[0] = gix * sx;
      q[HW] = giy * sy;
    } else {
      atomicAdd(q, gix * sx);
      atomicAdd(q + HW, giy * sy);
    }
  }
}

}  // namespace irr

using namespace irr;

extern "C" int irr_warp_bwd(const float* x, long long x_bs, const float* flow, long long flow_bs, const float* lin_x,
                            const float* lin_y, const float* grad_out, long long go_bs, float* grad_x, long long gx_bs,
                            float* grad_flow, long long gf_bs, int B, int C, int H, int W, int H_im, int W_im,
                            float div_flow, int grid_flags, irr_stream_t stream) {
  const char* fn = "irr_warp_bwd";
  IRR_REQUIRE(x && flow && grad_out && (grad_x || grad_flow), fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H_im > 0 && W_im > 0, fn, "non-positive size");
  IRR_REQUIRE(B <= 65535, fn, "batch too large");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  // d(ix)/d(u) = (W-1)/2 * 2 / max(W_im-1, 1) / div_flow   (pwc_modules.py:121-122 + grid_sampler's unnormalise)
  const float sx = (float)((double)(W - 1) / (double)g.den_x / (double)div_flow);
  const float sy = (float)((double)(H - 1) / (double)g.den_y / (double)div_flow);
  int cy = C < 8 ? 1 : (C + 7) / 8;   // channel groups per pixel: 8 channels per thread
  if (cy > 64) cy = 64;
  dim3 grid((H * W + 255) / 256, cy, B);
  warp_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_bs, flow, flow_bs, grad_out, go_bs, grad_x, gx_bs, grad_flow,
                                                       gf_bs, g, C, H, W, sx, sy);
  return check_launch(fn);
}

extern "C" int irr_warp_fwd(const float* x, long long x_bs, const float* flow, long long flow_bs, const float* lin_x,
                            const float* lin_y, const float* minuend, long long minuend_bs, float* out,
                            long long out_bs, float* mask_out, int B, int C, int H, int W, int H_im, int W_im,
                            float div_flow, int x_batch_shift, int grid_flags, int pitch, irr_stream_t stream) {
  const char* fn = "irr_warp_fwd";
  const int P = pitch > 0 ? pitch : W;
  IRR_REQUIRE(P >= W, fn, "row pitch smaller than the width");
  IRR_REQUIRE(x && flow && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H_im > 0 && W_im > 0, fn, "non-positive size");
  IRR_REQUIRE(x_batch_shift >= 0 && x_batch_shift < B, fn, "x_batch_shift out of range");
  IRR_REQUIRE(B <= 65535, fn, "batch too large");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  const int cg = C <= 4 ? 4 : 8;
  dim3 grid((H * W + 255) / 256, (C + cg - 1) / cg, B);
  auto kern = C <= 4 ? warp_kernel<4> : warp_kernel<8>;
  kern<<<grid, 256, 0, as_stream(stream)>>>(x, x_bs, flow, flow_bs, minuend, minuend_bs, out, out_bs, mask_out,
                                                   g, B, C, H, W, P, x_batch_shift);
  return check_launch(fn);
}
