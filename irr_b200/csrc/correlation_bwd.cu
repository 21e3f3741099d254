// Backward of the cost volume — replaces correlation_backward_input1 / _input2 of
// models/correlation_package/correlation_cuda_kernel.cu:116-300 for the PWC parameters (pad 4, k 1, md 4, s 1/1):
//
//   out[b, (dy+4)*9+(dx+4), y, x] = (1/C) sum_c f1[b,c,y,x] * f2[b,c,y+dy,x+dx]                  (forward)
//   grad_f1[b,c,y,x] = (1/C) sum_d g[b,d,y,x]       * f2[b,c,y+dy,x+dx]
//   grad_f2[b,c,y,x] = (1/C) sum_d g[b,d,y-dy,x-dx] * f1[b,c,y-dy,x-dx]          (zero outside the image)
//
// Both are the forward's windowed product with the roles turned: one kernel, FLIP selects the sign of the window.
// Tile = 8 x 32 output pixels, thread = pixel.  The 81 gradient values a pixel needs — g[d] at the pixel itself
// (grad_f1) or at the pixel displaced by -d (grad_f2) — are loaded ONCE per tile into registers (coalesced along x for
// every d); the other operand's 16 x 40 halo tile is staged per 8-channel chunk in shared memory (zero-filled outside
// the image = the zero padding), so one channel costs 81 LDS + 81 FFMA per thread and nothing is re-read from HBM
// except the 2.5x halo (L2).  The reference's backward re-reads the padded NHWC copies 81 x per element through
// global memory and needs atomics-free but uncoalesced gathers; it also materialises two padded copies first.
// Training is outside the inference north-star: this kernel is correct and coalesced, not tuned (no TMA ring yet).
#include "common.cuh"

namespace irr {

constexpr int BT_H = 8, BT_W = 32, B_MD = 4, B_ND = 9, B_CC = 8;
constexpr int BH_H = BT_H + 2 * B_MD, BH_W = BT_W + 2 * B_MD;  // 16 x 40 halo
constexpr int BH_P = 41;                                       // row pitch (odd: conflict-free for any window shift)

template <bool FLIP>
__global__ void __launch_bounds__(BT_H * BT_W) corr_bwd_kernel(const float* __restrict__ g, long long g_bs,
                                                               const float* __restrict__ other, long long o_bs,
                                                               float* __restrict__ grad, long long gr_bs, int C, int H,
                                                               int W, int tiles_x, int tiles_y) {
  __shared__ float tile[B_CC][BH_H][BH_P];
  const int tid = threadIdx.x;
  const int r = tid >> 5, xl = tid & 31;
  const int tile_id = blockIdx.x;
  const int tx = tile_id % tiles_x, ty = (tile_id / tiles_x) % tiles_y, b = tile_id / (tiles_x * tiles_y);
  const int y0 = ty * BT_H, x0 = tx * BT_W;
  const int y = y0 + r, x = x0 + xl;
  const size_t HW = (size_t)H * W;
  const bool inside = y < H && x < W;

  // the 81 gradient values of this pixel
  float gr[B_ND * B_ND];
  const float* gb = g + (size_t)b * g_bs;
#pragma unroll
  for (int d = 0; d < B_ND * B_ND; ++d) {
    const int dy = d / B_ND - B_MD, dx = d % B_ND - B_MD;
    const int sy = FLIP ? y - dy : y, sx = FLIP ? x - dx : x;
    const bool ok = inside && sy >= 0 && sy < H && sx >= 0 && sx < W;
    gr[d] = ok ? __ldg(gb + (size_t)d * HW + (size_t)sy * W + sx) : 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  const float* ob = other + (size_t)b * o_bs;
  float* out = grad + (size_t)b * gr_bs;
  for (int c0 = 0; c0 < C; c0 += B_CC) {
    __syncthreads();  // previous chunk fully consumed
    for (int i = tid; i < B_CC * BH_H * BH_W; i += BT_H * BT_W) {
      const int cc = i / (BH_H * BH_W), rem = i - cc * (BH_H * BH_W);
      const int hr = rem / BH_W, hx = rem - hr * BH_W;
      const int gy = y0 - B_MD + hr, gx = x0 - B_MD + hx, c = c0 + cc;
      const bool ok = c < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
      tile[cc][hr][hx] = ok ? __ldg(ob + (size_t)c * HW + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int cc = 0; cc < B_CC; ++cc) {
      float acc = 0.f;
#pragma unroll
      for (int dyi = 0; dyi < B_ND; ++dyi) {
        const float* row = &tile[cc][FLIP ? r + 2 * B_MD - dyi : r + dyi][xl];
#pragma unroll
        for (int dxi = 0; dxi < B_ND; ++dxi) acc = fmaf(gr[dyi * B_ND + dxi], row[FLIP ? 2 * B_MD - dxi : dxi], acc);
      }
      if (inside && c0 + cc < C) out[(size_t)(c0 + cc) * HW + (size_t)y * W + x] = acc * inv_c;
    }
  }
}

}  // namespace irr

using namespace irr;

extern "C" int irr_correlation_bwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs,
                                   const float* grad_out, long long go_bs, float* grad_f1, long long g1_bs,
                                   float* grad_f2, long long g2_bs, int B, int C, int H, int W, int max_disp,
                                   irr_stream_t stream) {
  const char* fn = "irr_correlation_bwd";
  IRR_REQUIRE(f1 && f2 && grad_out, fn, "null pointer");
  IRR_REQUIRE(grad_f1 || grad_f2, fn, "no gradient requested");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  IRR_REQUIRE(max_disp == B_MD, fn, "only max_disp == 4 is compiled");
  const int tiles_x = (W + BT_W - 1) / BT_W, tiles_y = (H + BT_H - 1) / BT_H;
  const long long nt = (long long)tiles_x * tiles_y * B;
  IRR_REQUIRE(nt <= 0x7fffffffLL, fn, "too many tiles");
  cudaStream_t st = as_stream(stream);
  if (grad_f1)
    corr_bwd_kernel<false><<<(unsigned)nt, BT_H * BT_W, 0, st>>>(grad_out, go_bs, f2, f2_bs, grad_f1, g1_bs, C, H, W, tiles_x,
                                                                 tiles_y);
  if (grad_f2)
    corr_bwd_kernel<true><<<(unsigned)nt, BT_H * BT_W, 0, st>>>(grad_out, go_bs, f1, f1_bs, grad_f2, g2_bs, C, H, W, tiles_x,
                                                                tiles_y);
  return check_launch(fn);
}
