/* CPU restatement of the reference's (orphaned) CUDA correlation extension.
 *
 * TEST INFRASTRUCTURE ONLY — compiled to oracle/_build/libcorr_ref.so by oracle/Makefile and loaded by
 * tests/ and bench.py's cpu_baseline / --impl reference legs; never linked into the product library.
 *
 * Restates, in plain C, what the three launches of
 *   models/correlation_package/correlation_cuda_kernel.cu:352-381 compute:
 *   - shape math ......... correlation_cuda.cc:17-36
 *   - pad + NCHW->NHWC .... correlation_cuda_kernel.cu:15-39  (folded into the bounds test below)
 *   - forward ............. correlation_cuda_kernel.cu:41-114
 * for generic (pad_size, kernel_size, max_displacement, stride1, stride2).  corr_type_multiply is accepted
 * and ignored exactly like the reference (correlation_cuda.cc:14; never read by the kernels).
 * The reduction order is a straight sum over (j, i, c); the CUDA kernel sums 32 lane-partials, so equality
 * with it is to rounding (≈1e-7), not bitwise.  Parity is pinned against models/pwc_modules.py:42-62
 * (compute_cost_volume) by oracle/gen_golden.py + tests/test_oracle.py.
 */
#include <math.h>
#include <stddef.h>

/* Output spatial size — correlation_cuda.cc:23-32. */
int corr_ref_out_shape(int H, int W, int pad, int ksize, int max_disp, int s1, int s2, int* oc, int* oh, int* ow) {
  if (ksize < 1 || (ksize & 1) == 0 || s1 < 1 || s2 < 1 || max_disp < 0 || pad < 0) return -1;
  int krad = (ksize - 1) / 2;
  int border = krad + max_disp;
  int pH = H + 2 * pad, pW = W + 2 * pad;
  int d = (max_disp / s2) * 2 + 1;
  *oc = d * d;
  *oh = (int)ceilf((float)(pH - 2 * border) / (float)s1);
  *ow = (int)ceilf((float)(pW - 2 * border) / (float)s1);
  return (*oh > 0 && *ow > 0) ? 0 : -1;
}

static inline float padded_at(const float* x, int C, int H, int W, int pad, int c, int py, int px) {
  /* value of the zero-padded NHWC copy rInput[n][py][px][c] (correlation_cuda_kernel.cu:35-38) */
  int y = py - pad, xx = px - pad;
  if (y < 0 || y >= H || xx < 0 || xx >= W) return 0.0f;
  return x[((size_t)c * H + y) * W + xx];
}

/* in1, in2: B x C x H x W contiguous; out: B x oc x oh x ow contiguous. Returns 0 on success. */
int corr_ref_forward(const float* in1, const float* in2, float* out, int B, int C, int H, int W, int pad, int ksize,
                     int max_disp, int s1, int s2) {
  int oc, oh, ow;
  if (corr_ref_out_shape(H, W, pad, ksize, max_disp, s1, s2, &oc, &oh, &ow)) return -1;
  int krad = (ksize - 1) / 2;
  int drad = max_disp / s2;
  int dsize = 2 * drad + 1;
  float nelems = (float)(ksize * ksize * C); /* .cu:73 */
  for (int n = 0; n < B; ++n) {
    const float* a = in1 + (size_t)n * C * H * W;
    const float* b = in2 + (size_t)n * C * H * W;
    for (int by = 0; by < oh; ++by)
      for (int bx = 0; bx < ow; ++bx) {
        int y1 = by * s1 + max_disp, x1 = bx * s1 + max_disp; /* .cu:61-62, padded coordinates */
        for (int tj = -drad; tj <= drad; ++tj)
          for (int ti = -drad; ti <= drad; ++ti) {
            int x2 = x1 + ti * s2, y2 = y1 + tj * s2; /* .cu:85-86 */
            float acc = 0.0f;
            for (int j = -krad; j <= krad; ++j)
              for (int i = -krad; i <= krad; ++i)
                for (int c = 0; c < C; ++c)
                  acc += padded_at(a, C, H, W, pad, c, y1 + j, x1 + i) * padded_at(b, C, H, W, pad, c, y2 + j, x2 + i);
            int tc = (tj + drad) * dsize + (ti + drad); /* .cu:105 */
            out[(((size_t)n * oc + tc) * oh + by) * ow + bx] = acc / nelems;
          }
      }
  }
  return 0;
}
