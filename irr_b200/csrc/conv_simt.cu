// conv() of models/pwc_modules.py:8-19 — fp32 CUDA-core implicit GEMM (IRR_MATH_FP32_SIMT).
//
//   y[b, n, oy, ox] = addend + alpha * lrelu( bias[n] + sum_{c,ky,kx} x[b, c, oy*s - pad + ky*dil, ox*s - pad + kx*dil] * w[n, c, ky, kx] )
//
// GEMM view: M = B*Ho*Wo output pixels (flattened, so odd sizes like 109x256 or 47x156 waste nothing), N = Cout,
// K = Cin*k*k in the reference's OIHW order (k index = c*k*k + ky*k + kx), so the dense blocks' "read the suffix of the
// concat buffer" needs no re-packing: channel c of the slice is just row block c of the packed weight.
// Tile: 128 (M) x BN (N) x 16 (K) per 256-thread CTA, thread tile 8 x BN/16, register-prefetch double buffering,
// im2col-free A gather (coalesced along ox, zero padding by predicate), weights pre-transposed to [K_pad][N_pad] at
// pack time so B loads are 16-byte coalesced and K needs no tail predicate.  Used for every layer in fp32 mode and,
// when the tcgen05 path is enabled, for the layers that are not a real contraction (Cin = 3, Cout <= 16, ...).
#include "common.cuh"

namespace irr {

constexpr int BM = 128, BK = 16, CONV_THREADS = 256;

struct ConvArgs {
  const float* x; long long x_bs;
  const float* w;  // packed [K_pad][N_pad]
  const float* bias;
  const float* addend; long long a_bs;
  float* y; long long y_bs;
  int B, Cin, H, W, Cout, Ho, Wo, stride, dil, pad;
  int K, N_pad;
  long long M;
  float slope, alpha;
};

__host__ __device__ static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <int KS, int BN>
__global__ void __launch_bounds__(CONV_THREADS) conv_simt_kernel(ConvArgs p) {
  constexpr int TN = BN / 16;
  constexpr int T = KS * KS;
  constexpr int B_F4 = BK * BN / 4;                                    // float4 per B tile
  constexpr int B_LD = (B_F4 + CONV_THREADS - 1) / CONV_THREADS;       // float4 per thread
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  const int tm = tid & 15, tn = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int HWo = p.Ho * p.Wo;
  const size_t HW = (size_t)p.H * p.W;

  // A-gather role: this thread always loads pixel (m0 + tid%128) for k rows (tid/128) + 2*i
  const int am = tid & (BM - 1);
  const int ak0 = tid >> 7;
  const long long mg = m0 + am;
  const bool m_ok = mg < p.M;
  int ab = 0, aoy = 0, aox = 0;
  if (m_ok) {
    ab = (int)(mg / HWo);
    int rem = (int)(mg - (long long)ab * HWo);
    aoy = rem / p.Wo;
    aox = rem - aoy * p.Wo;
  }
  const int iy0 = aoy * p.stride - p.pad, ix0 = aox * p.stride - p.pad;
  const float* xb = p.x + (size_t)ab * p.x_bs;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[8];
  float4 rb[B_LD];

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int k = k0 + ak0 + 2 * i;
      float v = 0.f;
      if (m_ok && k < p.K) {
        int c, iy, ix;
        if (KS == 1) {
          c = k; iy = iy0; ix = ix0;
        } else {
          c = k / T;
          int t = k - c * T;
          int ky = t / KS, kx = t - ky * KS;
          iy = iy0 + ky * p.dil; ix = ix0 + kx * p.dil;
        }
        if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) v = __ldg(xb + (size_t)c * HW + (size_t)iy * p.W + ix);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int f = tid + i * CONV_THREADS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < B_F4) {
        int kk = f / (BN / 4), nq = f - kk * (BN / 4);
        int n = n0 + nq * 4;
        if (n < p.N_pad) v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)(k0 + kk) * p.N_pad + n));
      }
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][ak0 + 2 * i][am] = ra[i];
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int f = tid + i * CONV_THREADS;
      if (f < B_F4) {
        int kk = f / (BN / 4), nq = f - kk * (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[i];
      }
    }
  };

  const int K_pad = round_up(p.K, BK);  // same formula as the packer
  const int ntiles = K_pad / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_tile((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], bv[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tm * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tm * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[buf][kk][tn * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + (i < 4 ? tm * 4 + i : 64 + tm * 4 + (i - 4));
    if (m >= p.M) continue;
    int b = (int)(m / HWo);
    int pix = (int)(m - (long long)b * HWo);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tn * TN + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j] + __ldg(p.bias + n);
      v = leaky(v, p.slope);
      v *= p.alpha;
      if (p.addend) v += __ldg(p.addend + (size_t)b * p.a_bs + (size_t)n * HWo + pix);
      p.y[(size_t)b * p.y_bs + (size_t)n * HWo + pix] = v;
    }
  }
}

// OIHW -> [K_pad][N_pad] (k = c*T + tap), zero padded.
__global__ void pack_simt_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int K, int K_pad,
                                 int N_pad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K_pad * N_pad) return;
  int n = (int)(i % N_pad), k = (int)(i / N_pad);
  out[i] = (n < Cout && k < K) ? __ldg(w + (size_t)n * K + k) : 0.f;
}

// ------------------------------------------------------------------------------------------------ direct kernel
// Layers that are not a contraction worth a tile pipeline — K = Cin*k*k <= 32 and Cout <= 16: the first pyramid layer
// 3 -> 16 (k 3, stride 2, pwc_modules.py:95-104) and the 16 -> 3 1x1 convs of the occlusion up-sampler's inputs
// (IRR_PWC.py:44-45) — are pure HBM streams: one thread per output pixel, all K taps loaded first (K independent loads in
// flight), weights broadcast from shared memory, every output channel of the pixel in registers, plane-coalesced stores.
// fp32 FMAs in the reference's k order (c, ky, kx).  On the tensor-core kernel's gather variant these layers ran 4-10x
// above their HBM time (3 -> 16 at 436 x 1024: 356 us against 31 us of bytes).  Row pitches are honoured (ABI 2).
struct DirectArgs {
  ConvArgs c;
  int Pi, Po;
};
template <int K, int KS, int NP>
__global__ void __launch_bounds__(256) conv_direct_kernel(DirectArgs q) {
  constexpr int T = KS * KS;
  const ConvArgs& p = q.c;
  __shared__ __align__(16) float ws[K * NP];
  __shared__ float bs[NP];
  for (int i = threadIdx.x; i < K * NP; i += 256) ws[i] = __ldg(p.w + (i / NP) * p.N_pad + (i % NP));   // packed [K_pad][N_pad]
  if (threadIdx.x < NP) bs[threadIdx.x] = threadIdx.x < p.Cout ? __ldg(p.bias + threadIdx.x) : 0.f;
  __syncthreads();
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= p.Ho * p.Wo) return;
  const int b = blockIdx.y;
  const int oy = pix / p.Wo, ox = pix - oy * p.Wo;
  const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
  const float* xb = p.x + (size_t)b * p.x_bs;
  const size_t HWi = (size_t)p.H * q.Pi;
  float v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int c = k / T, t = k - c * T, ky = t / KS, kx = t - ky * KS;
    const int iy = iy0 + ky * p.dil, ix = ix0 + kx * p.dil;
    const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
    v[k] = ok ? __ldg(xb + (size_t)c * HWi + (size_t)iy * q.Pi + ix) : 0.f;
  }
  float acc[NP];
#pragma unroll
  for (int n = 0; n < NP; ++n) acc[n] = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int n4 = 0; n4 < NP / 4; ++n4) {
      const float4 w = *reinterpret_cast<const float4*>(&ws[k * NP + n4 * 4]);
      acc[n4 * 4] = fmaf(v[k], w.x, acc[n4 * 4]);
      acc[n4 * 4 + 1] = fmaf(v[k], w.y, acc[n4 * 4 + 1]);
      acc[n4 * 4 + 2] = fmaf(v[k], w.z, acc[n4 * 4 + 2]);
      acc[n4 * 4 + 3] = fmaf(v[k], w.w, acc[n4 * 4 + 3]);
    }
  }
  const size_t HWo = (size_t)p.Ho * q.Po;
  const size_t o = (size_t)oy * q.Po + ox;
#pragma unroll
  for (int n = 0; n < NP; ++n) {
    if (n < p.Cout) {
      float r = leaky(acc[n] + bs[n], p.slope) * p.alpha;
      if (p.addend) r += __ldg(p.addend + (size_t)b * p.a_bs + (size_t)n * HWo + o);
      p.y[(size_t)b * p.y_bs + (size_t)n * HWo + o] = r;
    }
  }
}

bool simt_direct_supported(int Cout, int Cin, int ks) {
  return Cout <= 16 && ((ks == 3 && Cin == 3) || (ks == 1 && Cin == 16));
}

template <int KS>
static void launch_simt(const ConvArgs& a, cudaStream_t st) {
  unsigned gm = (unsigned)((a.M + BM - 1) / BM);
  if (a.Cout > 64) {
    dim3 g(gm, (a.Cout + 127) / 128);
    conv_simt_kernel<KS, 128><<<g, CONV_THREADS, 0, st>>>(a);
  } else if (a.Cout > 32) {
    dim3 g(gm, 1);
    conv_simt_kernel<KS, 64><<<g, CONV_THREADS, 0, st>>>(a);
  } else if (a.Cout > 16) {
    dim3 g(gm, 1);
    conv_simt_kernel<KS, 32><<<g, CONV_THREADS, 0, st>>>(a);
  } else {
    dim3 g(gm, 1);
    conv_simt_kernel<KS, 16><<<g, CONV_THREADS, 0, st>>>(a);
  }
}

size_t simt_packed_bytes(int Cout, int Cin, int ks) {
  return (size_t)round_up(Cin * ks * ks, BK) * round_up(Cout, 16) * sizeof(float);
}

int simt_pack(const float* w, void* out, int Cout, int Cin, int ks, cudaStream_t st) {
  int K = Cin * ks * ks, K_pad = round_up(K, BK), N_pad = round_up(Cout, 16);
  long long total = (long long)K_pad * N_pad;
  pack_simt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, (float*)out, Cout, K, K_pad, N_pad);
  return check_launch("irr_conv2d_pack_weights");
}

int simt_conv(const float* x, long long x_bs, const void* w, const float* bias, const float* addend, long long a_bs,
              float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ks, int stride, int dil,
              float slope, float alpha, cudaStream_t st, int pitch_in, int pitch_out) {
  ConvArgs a;
  a.x = x; a.x_bs = x_bs; a.w = (const float*)w; a.bias = bias; a.addend = addend; a.a_bs = a_bs; a.y = y; a.y_bs = y_bs;
  a.B = B; a.Cin = Cin; a.H = H; a.W = W; a.Cout = Cout; a.stride = stride; a.dil = dil;
  a.pad = ((ks - 1) * dil) / 2;
  a.Ho = (H + 2 * a.pad - dil * (ks - 1) - 1) / stride + 1;
  a.Wo = (W + 2 * a.pad - dil * (ks - 1) - 1) / stride + 1;
  a.K = Cin * ks * ks;
  a.N_pad = round_up(Cout, 16);
  a.M = (long long)B * a.Ho * a.Wo;
  a.slope = slope; a.alpha = alpha;
  const int Pi = pitch_in > 0 ? pitch_in : W, Po = pitch_out > 0 ? pitch_out : a.Wo;
  if (Pi < W || Po < a.Wo) return fail_arg("irr_conv2d_fwd", "row pitch smaller than the width");
  if (simt_direct_supported(Cout, Cin, ks) && B <= 65535) {
    DirectArgs d;
    d.c = a; d.Pi = Pi; d.Po = Po;
    dim3 g((unsigned)((a.Ho * a.Wo + 255) / 256), (unsigned)B);
    if (ks == 3) conv_direct_kernel<27, 3, 16><<<g, 256, 0, st>>>(d);
    else if (Cout <= 4) conv_direct_kernel<16, 1, 4><<<g, 256, 0, st>>>(d);
    else conv_direct_kernel<16, 1, 16><<<g, 256, 0, st>>>(d);
    return check_launch("irr_conv2d_fwd");
  }
  if (Pi != W || Po != a.Wo)
    return fail_arg("irr_conv2d_fwd", "row pitches are implemented by the IRR_MATH_TC_3XF16 path and the direct thin-layer kernel only");
  if (ks == 1) launch_simt<1>(a, st); else launch_simt<3>(a, st);
  return check_launch("irr_conv2d_fwd");
}

}  // namespace irr
