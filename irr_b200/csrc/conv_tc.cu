// conv() of models/pwc_modules.py:8-19 as an im2col-free implicit GEMM on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//   D[128 px x N] (fp32, TMEM) += A[128 px x K] * B[N x K]^T,   K = (channel chunk, tap) blocks of <= 32 channels.
//
// Why this shape:
//   * fp32 parity (<= 1e-4, north_star) rules out a plain TF32/BF16 contraction, so the default mode is 3xTF32: every
//     fp32 operand is split  v = hi + lo  (hi = rna_tf32(v), lo = rna_tf32(v - hi)) and the product is accumulated as
//     hi*hi + lo*hi + hi*lo in the fp32 TMEM accumulator — 2^-21-grade products at 1/3 of the TF32 rate.
//   * A (activations) never touches shared memory: the 8 producer warps gather the im2col row of "their" output pixel
//     straight from NCHW global memory (lanes = consecutive pixels -> 128-byte coalesced; zero padding, stride and
//     dilation are just address arithmetic + a predicate), split it in registers and tcgen05.st it into TMEM, where
//     tcgen05.mma reads it as the A operand (".kind::tf32 [d], [a_tmem], b_desc").  SS-mode TF32 would saturate the
//     128 B/clk shared-memory port with A+B operand reads; with A in TMEM only B streams through smem.
//   * B (weights) is pre-packed on the device, once per layer, into the exact shared-memory image the MMA wants
//     (K-major, 128-byte swizzle, hi and lo planes, one image per (N tile, K block)), so staging is one cp.async.bulk
//     per K block completing on an mbarrier — no tensor map, no per-element work.
//   * K order is (channel chunk outer, tap inner): the nine taps of a chunk re-read the same activation lines, which
//     stay in L1; L2 sees each activation ~once and the weight stream once per CTA.
//   * persistent CTAs, warp-specialised (roles and the measurements behind them are listed above the kernel and in
//     DESIGN.md §4.2): 8 producer warps in two alternating groups, two MMA-issuing warps (one per 128-row half), a
//     weight loader, four epilogue warps; hand-off is mbarrier-only (a_full / b_full / a_empty / b_empty via
//     tcgen05.commit / acc_full / acc_empty).
#include "common.cuh"

namespace irr {

constexpr int TC_BM = 128;      // pixels per CTA (UMMA M)
constexpr int TC_CK = 32;       // channels per K block (4 MMAs of K=8)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_ACC_COL = 0;     // accumulator columns [0, N)
constexpr int TC_A_COL = 256;     // A stages: columns [256, 256 + 4*64)

struct TcArgs {
  const float* x; long long x_bs;
  const uint8_t* wp;  // packed weight images
  const float* bias;
  const float* addend; long long a_bs;
  float* y; long long y_bs;
  int B, Cin, H, W, Cout, Ho, Wo, stride, dil, pad, ks;
  int n_tile, n_tiles, cchunks, nkb, passes, resident;
  unsigned b_smem_bytes;
  long long M;
  float slope, alpha;
};

__host__ __device__ static inline int tc_round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- tile geometry shared by packer, launcher and kernel
struct TcGeom {
  int n_tiles, n_tile, cchunks, taps, nkb;
  size_t img_bytes;  // one (n tile, k block) image: passes planes of n_tile x 128 B
};
static inline TcGeom tc_geom(int Cout, int Cin, int ks, int passes) {
  TcGeom g;
  g.n_tiles = (Cout + 127) / 128;
  g.n_tile = tc_round_up((Cout + g.n_tiles - 1) / g.n_tiles, 16);
  g.cchunks = (Cin + TC_CK - 1) / TC_CK;
  g.taps = ks * ks;
  g.nkb = g.cchunks * g.taps;
  g.img_bytes = (size_t)(passes == 3 ? 2 : 1) * g.n_tile * 128;
  return g;
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One lane of a converged warp (elect.sync): the tcgen05 issue path stays warp-uniform, so its operands live in
// uniform registers and ptxas does not wrap every UTCHMMA in a per-lane uniformisation loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 -> tf32, round-to-nearest (ties away): add half an ulp of the 10-bit mantissa to the magnitude, drop 13 bits.
// Two integer ops; ptxas expands cvt.rna.tf32.f32 into ~4 (it also handles Inf/NaN, which activations never are).
__device__ __forceinline__ uint32_t to_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start
// address >> 4 in bits [0,14), LBO (unused for swizzled K-major, set to 1) in [16,30), SBO = 1024 B >> 4 in [32,46),
// version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ------------------------------------------------------------------------------------------------ kernel
// Persistent: grid = #SMs; each CTA walks (n tile, 256-pixel M tile) work items.  One work item = two 128-row halves
// that share every weight stage (halves the weight stream per pixel) and own one TMEM accumulator each.
//   warps 0-3 : A producers, half 0      warps 4-7 : A producers, half 1
//   warps 8,9 : MMA issuers, one per half (measured: one issuer spends ~700 clk per K block in mbarrier waits / fences
//               / elect during which the tensor pipe drains; two issuers on alternate halves keep it fed)
//   warp 10   : weight loader            warps 12-15: epilogue (TMEM lane quadrant = warp % 4)
// TMEM columns: [0,256) accumulators (2 halves x N, double-buffered across work items when N <= 64),
//               [256,512) A ring: stage = 2 halves x (hi 32 | lo 32) columns -> 2 stages (3xTF32) / 4 stages (TF32).
// Weights: resident in shared memory for the whole kernel when the layer's images fit (K <= ~9 blocks at N=128, every
// N=32 layer), otherwise a 4-stage cp.async.bulk ring.
constexpr int TC_BM2 = 256;
constexpr int TC_THREADS2 = 512;
constexpr int TC_SB = 4;                   // weight ring depth (streaming mode)
constexpr size_t TC_RESIDENT_MAX = 168 * 1024;

__device__ float tc_zero_page[32];  // statically zero: source of out-of-image taps

struct TcSeq {  // flat (work item, k block) iterator shared by all roles
  int item, kb, nkb, items, stride;
  __device__ __forceinline__ bool valid() const { return item < items; }
  __device__ __forceinline__ void next() {
    if (++kb == nkb) { kb = 0; item += stride; }
  }
};

template <int KS, int PASSES>
__global__ void __launch_bounds__(TC_THREADS2, 1) conv_tc_kernel(TcArgs p) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  constexpr int T = KS * KS;
  constexpr int A_COLS = PASSES == 3 ? 64 : 32;       // per half per stage
  constexpr int SA = 256 / (2 * A_COLS);              // 2 or 4
  const int N = p.n_tile;
  const int NBUF = N <= 64 ? 2 : 1;
  const int acc_stride = N <= 64 ? 64 : 128;
  const uint32_t plane_bytes = (uint32_t)N * 128;
  const uint32_t img_bytes = plane_bytes * (PASSES == 3 ? 2 : 1);
  const int nkb = p.nkb;
  const bool resident = p.resident != 0;
  const int m_tiles = (int)((p.M + TC_BM2 - 1) / TC_BM2);
  const int items = m_tiles * p.n_tiles;

  uint8_t* smem_b = tc_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + p.b_smem_bytes);
  // barrier map
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s, int h) { return bar0 + 8u * (s * 2 + h); };            // [0, 8)
  auto a_empty = [&](int s) { return bar0 + 8u * (8 + s); };                      // [8, 12)
  auto b_full = [&](int s) { return bar0 + 8u * (12 + s); };                      // [12, 16)
  auto b_empty = [&](int s) { return bar0 + 8u * (16 + s); };                     // [16, 20)
  auto acc_full = [&](int b) { return bar0 + 8u * (20 + b); };                    // [20, 22)
  auto acc_empty = [&](int b) { return bar0 + 8u * (22 + b); };                   // [22, 24)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 24);
  float* bias_all = reinterpret_cast<float*>(bars + 26);  // n_tiles * N floats (<= 256)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < p.n_tiles * N; i += TC_THREADS2) bias_all[i] = i < p.Cout ? __ldg(p.bias + i) : 0.f;
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(a_full(s, 0), 4);  // 4 producer warps of the group that owns this K block
      mbar_init(a_full(s, 1), 4);
      mbar_init(a_empty(s), 2);  // one tcgen05.commit per MMA issuer (half)
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 2);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 2);
      mbar_init(acc_empty(b), 4);
    }
    mbar_fence_init();
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int HWo = p.Ho * p.Wo;

  TcSeq seq;
  seq.item = blockIdx.x; seq.kb = 0; seq.nkb = nkb; seq.items = items; seq.stride = gridDim.x;

  if (warp < 8) {
    // ======================= A producers =======================
    // Two groups of 4 warps take alternate K blocks (c % 2 == group).  A warp may touch TMEM lanes 32*(warp%4)..+31
    // only, but in BOTH halves' column ranges — so each thread owns two output pixels (row r of half 0 and row r of
    // half 1): 64 independent loads in flight per thread, and a full K-block period of slack for the
    // wait -> split -> tcgen05.st -> wait::st -> arrive chain.
    const int grp = warp >> 2, q = warp & 3;
    const size_t HW = (size_t)p.H * p.W;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int cur_item = -1;
    bool m_ok[2] = {false, false};
    int iy0[2] = {0, 0}, ix0[2] = {0, 0};
    const float* xb[2] = {p.x, p.x};
    int c = 0;
    for (; seq.valid(); seq.next(), ++c) {
      if ((c & 1) != grp) continue;
      if (seq.item != cur_item) {  // new work item: decode this thread's two output pixels
        cur_item = seq.item;
        const int mt = seq.item % m_tiles;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long mg = (long long)mt * TC_BM2 + h * 128 + q * 32 + lane;
          m_ok[h] = mg < p.M;
          int ab = 0, aoy = 0, aox = 0;
          if (m_ok[h]) {
            ab = (int)(mg / HWo);
            int rem = (int)(mg - (long long)ab * HWo);
            aoy = rem / p.Wo;
            aox = rem - aoy * p.Wo;
          }
          iy0[h] = aoy * p.stride - p.pad; ix0[h] = aox * p.stride - p.pad;
          xb[h] = p.x + (size_t)ab * p.x_bs;
        }
      }
      const int cc = seq.kb / T, tap = seq.kb - cc * T;
      const int ky = tap / KS, kx = tap - ky * KS;
      const int c0 = cc * TC_CK;
      const int nch = p.Cin - c0;
      float v[2][TC_CK];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int iy = iy0[h] + ky * p.dil, ix = ix0[h] + kx * p.dil;
        const bool ok = m_ok[h] && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
        // Zero padding without per-element predicates: an out-of-image tap reads a zero page with channel stride 0.
        const float* src = ok ? xb[h] + (size_t)c0 * HW + ((size_t)iy * p.W + ix) : tc_zero_page;
        const unsigned cstride = ok ? (unsigned)HW : 0u;
        if (nch >= TC_CK) {
#pragma unroll
          for (int j = 0; j < TC_CK; ++j) v[h][j] = __ldg(src + (size_t)(j * cstride));
        } else {  // ragged channel tail of the layer: never touch channels beyond the slice
#pragma unroll
          for (int j = 0; j < TC_CK; ++j) v[h][j] = (j < nch) ? __ldg(src + (size_t)(j * cstride)) : 0.f;
        }
      }
      const int s = c % SA;
      mbar_wait(a_empty(s), (uint32_t)(((c / SA) & 1) ^ 1));
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t a_addr = lane_addr + (uint32_t)(TC_A_COL + (s * 2 + h) * A_COLS);
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // 16 channels at a time bounds the live register set
          uint32_t hi[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) hi[j] = to_tf32(v[h][half * 16 + j]);
          tmem_st16(a_addr + half * 16, hi);
          if (PASSES == 3) {
            uint32_t lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) lo[j] = __float_as_uint(v[h][half * 16 + j] - __uint_as_float(hi[j]));  // exact;
            // the tensor core reads its top 19 bits (truncating an already 2^-12-relative residual: sign-random 2^-23)
            tmem_st16(a_addr + 32 + half * 16, lo);
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full(s, 0));
        mbar_arrive(a_full(s, 1));
      }
    }
  } else if (warp == 8 || warp == 9) {
    // ======================= MMA issuers (one per 128-row half; warp converged, one elected lane issues) ==========
    {
      const int h = warp - 8;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int c = 0, tcount = 0;
      if (resident) mbar_wait(b_full(0), 0);
      while (seq.valid()) {
        const int buf = NBUF == 2 ? (tcount & 1) : 0;
        const int use = NBUF == 2 ? (tcount >> 1) : tcount;
        mbar_wait(acc_empty(buf), (uint32_t)((use & 1) ^ 1));
        const uint32_t d_tmem = tmem_base + (uint32_t)(TC_ACC_COL + (buf * 2 + h) * acc_stride);
        for (int kb = 0; kb < nkb; ++kb, ++c) {
          const int s = c % SA;
          const uint32_t pha = (uint32_t)((c / SA) & 1);
          const int sb = c % TC_SB;
          const int cc = kb / T;
          const int nch = min(TC_CK, p.Cin - cc * TC_CK);
          const int nk = (nch + 7) >> 3;
          uint32_t b_addr;
          if (resident) {
            b_addr = smem_u32(smem_b + (size_t)kb * img_bytes);
          } else {
            mbar_wait(b_full(sb), (uint32_t)((c / TC_SB) & 1));
            b_addr = smem_u32(smem_b + (size_t)sb * img_bytes);
          }
          const uint64_t bd_hi = make_b_desc(b_addr);
          const uint64_t bd_lo = make_b_desc(b_addr + plane_bytes);
          const uint32_t a_hi = tmem_base + (uint32_t)(TC_A_COL + (s * 2 + h) * A_COLS);
          mbar_wait(a_full(s, h), pha);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll 4
            for (int j = 0; j < nk; ++j) {
              const uint64_t koff = (uint64_t)(2 * j);  // +32 B of K per MMA inside the 128-byte swizzled row (>>4)
              if (PASSES == 3) {
                tc_mma_ts(d_tmem, a_hi + 32 + 8 * j, bd_hi + koff, idesc, (kb | j) ? 1u : 0u);  // lo * hi
                tc_mma_ts(d_tmem, a_hi + 8 * j, bd_lo + koff, idesc, 1u);                       // hi * lo
                tc_mma_ts(d_tmem, a_hi + 8 * j, bd_hi + koff, idesc, 1u);                       // hi * hi
              } else {
                tc_mma_ts(d_tmem, a_hi + 8 * j, bd_hi + koff, idesc, (kb | j) ? 1u : 0u);
              }
            }
            tc_commit(a_empty(s));
            if (!resident) tc_commit(b_empty(sb));
            if (kb == nkb - 1) tc_commit(acc_full(buf));
          }
          __syncwarp();
        }
        ++tcount;
        seq.item += seq.stride;
      }
    }
  } else if (warp == 10) {
    // ======================= weight loader =======================
    if (lane == 0) {
      if (resident) {
        // n_tiles == 1 in resident mode: one expect_tx for the whole layer, nkb bulk copies
        mbar_expect_tx(b_full(0), img_bytes * (uint32_t)nkb);
        for (int kb = 0; kb < nkb; ++kb)
          bulk_g2s(smem_u32(smem_b + (size_t)kb * img_bytes), p.wp + (size_t)kb * img_bytes, img_bytes, b_full(0));
      } else {
        int c = 0;
        while (seq.valid()) {
          const int nt = seq.item / m_tiles;
          const int sb = c % TC_SB;
          mbar_wait(b_empty(sb), (uint32_t)(((c / TC_SB) & 1) ^ 1));
          mbar_expect_tx(b_full(sb), img_bytes);
          bulk_g2s(smem_u32(smem_b + (size_t)sb * img_bytes), p.wp + ((size_t)nt * nkb + seq.kb) * img_bytes, img_bytes,
                   b_full(sb));
          seq.next();
          ++c;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 12) {
    // ======================= epilogue: TMEM -> bias/LeakyReLU/alpha/addend -> NCHW slice =======================
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int tcount = 0;
    while (seq.valid()) {
      const int buf = NBUF == 2 ? (tcount & 1) : 0;
      const int use = NBUF == 2 ? (tcount >> 1) : tcount;
      const int mt = seq.item % m_tiles, nt = seq.item / m_tiles;
      const float* bias_s = bias_all + nt * N;
      mbar_wait(acc_full(buf), (uint32_t)(use & 1));
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const long long mg = (long long)mt * TC_BM2 + h * 128 + q * 32 + lane;
        const bool m_ok = mg < p.M;
        int ob = 0, opix = 0;
        if (m_ok) { ob = (int)(mg / HWo); opix = (int)(mg - (long long)ob * HWo); }
        const uint32_t acc_addr = lane_addr + (uint32_t)(TC_ACC_COL + (buf * 2 + h) * acc_stride);
        const float* ap = p.addend ? p.addend + (size_t)ob * p.a_bs + opix : nullptr;
        float* yp = p.y + (size_t)ob * p.y_bs + opix;
        for (int c0 = 0; c0 < N; c0 += 16) {
          const int nb = nt * N + c0;
          // residual / skip-connection operand: 16 independent loads in flight before the accumulator is touched
          float add[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            add[j] = (ap != nullptr && m_ok && nb + j < p.Cout) ? __ldg(ap + (size_t)(nb + j) * HWo) : 0.f;
          uint32_t r[16];
          tmem_ld16(acc_addr + (uint32_t)c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (m_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (nb + j < p.Cout) {
                float val = __uint_as_float(r[j]) + bias_s[c0 + j];
                val = fmaf(leaky(val, p.slope), p.alpha, add[j]);
                yp[(size_t)(nb + j) * HWo] = val;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      ++tcount;
      seq.item += seq.stride;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ weight packer
// One thread per (n tile, k block, plane, row n, k j): OIHW -> swizzled K-major image (see make_b_desc).
__global__ void pack_tc_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, int Cout, int Cin, int ks,
                               int n_tile, int n_tiles, int cchunks, int passes, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int T = ks * ks;
  const int planes = passes == 3 ? 2 : 1;
  int j = (int)(i % TC_CK);
  long long r = i / TC_CK;
  int n = (int)(r % n_tile); r /= n_tile;
  int plane = (int)(r % planes); r /= planes;
  int kb = (int)(r % (cchunks * T));
  int nt = (int)(r / (cchunks * T));
  int cc = kb / T, tap = kb - cc * T;
  int c = cc * TC_CK + j, ng = nt * n_tile + n;
  float v = 0.f;
  if (c < Cin && ng < Cout) v = __ldg(w + ((size_t)ng * Cin + c) * T + tap);
  uint32_t hi = to_tf32(v);
  uint32_t val = plane == 0 ? hi : to_tf32(v - __uint_as_float(hi));
  size_t img = (size_t)planes * n_tile * 128;
  size_t base = ((size_t)nt * cchunks * T + kb) * img + (size_t)plane * n_tile * 128;
  // row n: 128 bytes; 16-byte chunk (j/4) XOR (n % 8)   [Swizzle<3,4,3>]
  size_t off = (size_t)n * 128 + (size_t)(((j >> 2) ^ (n & 7)) << 4) + (size_t)(j & 3) * 4;
  *reinterpret_cast<uint32_t*>(out + base + off) = val;
}

bool tc_supported(int Cout, int Cin, int ks, int stride, int dil) {
  (void)stride; (void)dil;
  if (ks != 1 && ks != 3) return false;
  // Every layer shape of the PWC family maps: Cout is padded to a multiple of 16 (UMMA N), the channel tail of a K
  // block to a multiple of 8 (UMMA K).  Thin layers (Cout = 1, 2, 3, 9; Cin = 3, 11) waste tensor throughput they do
  // not need — they are bound by the activation gather, which is identical for any N.
  return Cout >= 1 && Cout <= 256 && Cin >= 1;
}

size_t tc_packed_bytes(int Cout, int Cin, int ks, int math) {
  int passes = math == IRR_MATH_TC_3XTF32 ? 3 : 1;
  TcGeom g = tc_geom(Cout, Cin, ks, passes);
  return (size_t)g.n_tiles * g.nkb * g.img_bytes;
}

int tc_pack(const float* w, void* out, int Cout, int Cin, int ks, int math, cudaStream_t st) {
  int passes = math == IRR_MATH_TC_3XTF32 ? 3 : 1;
  TcGeom g = tc_geom(Cout, Cin, ks, passes);
  long long total = (long long)g.n_tiles * g.nkb * (passes == 3 ? 2 : 1) * g.n_tile * TC_CK;
  pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, (uint8_t*)out, Cout, Cin, ks, g.n_tile, g.n_tiles,
                                                                  g.cchunks, passes, total);
  return check_launch("irr_conv2d_pack_weights");
}

template <int KS, int PASSES>
static int launch_tc(const TcArgs& a, size_t smem, cudaStream_t st) {
  static SmemAttrCache attr = {};
  if (int rc = ensure_dyn_smem(conv_tc_kernel<KS, PASSES>, smem, attr, "irr_conv2d_fwd")) return rc;
  long long items = ((a.M + TC_BM2 - 1) / TC_BM2) * a.n_tiles;
  int grid = (int)(items < sm_count() ? items : sm_count());  // persistent: one CTA per SM (TMEM: 512 columns each)
  conv_tc_kernel<KS, PASSES><<<grid, TC_THREADS2, smem, st>>>(a);
  return check_launch("irr_conv2d_fwd");
}

int tc_conv(const float* x, long long x_bs, const void* w, const float* bias, const float* addend, long long a_bs,
            float* y, long long y_bs, int B, int Cin, int H, int W, int Cout, int ks, int stride, int dil, float slope,
            float alpha, int math, cudaStream_t st) {
  int passes = math == IRR_MATH_TC_3XTF32 ? 3 : 1;
  TcGeom g = tc_geom(Cout, Cin, ks, passes);
  TcArgs a;
  a.x = x; a.x_bs = x_bs; a.wp = (const uint8_t*)w; a.bias = bias; a.addend = addend; a.a_bs = a_bs; a.y = y; a.y_bs = y_bs;
  a.B = B; a.Cin = Cin; a.H = H; a.W = W; a.Cout = Cout; a.stride = stride; a.dil = dil; a.ks = ks;
  a.pad = ((ks - 1) * dil) / 2;
  a.Ho = (H + 2 * a.pad - dil * (ks - 1) - 1) / stride + 1;
  a.Wo = (W + 2 * a.pad - dil * (ks - 1) - 1) / stride + 1;
  a.n_tile = g.n_tile; a.n_tiles = g.n_tiles; a.cchunks = g.cchunks; a.nkb = g.nkb; a.passes = passes;
  a.M = (long long)B * a.Ho * a.Wo;
  a.slope = slope; a.alpha = alpha;
  size_t total_b = (size_t)g.nkb * g.img_bytes;
  a.resident = (g.n_tiles == 1 && total_b <= TC_RESIDENT_MAX) ? 1 : 0;
  a.b_smem_bytes = (unsigned)(a.resident ? total_b : (size_t)TC_SB * g.img_bytes);
  size_t smem = (size_t)a.b_smem_bytes + 26 * 8 + 256 * 4 + 64;
  if (ks == 1) return passes == 3 ? launch_tc<1, 3>(a, smem, st) : launch_tc<1, 1>(a, smem, st);
  return passes == 3 ? launch_tc<3, 3>(a, smem, st) : launch_tc<3, 1>(a, smem, st);
}

}  // namespace irr
