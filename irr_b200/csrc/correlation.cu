// Cost volume (A1) fused with LeakyReLU (A3) and, optionally, with the bilinear warp + hard mask that feeds it (A2).
//
// Reference semantics: models/pwc_modules.py:42-62 (compute_cost_volume) == Correlation(pad 4, k 1, md 4, s 1/1) of
// models/correlation_package/correlation_cuda_kernel.cu:41-114:  out[b, (dy+4)*9+(dx+4), y, x] =
// (1/C) sum_c f1[b,c,y,x] * f2w[b,c,y+dy,x+dx], zero outside the image.
//
// B200 design (HBM-bound op sitting at the fp32-FMA ridge) — persistent, warp-specialised:
//   * grid = min(#tiles, #SMs) persistent CTAs of 16 warps; each CTA walks 8 x 32 output tiles of the (2B) batch.
//   * warps 0-8 = COMPUTE: warp w owns displacement row dy = w-4, lane = (tile row r, 8-pixel strip s); 8 px x 9 dx = 72
//     fp32 accumulators per thread, so one channel step is 2 + 4 LDS.128 for 72 FFMA — the register tile that keeps
//     the FMA pipe, not the LSU, the limiter (the reference kernel does 1 FMA per 2 global loads and re-reads f1 81x).
//   * warps 9-15 = PRODUCERS: they stage 8-channel chunks of the f1 tile [8][8][36] and the f2 halo tile [8][16][44]
//     into a 3-slot shared-memory ring (mbarrier full/empty) with cp.async (16-byte, zero-fill = the zero padding),
//     one chunk ahead of the compute warps and across tile boundaries, so the epilogue of tile t overlaps the loads
//     of tile t+1.  Fused variant: the sample coordinates / bilinear weights / hard mask of each halo position are
//     computed ONCE per tile (bit-exact recipe in common.cuh) and kept in registers; the tile's source footprint of f2
//     (24 x 48 texels per channel, placed from the min/max of the sample coordinates) is copied asynchronously into a
//     second ring and the four taps are read from shared memory; a tile whose flow is too divergent for the window
//     gathers from global memory instead.  The warped tensor never exists in HBM.
//   * row pitches 36 / 44 words (== 4, 12 mod 32) make every quarter-warp LDS.128 (8 rows, same strip) conflict-free.
//   * global reads are coalesced along W; f2's 2.5x halo re-read is served by the 126 MB L2 (the largest level-4 map
//     of cfg 3 is 28.6 MB), so DRAM traffic stays at the algorithmic B*H*W*(8C+324) bytes.
//   * epilogue: / C, LeakyReLU, 16-byte stores into the caller's channel slice of the estimator input buffer.
// No tensor cores by design (a 9x9-window dot product, not a GEMM).
#include <stdlib.h>

#include "common.cuh"

// This file is compiled twice: as is (8-row tiles, every entry point) and from correlation7.cu with IRR_CORR_TH = 7
// (namespace irr::corr7, only launch_corr<true> is used).  With 7-row tiles the 63 (row, displacement-row) pairs of a tile
// make exactly 8 compute warps — two per scheduler — where the 72 pairs of an 8-row tile make 9, which load the four
// schedulers 3:2:2:2 and leave the kernel waiting on scheduler 0.
#ifndef IRR_CORR_TH
#define IRR_CORR_TH 8
#endif
#ifndef IRR_CORR_FP_H
#define IRR_CORR_FP_H 24   // rows of the fused kernel's source-footprint window
#endif
#ifndef IRR_CORR_NFS
#define IRR_CORR_NFS 2     // footprint ring depth
#endif
#ifndef IRR_CORR_NS
#define IRR_CORR_NS corr8
#else
#define IRR_CORR_VARIANT_ONLY 1   // secondary translation unit: only launch_corr_fused_variant is used
#endif

namespace irr {
namespace IRR_CORR_NS {

constexpr int TH = IRR_CORR_TH, TW = 32, MD = 4, ND = 9, PX = 8;
constexpr int F1_P = 36;                    // f1 row pitch (floats)
constexpr int F2_H = TH + 2 * MD;           // 16 (15)
constexpr int F2_WV = TW + 2 * MD;          // 40 valid halo columns
constexpr int F2_P = 44;                    // f2 row pitch (floats)
constexpr int CC = 8;                       // channels per chunk
constexpr int NPAIR = TH * ND;              // (tile row, displacement row) pairs: 72 (63)
constexpr int NCOMP = ((NPAIR * (TW / PX) + 31) / 32) * 32;   // compute threads: 288 = 9 warps (256 = 8 warps)
constexpr int NPROD = 224;                  // 7 producer warps
constexpr int CORR_THREADS = NCOMP + NPROD; // 512 (480)
constexpr int CORR_STAGES = 3;
constexpr int F1_ELEMS = CC * TH * F1_P;    // 2304
constexpr int F2_ELEMS = CC * F2_H * F2_P;  // 5632
constexpr int STAGE_ELEMS = F1_ELEMS + F2_ELEMS;
constexpr int NHALO = F2_H * F2_WV;         // 640
constexpr int NF1 = TH * TW;                // 256
constexpr int FP_H = IRR_CORR_FP_H, FP_W = 48; // fused variant: source footprint window per channel (texels)
constexpr int FP_ELEMS = CC * FP_H * FP_W;     // 9216 floats per ring slot
// bf16 INPUT STORAGE (dtype_in = 1, TMA kernels only): the ring slots keep their fp32-sized regions and byte offsets, the
// bf16 tiles use the front of them.  Row pitches in bf16 elements are multiples of 8 (TMA box rows of whole 16-byte units):
// f1 [8][TH][40]; plain f2 [8][16][48] with the box origin at x0 - 8 (a 16-byte aligned inner coordinate), i.e. halo column
// hx sits at tile column hx + 4; fused: the footprint [8][FP_H][48] is bf16, the warped halo tile the samplers write stays fp32.
constexpr int F1_P16 = 40, F2_P16 = 48, F2_SKIP16 = 4;
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
template <bool BF>
__device__ __forceinline__ float ld_in(const void* base, long long i) {   // element i of an fp32 / bf16 array
  if (BF) return __uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(base)[i] << 16);
  return reinterpret_cast<const float*>(base)[i];
}
template <bool BF>
__device__ __forceinline__ float ldg_in(const void* base, long long i) {
  if (BF) return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(base) + i) << 16);
  return __ldg(reinterpret_cast<const float*>(base) + i);
}
constexpr int CORR_SMEM_PLAIN = CORR_STAGES * STAGE_ELEMS * 4 + 64;
constexpr int CORR_SMEM_FUSED = CORR_SMEM_PLAIN + CORR_STAGES * FP_ELEMS * 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");
}

__device__ float corr_zero_page[32];  // statically zero: source of dead halo positions

struct ProdPos {  // one halo (or f1) position owned by a producer thread
  int soff;       // offset inside the stage for channel 0 (floats); < 0 => unused slot
  int goff;       // global offset (y*W + x), clamped; for fused: clamped (y0*W + x0)
  int dx, dy;     // fused: +1 / +W when the second column / row is a distinct in-range texel, else 0
  float w00, w01, w10, w11;  // fused: bilinear weights with mask and bounds folded in; plain: w00 = 1 if in image else 0
};

// ---------------------------------------------------------------------------------------------------------
// COMPUTE role (warps 0-8 of either kernel): consumes the NS-slot ring of [f1 tile | f2 halo tile] chunks.
// Role cycle counters (block 0, one thread per role), filled when the launch is made with IRR_CORR_CTR=1 and read back
// with irrdbg_corr_counters(): [0] compute wait-full [1] set-up [2] epilogue [3] total | [8] issuer wait-table
// [9] wait-empty [10] wait-fpempty [11] total | [16] sampler wait-table [17] wait-fpfull [18] wait-empty [19] total.
__device__ unsigned long long corr_ctr[32];
__device__ __forceinline__ void mbar_wait_ctr(uint32_t bar, uint32_t parity, bool on, int idx) {
  if (on) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    corr_ctr[idx] += (unsigned long long)(clock64() - t0);
  } else {
    mbar_wait(bar, parity);
  }
}

struct NoTileHook {
  __device__ __forceinline__ void setup(int, int) const {}
};

// Channel split for launches with far fewer tiles than SMs (the 7x16 ... 14x32 pyramid levels: 16-32 tiles, up to 25
// serial 8-channel chunks each): the chunks of a tile are dealt to `ksplit` CTAs ("virtual tiles" vt = tile * ksplit +
// ks, chunk range [ks*cps, (ks+1)*cps)), each stores its RAW partial sums to workspace slice ks, and corr_split_finish
// adds the slices in a fixed order, scales by 1/C and applies the LeakyReLU — deterministic, like the convs' split-K.
struct CSplit {
  int ksplit, cps;
  float* ws; long long ws_stride;   // [ksplit][B][81][H][W]
};

template <int NS, bool PREFETCH, class Hook, bool SPLIT = false, bool BF1 = false, bool BF2 = false>
__device__ __forceinline__ void corr_compute(const Hook& hook, const float* smem, uint32_t full0, uint32_t empty0,
                                             const float* __restrict__ f1, long long f1_bs,
                                             const float* __restrict__ f2, long long f2_bs, float* __restrict__ out,
                                             long long out_bs, int B, int C, int H, int W, int P, int shift, float slope,
                                             int vec_ok, int tiles_x, int tiles_y, int ntiles, int ctr = 0,
                                             CSplit sp = CSplit{1, 0, nullptr, 0}) {
  const int tid = threadIdx.x;
  const int HW = H * P;   // channel stride: rows are stored with pitch P >= W (include/irr_b200.h)
  const int nchunks = (C + CC - 1) / CC;
  const int ksplit = SPLIT ? sp.ksplit : 1;
  const int nvt = ntiles * ksplit;
  const bool con = ctr && blockIdx.x == 0 && tid == 0;
  const long long ct0 = con ? clock64() : 0;
  // Thread -> (tile row r, displacement row dyi, 8-pixel strip s8).  The 72 (r, dyi) pairs are dealt to the 9 warps
  // sorted by the f2 halo row they read (h = r + dyi), 8 pairs x 4 strips per warp: a warp then touches only 2-4
  // distinct f2 rows and ~5 f1 rows, and lanes that share a row and strip read the SAME 16 bytes — one shared-memory
  // wavefront serves them all (broadcast).  With warp = dy (lane = row) every lane read its own 16 bytes and the
  // 6 LDS.128 per channel cost 24 wavefronts per warp against 18 issue slots of FFMA: shared memory, not the FMA
  // pipe, was the limiter.  Sorted, the same loads cost ~12 wavefronts.
  const int lane = tid & 31;
  const int s8 = lane & 3;
  int r, dyi;
  bool pair_ok;
  {
    // p-th pair in (h, r) order; halo row h holds min(h, TH-1, ND-1, TH+ND-2-h) + 1 pairs.  A 7-row tile has 63 pairs:
    // the last four lanes of warp 7 shadow pair 0 (their loads stay in bounds, they store nothing).
    int p = (tid >> 5) * 8 + (lane >> 2), h = 0;
    pair_ok = p < NPAIR;
    if (!pair_ok) p = 0;
    for (;;) {
      const int cnt = min(min(h, TH - 1), min(ND - 1, TH + ND - 2 - h)) + 1;
      if (p < cnt) break;
      p -= cnt;
      ++h;
    }
    r = max(h - (ND - 1), 0) + p;
    dyi = h - r;  // 0..8 -> dy = dyi - 4
  }
  int gchunk = 0, it = 0;
  if ((int)blockIdx.x < nvt) hook.setup((int)blockIdx.x / ksplit, 0);
  for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x, ++it) {
    const int tile = SPLIT ? vt / ksplit : vt, ks = SPLIT ? vt - tile * ksplit : 0;
    const int c_lo = SPLIT ? ks * sp.cps : 0, c_hi = SPLIT ? min(nchunks, c_lo + sp.cps) : nchunks;
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int b = tile / (tiles_x * tiles_y);
    const int y0 = ty * TH, x0 = tx * TW;
    {
      const long long t0 = con ? clock64() : 0;
      if (vt + (int)gridDim.x < nvt) hook.setup((vt + (int)gridDim.x) / ksplit, it + 1);  // accumulators are dead here
      if (con) corr_ctr[1] += (unsigned long long)(clock64() - t0);
    }
    float acc[ND][PX];
#pragma unroll
    for (int d = 0; d < ND; ++d)
#pragma unroll
      for (int p = 0; p < PX; ++p) acc[d][p] = 0.f;

    if (PREFETCH) {  // The compute warps spend most of a tile waiting on the producers: use them to pull the NEXT tile's f1 rows and
       // (un-warped) f2 neighbourhood into L2, so the producers' gathers find their lines there instead of in DRAM.
      const int ntile = tile + (int)gridDim.x;
      if (ntile < ntiles) {
        const int ntx = ntile % tiles_x, nty = (ntile / tiles_x) % tiles_y, nb = ntile / (tiles_x * tiles_y);
        int nb2 = nb + shift;
        if (nb2 >= B) nb2 -= B;
        const float* pf1 = f1 + (size_t)nb * f1_bs;
        const float* pf2 = f2 + (size_t)nb2 * f2_bs;
        const int ny0 = nty * TH, nx0 = ntx * TW;
        for (int i = tid; i < C * 40; i += NCOMP) {
          const int c = i / 40, rr = i - c * 40;
          const float* a;
          if (rr < 8) {
            const int y = min(ny0 + rr, H - 1);
            a = pf1 + (size_t)c * HW + (size_t)y * P + min(nx0, W - 1);
          } else {
            const int k = rr - 8;
            const int y = min(max(ny0 - MD + (k >> 1), 0), H - 1);
            const int x = min(max(nx0 - MD + (k & 1) * 32, 0), W - 1);
            a = pf2 + (size_t)c * HW + (size_t)y * P + x;
          }
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
    }

    for (int ci = c_lo; ci < c_hi; ++ci, ++gchunk) {
      const int s = gchunk % NS;
      const uint32_t ph = (uint32_t)((gchunk / NS) & 1);
      mbar_wait_ctr(full0 + 8u * s, ph, con, 0);
      const float* f1s = smem + s * STAGE_ELEMS;
      const float* f2s = f1s + F1_ELEMS;
#pragma unroll 2
      for (int cc = 0; cc < CC; ++cc) {
        float a[PX], bv[PX + 2 * MD];
        if (BF1) {   // 8 bf16 pixels = one 16-byte load; bf16 -> fp32 is a shift / a mask
          const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(f1s) + (cc * TH + r) * F1_P16 +
                                                          s8 * PX);
          a[0] = bf_lo(t.x); a[1] = bf_hi(t.x); a[2] = bf_lo(t.y); a[3] = bf_hi(t.y);
          a[4] = bf_lo(t.z); a[5] = bf_hi(t.z); a[6] = bf_lo(t.w); a[7] = bf_hi(t.w);
        } else {
          const float4* ap = reinterpret_cast<const float4*>(f1s + (cc * TH + r) * F1_P + s8 * PX);
          float4 t0 = ap[0], t1 = ap[1];
          a[0] = t0.x; a[1] = t0.y; a[2] = t0.z; a[3] = t0.w; a[4] = t1.x; a[5] = t1.y; a[6] = t1.z; a[7] = t1.w;
        }
        if (BF2) {   // 16 bf16 halo pixels starting 8 bytes into a 16-byte unit: four 8-byte loads
          const uint2* bp = reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned short*>(f2s) +
                                                           (cc * F2_H + r + dyi) * F2_P16 + F2_SKIP16 + s8 * PX);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint2 t = bp[q];
            bv[4 * q] = bf_lo(t.x); bv[4 * q + 1] = bf_hi(t.x); bv[4 * q + 2] = bf_lo(t.y); bv[4 * q + 3] = bf_hi(t.y);
          }
        } else {
          const float4* bp = reinterpret_cast<const float4*>(f2s + (cc * F2_H + r + dyi) * F2_P + s8 * PX);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 t = bp[q];
            bv[4 * q] = t.x; bv[4 * q + 1] = t.y; bv[4 * q + 2] = t.z; bv[4 * q + 3] = t.w;
          }
        }
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
          for (int p = 0; p < PX; ++p) acc[d][p] = fmaf(a[p], bv[p + d], acc[d][p]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8u * s);
    }

    // ---- epilogue: mean over channels (pwc_modules.py:59 / .cu:107 divide by nelems), LeakyReLU (IRR_PWC.py:94-95)
    const long long te0 = con ? clock64() : 0;
    const int gy = y0 + r;
    const int gx = x0 + s8 * PX;
    if (SPLIT) {
      // channel-split launch: raw partial sums of chunks [c_lo, c_hi) into this split's workspace slice
      if (pair_ok && gy < H && gx < W) {
        float* op = sp.ws + (size_t)ks * sp.ws_stride + ((size_t)b * (ND * ND) + (size_t)(dyi * ND)) * HW + (size_t)gy * P + gx;
        if ((P & 3) == 0 && gx + PX <= W) {
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            float4* q = reinterpret_cast<float4*>(op + (size_t)d * HW);
            q[0] = make_float4(acc[d][0], acc[d][1], acc[d][2], acc[d][3]);
            q[1] = make_float4(acc[d][4], acc[d][5], acc[d][6], acc[d][7]);
          }
        } else {
#pragma unroll
          for (int d = 0; d < ND; ++d)
#pragma unroll
            for (int p = 0; p < PX; ++p)
              if (gx + p < W) op[(size_t)d * HW + p] = acc[d][p];
        }
      }
    } else if (pair_ok && gy < H && gx < W) {
      const float inv_c = 1.0f / (float)C;  // mean over channels as one multiply (<= 1 ulp from the reference's divide)
      float* op = out + (size_t)b * out_bs + (size_t)(dyi * ND) * HW + (size_t)gy * P + gx;
      // scale in place first, then issue the stores back to back (a temporary per displacement makes every store
      // wait for the previous one to release its registers).  slope in [0, 1]: leaky(v) == max(v, v * slope).
      if (slope >= 0.f && slope <= 1.f) {
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const float v = acc[d][p] * inv_c;
            acc[d][p] = fmaxf(v, v * slope);
          }
      } else {
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
          for (int p = 0; p < PX; ++p) acc[d][p] = leaky(acc[d][p] * inv_c, slope);
      }
      if (vec_ok == 2 && gx + PX <= W) {
        // 32-byte stores (sm_100 st.global.v8.f32): one full sector per lane per instruction, 9 stores per thread
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op + (size_t)d * HW), "f"(acc[d][0]),
                       "f"(acc[d][1]), "f"(acc[d][2]), "f"(acc[d][3]), "f"(acc[d][4]), "f"(acc[d][5]), "f"(acc[d][6]),
                       "f"(acc[d][7])
                       : "memory");
        }
      } else if (vec_ok && gx + PX <= W) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          float4* q = reinterpret_cast<float4*>(op + (size_t)d * HW);
          q[0] = make_float4(acc[d][0], acc[d][1], acc[d][2], acc[d][3]);
          q[1] = make_float4(acc[d][4], acc[d][5], acc[d][6], acc[d][7]);
        }
      } else {
#pragma unroll
        for (int d = 0; d < ND; ++d)
#pragma unroll
          for (int p = 0; p < PX; ++p)
            if (gx + p < W) op[(size_t)d * HW + p] = acc[d][p];
      }
    }
    if (con) corr_ctr[2] += (unsigned long long)(clock64() - te0);
  }
  if (con) corr_ctr[3] += (unsigned long long)(clock64() - ct0);
}

template <bool FUSED>
__global__ void __launch_bounds__(CORR_THREADS, 1)
    corr_kernel(const float* __restrict__ f1, long long f1_bs, const float* __restrict__ f2, long long f2_bs,
                const float* __restrict__ flow, long long flow_bs, float* __restrict__ out, long long out_bs,
                GridArgs g, int B, int C, int H, int W, int P, int shift, float slope, int vec_ok, int vec_in, int tiles_x,
                int tiles_y, int ntiles) {
  extern __shared__ __align__(16) float smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CORR_STAGES * STAGE_ELEMS);
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (CORR_STAGES + s); };

  const int tid = threadIdx.x;
  const int HW = H * P;   // channel stride (pitched rows)
  const int nchunks = (C + CC - 1) / CC;
  if (tid == 0) {
    for (int s = 0; s < CORR_STAGES; ++s) {
      mbar_init(full(s), NPROD / 32);   // one elected arrival per producer warp
      mbar_init(empty(s), NCOMP / 32);  // one elected arrival per compute warp
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (tid >= NCOMP) {
    // ============================== PRODUCERS ==============================
    // Staging is asynchronous: every plain copy is a cp.async (16 B when rows are 16-byte aligned, else 4 B) with
    // zero-fill for the out-of-image part, issued two chunks ahead, so no producer thread ever waits on a global load.
    //   * f1 tile and, without a warp, the f2 halo tile land directly in their slot of the ring;
    //   * with the warp fused in, the tile's SOURCE FOOTPRINT of f2 (the halo box displaced by the flow, 24 x 48 texels
    //     per channel, positioned per tile from the min/max of the sample coordinates) is copied into a second ring and
    //     the four bilinear taps are read from shared memory (LDS latency instead of a DRAM round trip per tap group);
    //     tiles whose flow field is too divergent for the footprint fall back to gathering from global memory.
    const int pt = tid - NCOMP;
    const int lane = tid & 31;
    float* fpr = smem + CORR_STAGES * STAGE_ELEMS + 16;           // footprint ring (FUSED only), after the barriers
    int* red = reinterpret_cast<int*>(smem + CORR_STAGES * STAGE_ELEMS + 12);  // 4 ints: ymin, ymax, xmin, xmax
    int gchunk = 0;
    auto prod_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory"); };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const int y0 = ty * TH, x0 = tx * TW;
      int b2 = b + shift;
      if (b2 >= B) b2 -= B;
      const float* f1b = f1 + (size_t)b * f1_bs;
      const float* f2b = f2 + (size_t)b2 * f2_bs;

      // ---- per-tile: sample taps of the (<= 3) halo positions this thread owns
      ProdPos hp[3];
      int ya_[3], xa_[3];
      int lo_y = 1 << 30, hi_y = -1, lo_x = 1 << 30, hi_x = -1;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int h = pt + k * NPROD;
        ProdPos q;
        q.soff = -1; q.goff = 0; q.dx = 0; q.dy = 0; q.w00 = q.w01 = q.w10 = q.w11 = 0.f;
        ya_[k] = 0; xa_[k] = 0;
        if (h < NHALO) {
          const int hr = h / F2_WV, hx = h - hr * F2_WV;
          const int gy = y0 - MD + hr, gx = x0 - MD + hx;
          q.soff = F1_ELEMS + hr * F2_P + hx;
          if (FUSED && gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const float* fl = flow + (size_t)b * flow_bs + (size_t)gy * P + gx;
            float ix, iy;
            sample_coords(g, __ldg(fl), __ldg(fl + HW), gx, gy, W, H, ix, iy);
            Taps tp = make_taps(ix, iy, W, H);
            if (tp.mask != 0.f) {
              const int xa = min(max(tp.x0, 0), W - 1), xb = min(max(tp.x0 + 1, 0), W - 1);
              const int ya = min(max(tp.y0, 0), H - 1), yb = min(max(tp.y0 + 1, 0), H - 1);
              ya_[k] = ya; xa_[k] = xa;
              q.goff = ya * P + xa;
              q.dx = (xb != xa) ? 1 : 0;
              q.dy = (yb != ya) ? 1 : 0;  // in rows; scaled by the row pitch of whichever source is sampled
              // a clamped (out-of-range) tap has zero weight (make_taps), so aliasing it onto its in-range neighbour's
              // address is harmless: the four reads are always in bounds.
              q.w00 = tp.w00; q.w01 = tp.w01; q.w10 = tp.w10; q.w11 = tp.w11;
              lo_y = min(lo_y, ya); hi_y = max(hi_y, yb); lo_x = min(lo_x, xa); hi_x = max(hi_x, xb);
            }
          }
        }
        hp[k] = q;
      }
      // ---- FUSED: does the tile's sample footprint fit the shared-memory window?
      bool foot = false;
      int oy = 0, ox = 0;
      if (FUSED) {
        if (pt == 0) { red[0] = 1 << 30; red[1] = -1; red[2] = 1 << 30; red[3] = -1; }
        prod_sync();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
          lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
        }
        if (lane == 0) { atomicMin(&red[0], lo_y); atomicMax(&red[1], hi_y); atomicMin(&red[2], lo_x); atomicMax(&red[3], hi_x); }
        prod_sync();
        const int ymin = red[0], ymax = red[1], xmin = red[2], xmax = red[3];
        const bool any_live = ymax >= ymin;
        oy = any_live ? ymin : 0; ox = any_live ? (xmin & ~3) : 0;
        foot = vec_in && (!any_live || ((ymax - oy) < FP_H && (xmax - ox) < FP_W));
        prod_sync();  // red[] is re-initialised by the next tile
      }

      // ---- copy issue for one chunk (ring slot s), cooperative over the 224 producer threads
      auto issue_copies = [&](int ci, int s) {
        const int c0 = ci * CC;
        float* st = smem + s * STAGE_ELEMS;
        if (vec_in) {
          for (int j = pt; j < CC * TH * (TW / 4); j += NPROD) {  // f1: 8 ch x 8 rows x 8 float4
            const int xq = j & 7, rr = (j >> 3) & 7, cc = j >> 6;
            const int gy = y0 + rr, gx = x0 + xq * 4, c = c0 + cc;
            const bool ok = c < C && gy < H && gx < W;
            cp_async16(smem_u32(st + (cc * TH + rr) * F1_P + xq * 4), ok ? f1b + (size_t)c * HW + (size_t)gy * P + gx : f1b, ok);
          }
          if (!FUSED) {
            for (int j = pt; j < CC * F2_H * (F2_WV / 4); j += NPROD) {  // f2 halo: 8 ch x 16 rows x 10 float4
              const int cc = j / (F2_H * 10), r2 = j - cc * (F2_H * 10);
              const int hr = r2 / 10, xq = r2 - hr * 10;
              const int gy = y0 - MD + hr, gx = x0 - MD + xq * 4, c = c0 + cc;
              const bool ok = c < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
              cp_async16(smem_u32(st + F1_ELEMS + (cc * F2_H + hr) * F2_P + xq * 4), ok ? f2b + (size_t)c * HW + (size_t)gy * P + gx : f2b, ok);
            }
          } else if (foot) {
            float* fp = fpr + s * FP_ELEMS;
            for (int j = pt; j < CC * FP_H * (FP_W / 4); j += NPROD) {  // footprint: 8 ch x 24 rows x 12 float4
              const int cc = j / (FP_H * (FP_W / 4)), r2 = j - cc * (FP_H * (FP_W / 4));
              const int fr = r2 / (FP_W / 4), xq = r2 - fr * (FP_W / 4);
              const int gy = oy + fr, gx = ox + xq * 4, c = c0 + cc;
              const bool ok = c < C && gy < H && gx < W;  // oy, ox >= 0
              cp_async16(smem_u32(fp + (cc * FP_H + fr) * FP_W + xq * 4), ok ? f2b + (size_t)c * HW + (size_t)gy * P + gx : f2b, ok);
            }
          }
        } else {
          for (int j = pt; j < CC * TH * TW; j += NPROD) {
            const int xx = j & 31, rr = (j >> 5) & 7, cc = j >> 8;
            const int gy = y0 + rr, gx = x0 + xx, c = c0 + cc;
            const bool ok = c < C && gy < H && gx < W;
            cp_async4(smem_u32(st + (cc * TH + rr) * F1_P + xx), ok ? f1b + (size_t)c * HW + (size_t)gy * P + gx : f1b, ok);
          }
          if (!FUSED) {
            for (int j = pt; j < CC * NHALO; j += NPROD) {
              const int cc = j / NHALO, h2 = j - cc * NHALO;
              const int hr = h2 / F2_WV, hx = h2 - hr * F2_WV;
              const int gy = y0 - MD + hr, gx = x0 - MD + hx, c = c0 + cc;
              const bool ok = c < C && gy >= 0 && gy < H && gx >= 0 && gx < W;
              cp_async4(smem_u32(st + F1_ELEMS + (cc * F2_H + hr) * F2_P + hx), ok ? f2b + (size_t)c * HW + (size_t)gy * P + gx : f2b, ok);
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };

      // ---- pipeline: copies for chunk ci+1 are in flight while chunk ci is finished (sampled) and published
      {
        const int s0 = gchunk % CORR_STAGES;
        mbar_wait(empty(s0), (uint32_t)(((gchunk / CORR_STAGES) & 1) ^ 1));
        issue_copies(0, s0);
      }
      for (int ci = 0; ci < nchunks; ++ci) {
        const int gc = gchunk + ci;
        const int s = gc % CORR_STAGES;
        float* st = smem + s * STAGE_ELEMS;
        if (ci + 1 < nchunks) {
          const int s1 = (gc + 1) % CORR_STAGES;
          mbar_wait(empty(s1), (uint32_t)((((gc + 1) / CORR_STAGES) & 1) ^ 1));
          issue_copies(ci + 1, s1);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (FUSED) {
          const int c0 = ci * CC;
          if (foot) {
            prod_sync();  // every producer's share of the footprint has landed
            const float* fp = fpr + s * FP_ELEMS;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (hp[k].soff >= 0) {
                const bool live = hp[k].w00 != 0.f || hp[k].w01 != 0.f || hp[k].w10 != 0.f || hp[k].w11 != 0.f;
                // dead positions (outside the image / masked out) sample nothing: keep the address inside the window
                const float* p = fp + (live ? (ya_[k] - oy) * FP_W + (xa_[k] - ox) : 0);
                const int dx = live ? hp[k].dx : 0, dy = live ? hp[k].dy * FP_W : 0;
#pragma unroll
                for (int cc = 0; cc < CC; ++cc) {
                  const float* pc = p + cc * (FP_H * FP_W);
                  float a = __fmul_rn(pc[0], hp[k].w00);  // tap order of grid_sampler_2d
                  a = fmaf(pc[dx], hp[k].w01, a);
                  a = fmaf(pc[dy], hp[k].w10, a);
                  a = fmaf(pc[dy + dx], hp[k].w11, a);
                  st[hp[k].soff + cc * (F2_H * F2_P)] = live ? a : 0.f;
                }
              }
            }
          } else {  // divergent flow (or unaligned rows): gather the taps from global memory
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              if (hp[k].soff >= 0) {
                const bool live = hp[k].w00 != 0.f || hp[k].w01 != 0.f || hp[k].w10 != 0.f || hp[k].w11 != 0.f;
                const float* p = f2b + (size_t)c0 * HW + hp[k].goff;
                const int dx = hp[k].dx, dy = hp[k].dy * P;
                float t[4][CC];
#pragma unroll
                for (int cc = 0; cc < CC; ++cc) {
                  const bool okc = live && (c0 + cc < C);
                  const float* pc = p + (size_t)cc * HW;
                  t[0][cc] = okc ? __ldg(pc) : 0.f;
                  t[1][cc] = okc ? __ldg(pc + dx) : 0.f;
                  t[2][cc] = okc ? __ldg(pc + dy) : 0.f;
                  t[3][cc] = okc ? __ldg(pc + dy + dx) : 0.f;
                }
#pragma unroll
                for (int cc = 0; cc < CC; ++cc) {
                  float a = __fmul_rn(t[0][cc], hp[k].w00);
                  a = fmaf(t[1][cc], hp[k].w01, a);
                  a = fmaf(t[2][cc], hp[k].w10, a);
                  st[hp[k].soff + cc * (F2_H * F2_P)] = fmaf(t[3][cc], hp[k].w11, a);
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full(s));
      }
      gchunk += nchunks;
    }
  } else {
    corr_compute<CORR_STAGES, true>(NoTileHook(), smem, full(0), empty(0), f1, f1_bs, f2, f2_bs, out, out_bs, B, C, H, W, P, shift, slope,
                                    vec_ok, tiles_x, tiles_y, ntiles);
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-fed kernel (the default wherever rows are 16-byte aligned): same tiles, same compute warps, but the staging is
// done by the copy engine.  One elected thread issues, per 8-channel chunk, a [8 ch][8 rows][36] box of f1 and — plain —
// a [8][16][44] box of f2 straight into the ring slot (box origin (x0-4, y0-4): out-of-image rows / columns / channels
// arrive as zeros, which IS the correlation's zero padding and the ragged channel tail; the extra 4 columns make the
// row pitch bank-conflict-free).  No thread computes an address or a predicate per element any more.
// Fused with the warp: the f2 box is the tile's 24 x 48 SOURCE FOOTPRINT, placed from the min/max of the sample
// coordinates, delivered into a second ring; six SAMPLER warps turn footprint -> warped halo tile.  The taps / weights
// / hard mask of a tile's 640 halo positions (bit-exact recipe, common.cuh) are computed ONE TILE AHEAD by the compute
// warps — at the one point where their 72 accumulators are dead and they would otherwise wait for the samplers — into
// a double-buffered shared-memory table; the samplers and the copy issuer only read it (ncu: with the set-up on the
// sampler warps it was 48 % of their time and the compute warps waited 43 % of theirs).
constexpr int T_NS_PLAIN = 6, T_NS_FUSED = 3, T_NFS = IRR_CORR_NFS;
constexpr int T_PLAIN_THREADS = NCOMP + 32;   // plain: 9 compute warps + the issuer's warp (no register cap at 128)
constexpr int T_NSAMP = 192;                  // sampler threads (warps 9-11, 13-15)
// Warp w issues on scheduler w % 4: compute warps 0-8 load the schedulers 3:2:2:2, so the six sampler warps go to
// schedulers 1-3 and the (single-thread) copy issuer's warp is the one that shares scheduler 0.
constexpr int T_ISSUER_FUSED = 12 * 32;
constexpr int T_KPOS = (NHALO + T_NSAMP - 1) / T_NSAMP;  // 4 halo positions per sampler thread
constexpr int T_KSET = (NHALO + NCOMP - 1) / NCOMP;      // 3 halo positions per compute thread in the set-up
constexpr uint32_t T_F1_BYTES = F1_ELEMS * 4, T_F2_BYTES = F2_ELEMS * 4, T_FP_BYTES = FP_ELEMS * 4;
constexpr int T_TAB_WORDS = NHALO * 5;        // per slot: float4 weights[640], int off_dxy[640]
constexpr int CORR_SMEM_TMA_PLAIN = T_NS_PLAIN * STAGE_ELEMS * 4 + 256;
constexpr int CORR_SMEM_TMA_FUSED = T_NS_FUSED * STAGE_ELEMS * 4 + T_NFS * FP_ELEMS * 4 + 2 * T_TAB_WORDS * 4 + 256;

// Tap table set-up for one tile (runs on the 288 compute threads, one tile ahead).
struct TapSetup {
  const float* flow; long long flow_bs;
  GridArgs g;
  int H, W, P, tiles_x, tiles_y;
  int Pi, xalign;  // row pitch of f2 (elements of ITS dtype); the footprint origin is aligned down to xalign + 1 elements (16 bytes)
  float* tab;      // [2][T_TAB_WORDS]
  int* meta;       // [2][4] {oy, ox, foot, -}
  int* red;        // [3][4] {ymin, ymax, xmin, xmax}
  uint32_t tabfull0;
  __device__ __forceinline__ void setup(int tile, int j) const {
    const int tid = threadIdx.x, lane = tid & 31;
    const int HW = H * P;
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int b = tile / (tiles_x * tiles_y);
    const int y0 = ty * TH, x0 = tx * TW;
    const int slot = j & 1;
    int* rd = red + 4 * (j % 3);
    if (tid == 0) {  // the slot of the NEXT reduction; nobody touches it before the barrier below
      int* rn = red + 4 * ((j + 1) % 3);
      rn[0] = 1 << 30; rn[1] = -1; rn[2] = 1 << 30; rn[3] = -1;
    }
    float fu[T_KSET], fv[T_KSET];
    bool in_[T_KSET];
#pragma unroll
    for (int k = 0; k < T_KSET; ++k) {  // all flow loads in flight before any arithmetic
      const int h = tid + k * NCOMP;
      const int hr = h / F2_WV, hx = h - hr * F2_WV;
      const int gy = y0 - MD + hr, gx = x0 - MD + hx;
      in_[k] = h < NHALO && gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* fl = flow + (size_t)b * flow_bs + (in_[k] ? (size_t)gy * P + gx : 0);
      fu[k] = __ldg(fl); fv[k] = __ldg(fl + HW);
    }
    float4 wq[T_KSET];
    int ya_[T_KSET], xa_[T_KSET], dxy[T_KSET];
    int lo_y = 1 << 30, hi_y = -1, lo_x = 1 << 30, hi_x = -1;
#pragma unroll
    for (int k = 0; k < T_KSET; ++k) {
      const int h = tid + k * NCOMP;
      wq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      ya_[k] = 0; xa_[k] = 0; dxy[k] = -1;  // dead
      if (in_[k]) {
        const int hr = h / F2_WV, hx = h - hr * F2_WV;
        const int gy = y0 - MD + hr, gx = x0 - MD + hx;
        float ix, iy;
        sample_coords(g, fu[k], fv[k], gx, gy, W, H, ix, iy);
        Taps tp = make_taps(ix, iy, W, H);
        if (tp.mask != 0.f && (tp.w00 != 0.f || tp.w01 != 0.f || tp.w10 != 0.f || tp.w11 != 0.f)) {
          const int xa = min(max(tp.x0, 0), W - 1), xb = min(max(tp.x0 + 1, 0), W - 1);
          const int ya = min(max(tp.y0, 0), H - 1), yb = min(max(tp.y0 + 1, 0), H - 1);
          ya_[k] = ya; xa_[k] = xa;
          // a clamped (out-of-range) tap has zero weight (make_taps): aliasing it onto its in-range neighbour's
          // address is harmless, the four reads are always in bounds.
          dxy[k] = ((xb != xa) ? 1 : 0) | ((yb != ya) ? 2 : 0);
          wq[k] = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
          lo_y = min(lo_y, ya); hi_y = max(hi_y, yb); lo_x = min(lo_x, xa); hi_x = max(hi_x, xb);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
      lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
    }
    if (lane == 0 && hi_y >= 0) { atomicMin(&rd[0], lo_y); atomicMax(&rd[1], hi_y); atomicMin(&rd[2], lo_x); atomicMax(&rd[3], hi_x); }
    asm volatile("bar.sync 2, %0;" ::"n"(NCOMP) : "memory");
    const int ymin = rd[0], ymax = rd[1], xmin = rd[2], xmax = rd[3];
    const bool any_live = ymax >= ymin;
    // TMA traps on an inner coordinate that is not 16-byte aligned (measured): align the window origin down
    const int oy = any_live ? ymin : 0, ox = any_live ? (xmin & ~xalign) : 0;
    const int foot = (!any_live || ((ymax - oy) < FP_H && (xmax - ox) < FP_W)) ? 1 : 0;
    float4* tw = reinterpret_cast<float4*>(tab + slot * T_TAB_WORDS);
    int* to = reinterpret_cast<int*>(tab + slot * T_TAB_WORDS + NHALO * 4);
#pragma unroll
    for (int k = 0; k < T_KSET; ++k) {
      const int h = tid + k * NCOMP;
      if (h < NHALO) {
        tw[h] = wq[k];
        const int off = foot ? (ya_[k] - oy) * FP_W + (xa_[k] - ox) : ya_[k] * Pi + xa_[k];
        to[h] = dxy[k] < 0 ? -1 : ((off << 2) | dxy[k]);
      }
    }
    if (tid == 0) {
      int* mt = meta + 4 * slot;
      mt[0] = oy; mt[1] = ox; mt[2] = foot;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(tabfull0 + 8u * slot);  // release: readers acquire the table and meta through the wait
  }
};

// Tap table pre-pass (PRETAB): one thread per pixel computes what TapSetup::setup computes per halo position — the
// bilinear weights with the hard mask and the bounds folded in (bit-exact recipe, common.cuh) and the clamped integer
// corner — once per PIXEL instead of once per tile-halo position (2.5x), and OFF the compute warps of the correlation
// kernel, where it was 19 % of their time (role counters, profiles/r01_corr_role_counters.txt).  Entry: float4 weights +
// one int  (ya << 16) | (xa << 2) | dxy  (dxy bit 0 / 1: the second column / row is a distinct in-range texel), -1 = the
// pixel samples nothing (masked out or fully out of bounds).  Needs H < 32768 and W < 16384.
__global__ void __launch_bounds__(256) corr_taps_kernel(const float* __restrict__ flow, long long flow_bs, GridArgs g, int H,
                                                        int W, int P, float4* __restrict__ tabW, int* __restrict__ tabI) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int b = blockIdx.y;
  const int gy = pix / W, gx = pix - gy * W;
  const float* fl = flow + (size_t)b * flow_bs + (size_t)gy * P + gx;
  float ix, iy;
  sample_coords(g, __ldg(fl), __ldg(fl + (size_t)H * P), gx, gy, W, H, ix, iy);
  const Taps tp = make_taps(ix, iy, W, H);
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
  int e = -1;
  if (tp.mask != 0.f && (tp.w00 != 0.f || tp.w01 != 0.f || tp.w10 != 0.f || tp.w11 != 0.f)) {
    const int xa = min(max(tp.x0, 0), W - 1), xb = min(max(tp.x0 + 1, 0), W - 1);
    const int ya = min(max(tp.y0, 0), H - 1), yb = min(max(tp.y0 + 1, 0), H - 1);
    e = (ya << 16) | (xa << 2) | ((xb != xa) ? 1 : 0) | ((yb != ya) ? 2 : 0);
    w = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
  }
  tabW[(size_t)b * HW + pix] = w;
  tabI[(size_t)b * HW + pix] = e;
}

// BF: f1 / f2 are bf16 storage (their batch strides and the pitch Pi in bf16 elements); flow / out stay fp32 with pitch P.
template <bool FUSED, bool SPLIT, bool PRETAB, bool BF = false>
__global__ void __launch_bounds__(FUSED ? CORR_THREADS : T_PLAIN_THREADS, 1)
    corr_tma_kernel(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
                    const float* __restrict__ f1, long long f1_bs, const float* __restrict__ f2, long long f2_bs,
                    const float* __restrict__ flow, long long flow_bs, float* __restrict__ out, long long out_bs,
                    GridArgs g, int B, int C, int H, int W, int P, int Pi, int shift, float slope, int vec_ok, int tiles_x,
                    int tiles_y, int ntiles, int ctr, CSplit sp, const float4* __restrict__ tabW,
                    const int* __restrict__ tabI) {
  static_assert(!(BF && PRETAB), "the pre-pass variant is fp32 only");
  constexpr uint32_t F1_BYTES = BF ? CC * TH * F1_P16 * 2 : T_F1_BYTES;
  constexpr uint32_t F2_BYTES = BF ? CC * F2_H * F2_P16 * 2 : T_F2_BYTES;
  constexpr uint32_t FP_BYTES = BF ? T_FP_BYTES / 2 : T_FP_BYTES;
  constexpr int XORG = BF ? 2 * MD : MD;   // plain f2 box origin x0 - XORG: a 16-byte aligned inner coordinate
  constexpr int NS = FUSED ? T_NS_FUSED : T_NS_PLAIN;
  const int ksplit = SPLIT ? sp.ksplit : 1;
  const int nvt = ntiles * ksplit;
  extern __shared__ __align__(1024) float smem[];
  float* fpr = smem + NS * STAGE_ELEMS;                          // footprint ring (FUSED)
  float* tab = fpr + (FUSED ? T_NFS * FP_ELEMS : 0);             // tap tables (FUSED)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + (FUSED ? 2 * T_TAB_WORDS : 0));
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (NS + s); };
  auto fpfull = [&](int s) { return bar0 + 8u * (2 * NS + s); };
  auto fpempty = [&](int s) { return bar0 + 8u * (2 * NS + T_NFS + s); };
  auto tabfull = [&](int s) { return bar0 + 8u * (2 * NS + 2 * T_NFS + s); };
  auto tabempty = [&](int s) { return bar0 + 8u * (2 * NS + 2 * T_NFS + 2 + s); };   // PRETAB: samplers have read meta[s]
  int* meta = reinterpret_cast<int*>(bars + 2 * NS + 2 * T_NFS + 4);  // 2 x {oy, ox, foot, -}, then red[3][4]
  int* red = meta + 8;

  const int tid = threadIdx.x;
  const int HWi = H * Pi;  // channel stride of f1 / f2 (their own pitch and element type)
  const int HWd = H * W;   // dense: the pre-pass' tap table
  const int nchunks = (C + CC - 1) / CC;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full(s), 1 + (FUSED ? T_NSAMP / 32 : 0));  // the copy's expect_tx arrival (+ one per sampler warp)
      mbar_init(empty(s), NCOMP / 32);
    }
    for (int s = 0; s < T_NFS; ++s) {
      mbar_init(fpfull(s), 1);
      mbar_init(fpempty(s), T_NSAMP / 32);
    }
    mbar_init(tabfull(0), PRETAB ? 1 : NCOMP / 32);
    mbar_init(tabfull(1), PRETAB ? 1 : NCOMP / 32);
    mbar_init(tabempty(0), T_NSAMP / 32);
    mbar_init(tabempty(1), T_NSAMP / 32);
    for (int i = 0; i < 3; ++i) { red[4 * i] = 1 << 30; red[4 * i + 1] = -1; red[4 * i + 2] = 1 << 30; red[4 * i + 3] = -1; }
    mbar_fence_init();
  }
  __syncthreads();

  if (tid < NCOMP) {
    if (FUSED && !PRETAB) {
      TapSetup hook;
      hook.flow = flow; hook.flow_bs = flow_bs; hook.g = g; hook.H = H; hook.W = W; hook.P = P; hook.tiles_x = tiles_x;
      hook.tiles_y = tiles_y; hook.tab = tab; hook.meta = meta; hook.red = red; hook.tabfull0 = tabfull(0);
      hook.Pi = Pi; hook.xalign = BF ? 7 : 3;
      corr_compute<NS, false, TapSetup, SPLIT, BF, false>(hook, smem, full(0), empty(0), f1, f1_bs, f2, f2_bs, out, out_bs, B, C,
                                                          H, W, P, shift, slope, vec_ok, tiles_x, tiles_y, ntiles, ctr, sp);
    } else {
      corr_compute<NS, false, NoTileHook, SPLIT, BF, BF && !FUSED>(NoTileHook(), smem, full(0), empty(0), f1, f1_bs, f2, f2_bs,
                                                                   out, out_bs, B, C, H, W, P, shift, slope, vec_ok, tiles_x,
                                                                   tiles_y, ntiles, ctr, sp);
    }
  } else if (FUSED && PRETAB && (tid >> 5) == (T_ISSUER_FUSED >> 5)) {
    // ============================== COPY ISSUER, table from the pre-pass (one warp) ==============================
    // The whole warp finds a tile's source footprint from the per-pixel tap table (20 entries per lane), ONE TILE AHEAD
    // (the loads of tile t+1 are in flight while lane 0 issues the copies of tile t); lane 0 publishes {oy, ox, foot}
    // for the samplers and issues the copies.
    const int lane = tid & 31;
    int gchunk = 0, it = 0;
    // footprint of virtual tile vt -> (oy, ox, foot), identical in every lane
    auto footprint = [&](int vt, int& oy, int& ox, int& foot) {
      const int tile = SPLIT ? vt / ksplit : vt;
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const int y0 = ty * TH, x0 = tx * TW;
      int lo_y = 1 << 30, hi_y = -1, lo_x = 1 << 30, hi_x = -1;
      const int* tI = tabI + (size_t)b * HWd;
      int e[(NHALO + 31) / 32];
#pragma unroll
      for (int k = 0; k < (NHALO + 31) / 32; ++k) {   // all loads in flight first
        const int h = lane + 32 * k;
        const int hr = h / F2_WV, hx = h - hr * F2_WV;
        const int gy = y0 - MD + hr, gx = x0 - MD + hx;
        e[k] = (h < NHALO && gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(tI + (size_t)gy * W + gx) : -1;
      }
#pragma unroll
      for (int k = 0; k < (NHALO + 31) / 32; ++k) {
        if (e[k] >= 0) {
          const int ya = e[k] >> 16, xa = (e[k] >> 2) & 0x3fff;
          lo_y = min(lo_y, ya); hi_y = max(hi_y, ya + ((e[k] >> 1) & 1));
          lo_x = min(lo_x, xa); hi_x = max(hi_x, xa + (e[k] & 1));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
        lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
      }
      const bool any_live = hi_y >= lo_y;
      oy = any_live ? lo_y : 0;
      ox = any_live ? (lo_x & ~3) : 0;   // TMA: 16-byte aligned inner coordinate
      foot = (!any_live || ((hi_y - oy) < FP_H && (hi_x - ox) < FP_W)) ? 1 : 0;
    };
    auto publish = [&](int j, int oy, int ox, int foot) {   // lane 0: meta of the j-th tile of this CTA
      const int slot = j & 1;
      if (j >= 2) mbar_wait(tabempty(slot), (uint32_t)(((j >> 1) - 1) & 1));   // samplers are done with this slot
      volatile int* mt = meta + 4 * slot;
      mt[0] = oy; mt[1] = ox; mt[2] = foot;
      mbar_arrive(tabfull(slot));   // release: the samplers acquire meta through their wait
    };
    int oy = 0, ox = 0, foot = 1;
    if ((int)blockIdx.x < nvt) {
      footprint(blockIdx.x, oy, ox, foot);
      if (lane == 0) publish(0, oy, ox, foot);
    }
    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x, ++it) {
      const int tile = SPLIT ? vt / ksplit : vt, ks = SPLIT ? vt - tile * ksplit : 0;
      const int c_lo = SPLIT ? ks * sp.cps : 0, c_hi = SPLIT ? min(nchunks, c_lo + sp.cps) : nchunks;
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const int y0 = ty * TH, x0 = tx * TW;
      int b2 = b + shift;
      if (b2 >= B) b2 -= B;
      int noy = 0, nox = 0, nfoot = 1;
      const bool has_next = vt + (int)gridDim.x < nvt;
      if (has_next) footprint(vt + (int)gridDim.x, noy, nox, nfoot);
      if (lane == 0) {
        if (has_next) publish(it + 1, noy, nox, nfoot);
        for (int ci = c_lo; ci < c_hi; ++ci, ++gchunk) {
          const int s = gchunk % NS;
          mbar_wait(empty(s), (uint32_t)(((gchunk / NS) & 1) ^ 1));
          const uint32_t st = smem_u32(smem + s * STAGE_ELEMS);
          mbar_expect_tx(full(s), T_F1_BYTES);
          tma_load_4d(st, &m1, x0, y0, ci * CC, b, full(s));
          const int fs = gchunk % T_NFS;
          mbar_wait(fpempty(fs), (uint32_t)(((gchunk / T_NFS) & 1) ^ 1));
          if (foot) {
            mbar_expect_tx(fpfull(fs), T_FP_BYTES);
            tma_load_4d(smem_u32(fpr + fs * FP_ELEMS), &m2, ox, oy, ci * CC, b2, fpfull(fs));
          } else {
            mbar_arrive(fpfull(fs));  // divergent tile: the samplers gather from global memory, keep the phases moving
          }
        }
      }
      __syncwarp();
      oy = noy; ox = nox; foot = nfoot;
    }
  } else if (!(FUSED && PRETAB) && tid == (FUSED ? T_ISSUER_FUSED : NCOMP)) {
    // ============================== COPY ISSUER (one thread) ==============================
    int gchunk = 0, it = 0;
    const bool con = ctr && blockIdx.x == 0;
    const long long ct0 = con ? clock64() : 0;
    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x, ++it) {
      const int tile = SPLIT ? vt / ksplit : vt, ks = SPLIT ? vt - tile * ksplit : 0;
      const int c_lo = SPLIT ? ks * sp.cps : 0, c_hi = SPLIT ? min(nchunks, c_lo + sp.cps) : nchunks;
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const int y0 = ty * TH, x0 = tx * TW;
      int b2 = b + shift;
      if (b2 >= B) b2 -= B;
      int oy = 0, ox = 0, foot = 0;
      if (FUSED) {
        mbar_wait_ctr(tabfull(it & 1), (uint32_t)((it >> 1) & 1), con, 8);
        const volatile int* mt = meta + 4 * (it & 1);
        oy = mt[0]; ox = mt[1]; foot = mt[2];
      }
      for (int ci = c_lo; ci < c_hi; ++ci, ++gchunk) {
        const int s = gchunk % NS;
        mbar_wait_ctr(empty(s), (uint32_t)(((gchunk / NS) & 1) ^ 1), con, 9);
        const uint32_t st = smem_u32(smem + s * STAGE_ELEMS);
        if (!FUSED) {
          mbar_expect_tx(full(s), F1_BYTES + F2_BYTES);
          tma_load_4d(st, &m1, x0, y0, ci * CC, b, full(s));
          tma_load_4d(st + T_F1_BYTES, &m2, x0 - XORG, y0 - MD, ci * CC, b2, full(s));   // the f2 REGION keeps its fp32 offset
        } else {
          mbar_expect_tx(full(s), F1_BYTES);
          tma_load_4d(st, &m1, x0, y0, ci * CC, b, full(s));
          const int fs = gchunk % T_NFS;
          mbar_wait_ctr(fpempty(fs), (uint32_t)(((gchunk / T_NFS) & 1) ^ 1), con, 10);
          if (foot) {
            mbar_expect_tx(fpfull(fs), FP_BYTES);
            tma_load_4d(smem_u32(fpr + fs * FP_ELEMS), &m2, ox, oy, ci * CC, b2, fpfull(fs));
          } else {
            mbar_arrive(fpfull(fs));  // divergent tile: the samplers gather from global memory, keep the phases moving
          }
        }
      }
    }
    if (con) corr_ctr[11] += (unsigned long long)(clock64() - ct0);
  } else if (FUSED && tid >= NCOMP && (tid >> 5) != (T_ISSUER_FUSED >> 5)) {
    // ============================== SAMPLERS (warps 9-11, 13-15) ==============================
    const int pt = tid - NCOMP - ((tid >> 5) > (T_ISSUER_FUSED >> 5) ? 32 : 0);
    const int lane = tid & 31;
    int gchunk = 0, it = 0;
    const bool con = ctr && blockIdx.x == 0 && pt == 0;
    const long long ct0 = con ? clock64() : 0;
    // PRETAB: this thread's raw table entries of the NEXT tile (loaded while the current tile is sampled)
    int pre_e[T_KPOS];
    float4 pre_w[T_KPOS];
    auto load_entries = [&](int vtile) {
      const int tl = SPLIT ? vtile / ksplit : vtile;
      const int tx = tl % tiles_x;
      const int ty = (tl / tiles_x) % tiles_y;
      const int tb = tl / (tiles_x * tiles_y);
      const int y0 = ty * TH, x0 = tx * TW;
      const float4* tW = tabW + (size_t)tb * HWd;
      const int* tI = tabI + (size_t)tb * HWd;
#pragma unroll
      for (int k = 0; k < T_KPOS; ++k) {
        const int h = pt + k * T_NSAMP;
        const int hr = h / F2_WV, hx = h - hr * F2_WV;
        const int gy = y0 - MD + hr, gx = x0 - MD + hx;
        const bool in = h < NHALO && gy >= 0 && gy < H && gx >= 0 && gx < W;
        pre_e[k] = in ? __ldg(tI + (size_t)gy * W + gx) : -1;
        pre_w[k] = in ? __ldg(tW + (size_t)gy * W + gx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (PRETAB && (int)blockIdx.x < nvt) load_entries(blockIdx.x);
    for (int vt = blockIdx.x; vt < nvt; vt += gridDim.x, ++it) {
      const int tile = SPLIT ? vt / ksplit : vt, ks = SPLIT ? vt - tile * ksplit : 0;
      const int c_lo = SPLIT ? ks * sp.cps : 0, c_hi = SPLIT ? min(nchunks, c_lo + sp.cps) : nchunks;
      const int b = tile / (tiles_x * tiles_y);
      int b2 = b + shift;
      if (b2 >= B) b2 -= B;
      // f2 of this tile's image (element offsets in f2's own type)
      const void* f2b = BF ? static_cast<const void*>(reinterpret_cast<const unsigned short*>(f2) + (size_t)b2 * f2_bs)
                           : static_cast<const void*>(f2 + (size_t)b2 * f2_bs);
      // this tile's taps, from the table the compute warps filled one tile ago (PRETAB: from the pre-pass' table)
      const int slot = it & 1;
      mbar_wait_ctr(tabfull(slot), (uint32_t)((it >> 1) & 1), con, 16);
      const int foot = reinterpret_cast<const volatile int*>(meta)[4 * slot + 2];
      float4 wq[T_KPOS];
      int od[T_KPOS];
      if (PRETAB) {
        const int oy = reinterpret_cast<const volatile int*>(meta)[4 * slot], ox = reinterpret_cast<const volatile int*>(meta)[4 * slot + 1];
        __syncwarp();
        if (lane == 0) mbar_arrive(tabempty(slot));   // meta[slot] may be overwritten for tile it + 2
#pragma unroll
        for (int k = 0; k < T_KPOS; ++k) {   // entries were loaded one tile ago (pre_e / pre_w)
          wq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          od[k] = -1;
          if (pre_e[k] >= 0) {
            const int ya = pre_e[k] >> 16, xa = (pre_e[k] >> 2) & 0x3fff;
            const int off = foot ? (ya - oy) * FP_W + (xa - ox) : ya * Pi + xa;
            od[k] = (off << 2) | (pre_e[k] & 3);
            wq[k] = pre_w[k];
          }
        }
        if (vt + (int)gridDim.x < nvt) load_entries(vt + (int)gridDim.x);   // next tile's entries, in flight during this tile
      } else {
        const float4* tw = reinterpret_cast<const float4*>(tab + slot * T_TAB_WORDS);
        const int* to = reinterpret_cast<const int*>(tab + slot * T_TAB_WORDS + NHALO * 4);
#pragma unroll
        for (int k = 0; k < T_KPOS; ++k) {
          const int h = pt + k * T_NSAMP;
          wq[k] = h < NHALO ? tw[h] : make_float4(0.f, 0.f, 0.f, 0.f);
          od[k] = h < NHALO ? to[h] : -1;
        }
      }
      for (int ci = c_lo; ci < c_hi; ++ci, ++gchunk) {
        const int s = gchunk % NS, fs = gchunk % T_NFS;
        mbar_wait_ctr(fpfull(fs), (uint32_t)((gchunk / T_NFS) & 1), con, 17);
        mbar_wait_ctr(empty(s), (uint32_t)(((gchunk / NS) & 1) ^ 1), con, 18);
        float* st = smem + s * STAGE_ELEMS + F1_ELEMS;
        const int c0 = ci * CC;
        if (foot) {
          const void* fp = fpr + fs * FP_ELEMS;   // bf16 footprints use the first half of the slot
#pragma unroll
          for (int k = 0; k < T_KPOS; ++k) {
            const int h = pt + k * T_NSAMP;
            if (h < NHALO) {
              const int hr = h / F2_WV, hx = h - hr * F2_WV;
              float* dst = st + hr * F2_P + hx;
              const bool live = od[k] >= 0;
              const int pq = live ? (od[k] >> 2) : 0;
              const int dx = od[k] & 1, dy = (od[k] & 2) ? FP_W : 0;  // dead: od = -1 -> dx = 1, dy = FP_W, still inside
              // all 32 tap loads of the position first, then the arithmetic, then the stores: a shared-memory store between
              // the channels would serialise them (the compiler cannot prove dst and the footprint do not alias)
              float t[4][CC];
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) {
                const int pc = pq + cc * (FP_H * FP_W);
                t[0][cc] = ld_in<BF>(fp, pc); t[1][cc] = ld_in<BF>(fp, pc + dx);
                t[2][cc] = ld_in<BF>(fp, pc + dy); t[3][cc] = ld_in<BF>(fp, pc + dy + dx);
              }
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) {
                float a = __fmul_rn(t[0][cc], wq[k].x);  // tap order of grid_sampler_2d
                a = fmaf(t[1][cc], wq[k].y, a);
                a = fmaf(t[2][cc], wq[k].z, a);
                a = fmaf(t[3][cc], wq[k].w, a);
                t[0][cc] = live ? a : 0.f;
              }
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) dst[cc * (F2_H * F2_P)] = t[0][cc];
            }
          }
        } else {  // divergent flow: gather the taps from global memory
#pragma unroll
          for (int k = 0; k < T_KPOS; ++k) {
            const int h = pt + k * T_NSAMP;
            if (h < NHALO) {
              const int hr = h / F2_WV, hx = h - hr * F2_WV;
              float* dst = st + hr * F2_P + hx;
              const bool live = od[k] >= 0;
              const long long pq = (long long)c0 * HWi + (live ? (od[k] >> 2) : 0);
              const int dx = live ? (od[k] & 1) : 0, dy = (live && (od[k] & 2)) ? Pi : 0;
              float t[4][CC];
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) {
                const bool okc = live && (c0 + cc < C);
                const long long pc = pq + (long long)cc * HWi;
                t[0][cc] = okc ? ldg_in<BF>(f2b, pc) : 0.f;
                t[1][cc] = okc ? ldg_in<BF>(f2b, pc + dx) : 0.f;
                t[2][cc] = okc ? ldg_in<BF>(f2b, pc + dy) : 0.f;
                t[3][cc] = okc ? ldg_in<BF>(f2b, pc + dy + dx) : 0.f;
              }
#pragma unroll
              for (int cc = 0; cc < CC; ++cc) {
                float a = __fmul_rn(t[0][cc], wq[k].x);
                a = fmaf(t[1][cc], wq[k].y, a);
                a = fmaf(t[2][cc], wq[k].z, a);
                dst[cc * (F2_H * F2_P)] = fmaf(t[3][cc], wq[k].w, a);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(full(s));
          mbar_arrive(fpempty(fs));
        }
      }
    }
    if (con) corr_ctr[19] += (unsigned long long)(clock64() - ct0);
  }
}

// Generic correlation (any odd kernel_size, strides): one thread per output element.  Restates
// correlation_cuda_kernel.cu:41-114 directly on NCHW with on-the-fly zero padding; API completeness only.
__global__ void corr_generic_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                                    float* __restrict__ out, int C, int H, int W, int pad, int krad, int md, int s1,
                                    int s2, int drad, int oc, int oh, int ow, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int bx = (int)(i % ow);
  long long rr = i / ow;
  int by = (int)(rr % oh);
  rr /= oh;
  int tc = (int)(rr % oc);
  int n = (int)(rr / oc);
  int dsize = 2 * drad + 1;
  int tj = tc / dsize - drad, ti = tc % dsize - drad;
  int y1 = by * s1 + md - pad, x1 = bx * s1 + md - pad;  // un-padded coordinates
  int y2 = y1 + tj * s2, x2 = x1 + ti * s2;
  const float* a = in1 + (size_t)n * C * H * W;
  const float* bq = in2 + (size_t)n * C * H * W;
  float acc = 0.f;
  for (int j = -krad; j <= krad; ++j)
    for (int ii = -krad; ii <= krad; ++ii) {
      int ya = y1 + j, xa = x1 + ii, yb = y2 + j, xb = x2 + ii;
      if (ya < 0 || ya >= H || xa < 0 || xa >= W || yb < 0 || yb >= H || xb < 0 || xb >= W) continue;
      for (int c = 0; c < C; ++c)
        acc = fmaf(__ldg(a + ((size_t)c * H + ya) * W + xa), __ldg(bq + ((size_t)c * H + yb) * W + xb), acc);
    }
  int ks = 2 * krad + 1;
  out[i] = acc / (float)(ks * ks * C);
}

// out = act((1/C) * sum_ks ws[ks]) — the channel-split launches' second pass (fixed summation order: deterministic).
__global__ void corr_split_finish(const float* __restrict__ ws, long long ws_stride, int ksplit, float* __restrict__ out,
                                  long long out_bs, int HW, long long total, float inv_c, float slope) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float acc = 0.f;
  for (int k = 0; k < ksplit; ++k) acc += __ldg(ws + (size_t)k * ws_stride + i);
  const long long per = (long long)(ND * ND) * HW;
  const long long b = i / per;
  out[(size_t)b * out_bs + (size_t)(i - b * per)] = leaky(acc * inv_c, slope);
}

// Channel-split plan: only when the launch has at most half as many tiles as the GPU has SMs and >= 2 chunks.
static int corr_ksplit(int ntiles, int nchunks, int* cps_out) {
  int ksplit = 1, cps = nchunks;
  const int sms = sm_count();
  if (nchunks >= 2 && ntiles * 2 <= sms) {
    int want = sms / ntiles;
    if (want > nchunks) want = nchunks;
    cps = (nchunks + want - 1) / want;
    ksplit = (nchunks + cps - 1) / cps;
  }
  *cps_out = cps;
  return ksplit;
}
static bool corr_no_split() {  // IRR_CORR_NO_SPLIT=1: never split channels (A/B measurements)
  const char* e = getenv("IRR_CORR_NO_SPLIT");
  return e && e[0] == '1';
}

// Workspace layout: [tap table: B*H*W float4, B*H*W int (fused launches; 256-byte aligned sections)] [split slices].
static size_t corr_tab_bytes(int B, int H, int W) {
  const size_t n = (size_t)B * H * W;
  return ((n * 16 + 255) / 256) * 256 + ((n * 4 + 255) / 256) * 256;
}
// The pre-pass variant is OFF by default: measured on the B200 (profiles/r02_bench_corr_variants.txt) it is SLOWER than
// the in-kernel set-up (level 4 fused: 188 us vs 143 us) — the compute warps are not the critical path of the fused
// kernel, the sampler / copy chain is, and the table look-ups lengthen it.  IRR_CORR_PRETAB=1 enables it (A/B runs, tests).
static bool corr_no_pretab() {
  const char* e = getenv("IRR_CORR_PRETAB");
  return !(e && e[0] == '1');
}

size_t corr_workspace_bytes(int B, int C, int H, int W, int fused) {
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const long long nt = (long long)tiles_x * tiles_y * B;
  if (nt > 0x7fffffffLL) return 0;
  int cps;
  const int k = corr_ksplit((int)nt, (C + CC - 1) / CC, &cps);
  size_t n = k > 1 ? (size_t)k * B * (ND * ND) * H * ((W + 3) & ~3) * sizeof(float) : 0;
  if (fused && !corr_no_pretab() && H < 32768 && W < 16384) n += corr_tab_bytes(B, H, W);
  return n;
}

static bool corr_no_tma() {  // IRR_CORR_NO_TMA=1: force the cp.async kernel (A/B measurements, tests of the fallback)
  const char* e = getenv("IRR_CORR_NO_TMA");
  return e && e[0] == '1';
}

static bool corr_ctr_on() {  // IRR_CORR_CTR=1: role cycle counters (debug)
  const char* e = getenv("IRR_CORR_CTR");
  return e && e[0] == '1';
}

template <bool FUSED>
static int launch_corr(const char* fn, const float* f1, long long f1_bs, const float* f2, long long f2_bs,
                       const float* flow, long long flow_bs, float* out, long long out_bs, const GridArgs& g, int B,
                       int C, int H, int W, int shift, float slope, void* ws, size_t ws_bytes, cudaStream_t st,
                       int pitch = 0, int dt = 0, int pitch_in = 0) {
  // dt = 1: f1 / f2 point to bf16 storage; f1_bs / f2_bs / pitch_in are in bf16 elements (pitch_in = 0: same as `pitch`)
  constexpr int CORR_SMEM = FUSED ? CORR_SMEM_FUSED : CORR_SMEM_PLAIN;
  static SmemAttrCache attr = {};
  if (int rc = ensure_dyn_smem(corr_kernel<FUSED>, CORR_SMEM, attr, fn)) return rc;
  // every H x W tensor of the call (f1, f2, flow, out) is stored with row pitch P >= W; P % 4 == 0 puts any width on the
  // TMA / vector paths (the tensor maps zero-fill columns >= W)
  const int P = pitch > 0 ? pitch : W;
  const int Pi = pitch_in > 0 ? pitch_in : P;
  if (P < W || Pi < W) return fail_arg(fn, "row pitch smaller than the width");
  if (dt != 0 && dt != 1) return fail_arg(fn, "dtype_in must be 0 (fp32) or 1 (bf16)");
  if (dt == 0 && Pi != P) return fail_arg(fn, "fp32 inputs share the row pitch of flow / out");
  int vec_ok = (P % 4 == 0) && (out_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec_ok && (P % 8 == 0) && (((long long)H * P) % 8 == 0) && (out_bs % 8 == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0))
    vec_ok = 2;  // every 8-pixel strip of every output plane is 32-byte aligned
  // 16-byte cp.async staging needs 16-byte aligned rows in both operands
  int vec_in = (P % 4 == 0) && (f1_bs % 4 == 0) && (f2_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(f1) & 15) == 0) &&
               ((reinterpret_cast<uintptr_t>(f2) & 15) == 0);
  int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  long long nt = (long long)tiles_x * tiles_y * B;
  if (nt > 0x7fffffffLL) return fail_arg(fn, "too many tiles");
  int ntiles = (int)nt;
  int grid = ntiles < sm_count() ? ntiles : sm_count();  // persistent: one CTA per SM
  if (dt == 1) vec_in = 1;   // bf16: the tensor maps check the alignment (pitch % 8, strides % 8, 16-byte bases)
  if (vec_in && (dt == 1 || !corr_no_tma())) {
    constexpr int TSMEM = FUSED ? CORR_SMEM_TMA_FUSED : CORR_SMEM_TMA_PLAIN;
    static SmemAttrCache tattr = {}, tattr_split = {}, tattr_pre = {}, tattr_pre_split = {}, tattr_bf = {}, tattr_bf_split = {};
    CUtensorMap m1, m2;
    const bool maps_ok =
        dt == 1 ? (make_nchw_map(&m1, f1, f1_bs, B, C, H, W, F1_P16, TH, CC, Pi, 2) &&
                   make_nchw_map(&m2, f2, f2_bs, B, C, H, W, FUSED ? FP_W : F2_P16, FUSED ? FP_H : F2_H, CC, Pi, 2))
                : (make_nchw_map(&m1, f1, f1_bs, B, C, H, W, F1_P, TH, CC, P) &&
                   make_nchw_map(&m2, f2, f2_bs, B, C, H, W, FUSED ? FP_W : F2_P, FUSED ? FP_H : F2_H, CC, P));
    if (maps_ok) {
      const int ctr = corr_ctr_on() ? 1 : 0;
      if (ctr) {
        static const unsigned long long zeros[32] = {0};
        cudaMemcpyToSymbolAsync(corr_ctr, zeros, sizeof(zeros), 0, cudaMemcpyHostToDevice, st);
      }
      CSplit sp = {1, 0, nullptr, 0};
      uint8_t* wsp = reinterpret_cast<uint8_t*>(ws);
      size_t ws_left = ws_bytes;
      float4* tabW = nullptr;
      int* tabI = nullptr;
      // tap table from a pre-pass (fused launches): the table goes first in the workspace
      if (FUSED && dt == 0 && wsp != nullptr && (reinterpret_cast<uintptr_t>(wsp) & 255) == 0 && !corr_no_pretab() && !ctr && H < 32768 &&
          W < 16384 && corr_tab_bytes(B, H, W) <= ws_left) {
        const size_t n = (size_t)B * H * W;
        tabW = reinterpret_cast<float4*>(wsp);
        tabI = reinterpret_cast<int*>(wsp + ((n * 16 + 255) / 256) * 256);
        wsp += corr_tab_bytes(B, H, W);
        ws_left -= corr_tab_bytes(B, H, W);
        dim3 pg((unsigned)((H * W + 255) / 256), (unsigned)B);
        corr_taps_kernel<<<pg, 256, 0, st>>>(flow, flow_bs, g, H, W, P, tabW, tabI);
        if (int rc = check_launch(fn)) return rc;
      }
      if (wsp != nullptr && !corr_no_split() && (reinterpret_cast<uintptr_t>(wsp) & 15) == 0) {
        int cps;
        const int k = corr_ksplit(ntiles, (C + CC - 1) / CC, &cps);
        const size_t slice = (size_t)B * (ND * ND) * H * P;   // the partial-sum slices are pitched like `out`
        if (k > 1 && (size_t)k * slice * sizeof(float) <= ws_left) {
          sp.ksplit = k; sp.cps = cps; sp.ws = reinterpret_cast<float*>(wsp); sp.ws_stride = (long long)slice;
          const long long nvt = (long long)ntiles * k;
          grid = nvt < sm_count() ? (int)nvt : sm_count();
        }
      }
      const int nthr = FUSED ? CORR_THREADS : T_PLAIN_THREADS;
#define IRR_CORR_LAUNCH(SPLIT_, PRE_, CACHE_, BF_)                                                                         \
  do {                                                                                                                     \
    if (int rc = ensure_dyn_smem(corr_tma_kernel<FUSED, SPLIT_, PRE_, BF_>, TSMEM, CACHE_, fn)) return rc;                  \
    corr_tma_kernel<FUSED, SPLIT_, PRE_, BF_><<<grid, nthr, TSMEM, st>>>(m1, m2, f1, f1_bs, f2, f2_bs, flow, flow_bs, out,  \
                                                                        out_bs, g, B, C, H, W, P, Pi, shift, slope, vec_ok, \
                                                                        tiles_x, tiles_y, ntiles, ctr, sp, tabW, tabI);     \
  } while (0)
      if (dt == 1) {
        if (sp.ksplit > 1) IRR_CORR_LAUNCH(true, false, tattr_bf_split, true);
        else IRR_CORR_LAUNCH(false, false, tattr_bf, true);
      } else if (FUSED && tabW != nullptr) {
        if (sp.ksplit > 1) IRR_CORR_LAUNCH(true, FUSED, tattr_pre_split, false);
        else IRR_CORR_LAUNCH(false, FUSED, tattr_pre, false);
      } else {
        if (sp.ksplit > 1) IRR_CORR_LAUNCH(true, false, tattr_split, false);
        else IRR_CORR_LAUNCH(false, false, tattr, false);
      }
#undef IRR_CORR_LAUNCH
      if (int rc = check_launch(fn)) return rc;
      if (sp.ksplit > 1) {
        const long long total = (long long)B * (ND * ND) * H * P;
        corr_split_finish<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sp.ws, sp.ws_stride, sp.ksplit, out, out_bs, H * P,
                                                                          total, 1.0f / (float)C, slope);
        return check_launch(fn);
      }
      return 0;
    }
  }
  if (dt == 1)
  {
    set_error("%s: bf16 inputs need 16-byte aligned rows (base %% 16 == 0, row pitch %% 8 == 0, batch stride %% 8 == 0)", fn);
    return IRR_E_ALIGN;
  }
  corr_kernel<FUSED><<<grid, CORR_THREADS, CORR_SMEM, st>>>(f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H,
                                                            W, P, shift, slope, vec_ok, vec_in, tiles_x, tiles_y, ntiles);
  return check_launch(fn);
}

// The fused launcher under a plain name, so the other translation unit (7-row tiles) can be called from this one.
int launch_corr_fused_variant(const char* fn, const float* f1, long long f1_bs, const float* f2, long long f2_bs,
                              const float* flow, long long flow_bs, float* out, long long out_bs, const GridArgs& g, int B,
                              int C, int H, int W, int shift, float slope, void* ws, size_t ws_bytes, cudaStream_t st,
                              int pitch) {
  return launch_corr<true>(fn, f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W, shift, slope, ws, ws_bytes, st,
                           pitch);
}

}  // namespace IRR_CORR_NS
}  // namespace irr

#ifndef IRR_CORR_VARIANT_ONLY
namespace irr {
namespace corr7 {   // correlation7.cu: 7-row tiles
int launch_corr_fused_variant(const char* fn, const float* f1, long long f1_bs, const float* f2, long long f2_bs,
                              const float* flow, long long flow_bs, float* out, long long out_bs, const GridArgs& g, int B,
                              int C, int H, int W, int shift, float slope, void* ws, size_t ws_bytes, cudaStream_t st,
                              int pitch);
size_t corr_workspace_bytes(int B, int C, int H, int W, int fused);
}  // namespace corr7
}  // namespace irr

using namespace irr;
using namespace irr::corr8;

// The 7-row-tile variant is OFF by default: measured on the B200 it is slower than 8-row tiles (level 4 fused: 171 us vs
// 143 us — 14 % more tiles and 7 % more halo per pixel outweigh the balanced schedulers).  IRR_CORR_TH7=1 selects it.
static bool corr_force_th8() {
  const char* e = getenv("IRR_CORR_TH7");
  return !(e && e[0] == '1');
}

extern "C" {

int irr_correlation_fwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, float* out,
                        long long out_bs, int B, int C, int H, int W, int max_disp, int f2_batch_shift,
                        float leaky_slope, irr_stream_t stream) {
  const char* fn = "irr_correlation_fwd";
  IRR_REQUIRE(f1 && f2 && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  IRR_REQUIRE(max_disp == MD, fn, "only max_disp == 4 is compiled (use irr_correlation_generic_fwd)");
  IRR_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < B, fn, "f2_batch_shift out of range");
  GridArgs g = make_grid_args(nullptr, nullptr, H, W, H, W, 1.f, 0);
  return launch_corr<false>(fn, f1, f1_bs, f2, f2_bs, nullptr, 0, out, out_bs, g, B, C, H, W, f2_batch_shift,
                            leaky_slope, nullptr, 0, as_stream(stream));
}

size_t irr_correlation_workspace_bytes(int B, int C, int H, int W, int fused) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  const size_t a = corr8::corr_workspace_bytes(B, C, H, W, fused);
  const size_t b = fused ? corr7::corr_workspace_bytes(B, C, H, W, fused) : 0;   // fused launches use 7-row tiles
  return a > b ? a : b;
}

int irr_warp_correlation_fwd_ws(const float* f1, long long f1_bs, const float* f2, long long f2_bs, const float* flow,
                                long long flow_bs, const float* lin_x, const float* lin_y, float* out, long long out_bs,
                                int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                                int f2_batch_shift, float leaky_slope, int grid_flags, void* workspace,
                                size_t workspace_bytes, int pitch, irr_stream_t stream) {
  const char* fn = "irr_warp_correlation_fwd_ws";
  IRR_REQUIRE(f1 && f2 && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  IRR_REQUIRE(max_disp == MD, fn, "only max_disp == 4 is compiled");
  IRR_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < B, fn, "f2_batch_shift out of range");
  if (flow == nullptr) {  // no warp: the plain cost volume
    GridArgs g = make_grid_args(nullptr, nullptr, H, W, H, W, 1.f, 0);
    return launch_corr<false>(fn, f1, f1_bs, f2, f2_bs, nullptr, 0, out, out_bs, g, B, C, H, W, f2_batch_shift, leaky_slope,
                              workspace, workspace_bytes, as_stream(stream), pitch);
  }
  IRR_REQUIRE(H_im > 0 && W_im > 0, fn, "non-positive image size");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  if (!corr_force_th8())   // experimental: 7-row tiles, 8 compute warps, two per scheduler (correlation7.cu)
    return corr7::launch_corr_fused_variant(fn, f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W,
                                            f2_batch_shift, leaky_slope, workspace, workspace_bytes, as_stream(stream), pitch);
  return launch_corr<true>(fn, f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W, f2_batch_shift, leaky_slope,
                           workspace, workspace_bytes, as_stream(stream), pitch);
}

int irr_warp_correlation_fwd_dt(const void* f1, long long f1_bs, const void* f2, long long f2_bs, int dtype_in, int in_pitch,
                                const float* flow, long long flow_bs, const float* lin_x, const float* lin_y, float* out,
                                long long out_bs, int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                                int f2_batch_shift, float leaky_slope, int grid_flags, void* workspace,
                                size_t workspace_bytes, int pitch, irr_stream_t stream) {
  const char* fn = "irr_warp_correlation_fwd_dt";
  IRR_REQUIRE(f1 && f2 && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, fn, "non-positive size");
  IRR_REQUIRE(max_disp == MD, fn, "only max_displacement = 4 (PWC parameters)");
  IRR_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < B, fn, "f2_batch_shift out of range");
  IRR_REQUIRE(dtype_in == IRR_DTYPE_F32 || dtype_in == IRR_DTYPE_BF16, fn, "dtype_in must be IRR_DTYPE_F32 or IRR_DTYPE_BF16");
  if (dtype_in == IRR_DTYPE_F32)
    return irr_warp_correlation_fwd_ws(static_cast<const float*>(f1), f1_bs, static_cast<const float*>(f2), f2_bs, flow, flow_bs,
                                       lin_x, lin_y, out, out_bs, B, C, H, W, H_im, W_im, div_flow, max_disp, f2_batch_shift,
                                       leaky_slope, grid_flags, workspace, workspace_bytes, pitch, stream);
  const float* a = static_cast<const float*>(f1);   // the bf16 kernels reinterpret the pointers
  const float* b = static_cast<const float*>(f2);
  if (flow == nullptr) {
    GridArgs g = make_grid_args(nullptr, nullptr, H, W, H, W, 1.f, 0);
    return launch_corr<false>(fn, a, f1_bs, b, f2_bs, nullptr, 0, out, out_bs, g, B, C, H, W, f2_batch_shift, leaky_slope,
                              workspace, workspace_bytes, as_stream(stream), pitch, 1, in_pitch);
  }
  IRR_REQUIRE(H_im > 0 && W_im > 0, fn, "non-positive image size");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  return launch_corr<true>(fn, a, f1_bs, b, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W, f2_batch_shift, leaky_slope,
                           workspace, workspace_bytes, as_stream(stream), pitch, 1, in_pitch);
}

int irr_warp_correlation_fwd(const float* f1, long long f1_bs, const float* f2, long long f2_bs, const float* flow,
                             long long flow_bs, const float* lin_x, const float* lin_y, float* out, long long out_bs,
                             int B, int C, int H, int W, int H_im, int W_im, float div_flow, int max_disp,
                             int f2_batch_shift, float leaky_slope, int grid_flags, irr_stream_t stream) {
  const char* fn = "irr_warp_correlation_fwd";
  IRR_REQUIRE(f1 && f2 && flow && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H_im > 0 && W_im > 0, fn, "non-positive size");
  IRR_REQUIRE(max_disp == MD, fn, "only max_disp == 4 is compiled");
  IRR_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < B, fn, "f2_batch_shift out of range");
  GridArgs g = make_grid_args(lin_x, lin_y, H, W, H_im, W_im, div_flow, grid_flags);
  return launch_corr<true>(fn, f1, f1_bs, f2, f2_bs, flow, flow_bs, out, out_bs, g, B, C, H, W, f2_batch_shift,
                           leaky_slope, nullptr, 0, as_stream(stream));
}

// Debug hook (not part of the ABI in include/irr_b200.h), scripts/corr_counters.py
int irrdbg_corr_counters(unsigned long long* out32) {
  if (!out32) return fail_arg("irrdbg_corr_counters", "null pointer");
  cudaError_t e = cudaMemcpyFromSymbol(out32, corr_ctr, 32 * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    set_error("irrdbg_corr_counters: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int irr_correlation_generic_out_shape(int H, int W, int pad_size, int kernel_size, int max_displacement, int stride1,
                                      int stride2, int* out_c, int* out_h, int* out_w) {
  const char* fn = "irr_correlation_generic_out_shape";
  IRR_REQUIRE(H > 0 && W > 0 && pad_size >= 0 && max_displacement >= 0 && stride1 >= 1 && stride2 >= 1, fn, "sizes");
  IRR_REQUIRE(kernel_size >= 1 && (kernel_size & 1), fn, "kernel_size must be odd");
  int krad = (kernel_size - 1) / 2, border = krad + max_displacement;  // correlation_cuda.cc:23-32
  int d = (max_displacement / stride2) * 2 + 1;
  int pH = H + 2 * pad_size, pW = W + 2 * pad_size;
  int oh = (pH - 2 * border + stride1 - 1) / stride1, ow = (pW - 2 * border + stride1 - 1) / stride1;
  IRR_REQUIRE(oh > 0 && ow > 0, fn, "empty output");
  if (out_c) *out_c = d * d;
  if (out_h) *out_h = oh;
  if (out_w) *out_w = ow;
  return 0;
}

int irr_correlation_generic_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                                int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                irr_stream_t stream) {
  const char* fn = "irr_correlation_generic_fwd";
  IRR_REQUIRE(in1 && in2 && out, fn, "null pointer");
  IRR_REQUIRE(B > 0 && C > 0, fn, "non-positive size");
  int oc, oh, ow;
  int rc = irr_correlation_generic_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2, &oc, &oh,
                                             &ow);
  if (rc) return rc;
  long long total = (long long)B * oc * oh * ow;
  int krad = (kernel_size - 1) / 2;
  corr_generic_kernel<<<(unsigned)((total + 127) / 128), 128, 0, as_stream(stream)>>>(
      in1, in2, out, C, H, W, pad_size, krad, max_displacement, stride1, stride2, max_displacement / stride2, oc, oh, ow,
      total);
  return check_launch(fn);
}

}  // extern "C"
#endif  // IRR_CORR_VARIANT_ONLY
