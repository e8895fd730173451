// First-layer kernels: 3x3 convolution whose INPUT has 1..4 channels (the image: unet.py:49-52 with in_channels 1 / 3).
// K = 9*Cin <= 36 is far below one tensor-core K step worth feeding through TMA (2..8 bytes per pixel), and the layer
// is bound by writing (forward) or reading (backward-weights) the 64-channel full-resolution tensor: 128 B / pixel
// against 576*Cin FMAs / pixel.  So: CUDA cores, fp32 weights in shared memory, and a thread mapping in which the
// 4 lanes that share a pixel group cover its 64 output channels -> every global access is a full 32-byte sector and
// a warp touches whole 128-byte lines.
//   forward : thread = 4 consecutive pixels x 16 output channels (64 accumulators, 16 FMAs per LDS.128 of weights)
//   wgrad   : thread = CPT output channels x all 9*Cin taps, grid-strided over pixels; warp-shuffle + shared-memory
//             block reduction, per-block partials, fixed-order final reduction (reproducible); bias gradient for free
#include <cstdlib>

#include "conv_impl.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kScThreads = 256;

// Forward.  ncu on the first version (profiles/r02_hbm_metrics.txt): 313 M warp instructions per launch and an issue rate of
// 0.6 — every one of the 4 lanes of a pixel group fetched the 18 input values itself (64-bit index arithmetic + bounds
// tests per value), bias and ReLU were separate fp32 passes over the 64 accumulators, and each FFMA2 needed its input
// duplicated into a register pair.  A second version staged the window per BLOCK behind a barrier and was no faster: two
// blocks per SM, both regularly parked in the load -> store -> barrier sequence (ncu: issue 0.42, stalls long-scoreboard /
// barrier / wait).  Now every WARP owns 32 consecutive output pixels of one row: its 3 x 34 input window lives in a
// warp-private, double-buffered shared-memory slab as (x, x) pairs (one LDS.128 = two FFMA2 operands); the window of the
// warp's NEXT unit is loaded into registers before the FMAs of the current one and parked afterwards (__syncwarp only, no
// block barrier), the accumulators start from the bias, and ReLU is the .relu of the bf16x2 conversion.
constexpr int kScPitch = 36;  // float2 per staged row: 34 used, rows stay 16-byte aligned

template <int CIN>
__global__ void __launch_bounds__(kScThreads, 2)
smallc_fwd_kernel(DView src, DView dst, const float* __restrict__ w, const float* __restrict__ bias, int relu, int pad) {
  __shared__ __align__(16) float ws[9 * CIN * 64];
  __shared__ __align__(16) float bs[64];
  extern __shared__ __align__(16) uint8_t sc_dyn[];
  float2 (*xs)[2][3 * CIN][kScPitch] = reinterpret_cast<float2 (*)[2][3 * CIN][kScPitch]>(sc_dyn);  // [warp][buffer][row][column]
  const int cout = dst.c;
  const int o_base = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 9 * CIN * 64; i += kScThreads) {
    const int o = i & 63, tc = i >> 6;  // tc = tap * CIN + c
    const int tap = tc / CIN, c = tc - tap * CIN;
    ws[i] = (o_base + o < cout) ? w[((long long)(o_base + o) * CIN + c) * 9 + tap] : 0.f;
  }
  if (threadIdx.x < 64) bs[threadIdx.x] = (bias && o_base + threadIdx.x < cout) ? bias[o_base + threadIdx.x] : 0.f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = lane & 3, g = lane >> 2;
  const bool lane_on = o_base + q4 * 8 < cout;
  const bool second_half = o_base + 32 + q4 * 8 < cout;  // cout is a multiple of 16, not necessarily of 64
  const unsigned segs = (unsigned)(dst.w + 31) / 32;
  const unsigned total = (unsigned)dst.n * dst.h * segs;
  const unsigned step = gridDim.x * (kScThreads / 32);

  // the window of a unit: rows oy - pad .. +2, columns seg * 32 - pad .. +33; lane l fetches column l, lanes 0 / 1 also 32 / 33
  float v[3 * CIN][2];
  auto fetch = [&](unsigned unit) {
    const unsigned seg = unit % segs, t = unit / segs;
    const int oy = (int)(t % dst.h), n = (int)(t / dst.h);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy + r - pad;
      const bool yok = iy >= 0 && iy < src.h;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ix = (int)seg * 32 + lane + 32 * e - pad;
        const bool ok = yok && ix >= 0 && ix < src.w && (e == 0 || lane < 2);
        const long long so = ok ? src.off(n, iy, ix) : 0;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          float x = 0.f;
          if (ok) {
            x = bf2f(src.p[so + c]);
            if (src.lo) x += bf2f(src.lo[so + c]);  // split tier: the image is carried as hi + lo
          }
          v[r * CIN + c][e] = x;
        }
      }
    }
  };
  auto park = [&](int buf) {
#pragma unroll
    for (int rc = 0; rc < 3 * CIN; ++rc) {
      xs[warp][buf][rc][lane] = make_float2(v[rc][0], v[rc][0]);
      if (lane < 2) xs[warp][buf][rc][32 + lane] = make_float2(v[rc][1], v[rc][1]);
    }
  };

  unsigned unit = blockIdx.x * (kScThreads / 32) + warp;
  int buf = 0;
  if (unit < total) {
    fetch(unit);
    park(0);
  }
  __syncwarp();
  for (; unit < total; unit += step, buf ^= 1) {
    const bool more = unit + step < total;
    if (more) fetch(unit + step);  // in flight during the FMAs below
    const unsigned seg = unit % segs, t = unit / segs;
    const int oy = (int)(t % dst.h), n = (int)(t / dst.h);
    const int ox0 = (int)seg * 32 + g * 4;
    if (lane_on && ox0 < dst.w) {
      float2 acc[4][8];  // 4 pixels x 16 channels as fp32 pairs: FFMA2 does two channels per instruction
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 b2 = *reinterpret_cast<const float2*>(bs + (k >> 2) * 32 + q4 * 8 + (k & 3) * 2);
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p][k] = b2;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          float2 xin[6];
          const float4* xp = reinterpret_cast<const float4*>(&xs[warp][buf][r * CIN + c][g * 4]);
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float4 tt = xp[j];
            xin[2 * j] = make_float2(tt.x, tt.y);
            xin[2 * j + 1] = make_float2(tt.z, tt.w);
          }
#pragma unroll
          for (int sx = 0; sx < 3; ++sx) {
            // lane q4 owns channels [8 q4, 8 q4 + 8) and [32 + 8 q4, 32 + 8 q4 + 8): the four lanes of a pixel then write
            // two contiguous 64-byte runs (full 32-byte sectors per store instruction)
            const float* wp = ws + ((r * 3 + sx) * CIN + c) * 64 + q4 * 8;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 wv = *reinterpret_cast<const float4*>(wp + (j4 >> 1) * 32 + (j4 & 1) * 4);
              const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
#pragma unroll
              for (int p = 0; p < 4; ++p) {
                acc[p][j4 * 2 + 0] = ffma2(xin[p + sx], w01, acc[p][j4 * 2 + 0]);
                acc[p][j4 * 2 + 1] = ffma2(xin[p + sx], w23, acc[p][j4 * 2 + 1]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int ox = ox0 + p;
        if (ox >= dst.w) break;
        const long long oo = dst.off(n, oy, ox) + o_base + q4 * 8;
        if (dst.lo == nullptr) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !second_half) break;
            uint4 cc;
            if (relu) {
              cc.x = pack_bf16x2_relu(acc[p][4 * h + 0].x, acc[p][4 * h + 0].y);
              cc.y = pack_bf16x2_relu(acc[p][4 * h + 1].x, acc[p][4 * h + 1].y);
              cc.z = pack_bf16x2_relu(acc[p][4 * h + 2].x, acc[p][4 * h + 2].y);
              cc.w = pack_bf16x2_relu(acc[p][4 * h + 3].x, acc[p][4 * h + 3].y);
            } else {
              cc.x = pack_bf16x2(acc[p][4 * h + 0].x, acc[p][4 * h + 0].y);
              cc.y = pack_bf16x2(acc[p][4 * h + 1].x, acc[p][4 * h + 1].y);
              cc.z = pack_bf16x2(acc[p][4 * h + 2].x, acc[p][4 * h + 2].y);
              cc.w = pack_bf16x2(acc[p][4 * h + 3].x, acc[p][4 * h + 3].y);
            }
            *reinterpret_cast<uint4*>(dst.p + oo + 32 * h) = cc;
          }
        } else {
          float f0[8], f1[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            f0[2 * k] = acc[p][k].x;
            f0[2 * k + 1] = acc[p][k].y;
            f1[2 * k] = acc[p][4 + k].x;
            f1[2 * k + 1] = acc[p][4 + k].y;
          }
          if (relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              f0[k] = fmaxf(f0[k], 0.f);
              f1[k] = fmaxf(f1[k], 0.f);
            }
          }
          store8s(dst.p, dst.lo, oo, f0);
          if (second_half) store8s(dst.p, dst.lo, oo + 32, f1);
        }
      }
    }
    if (more) park(buf ^ 1);  // the other slab was last read one iteration ago, before the __syncwarp below
    __syncwarp();
  }
}

// ------------------------------------------------------------------ forward on the tensor cores (bf16 tier)
// The CUDA-core kernel above stays issue-/latency-bound at ~22 TFLOP/s (ncu: FMA pipe 42 % busy, issue 42 %: 576 * CIN FMAs per
// pixel are simply too many instructions for a layer that should cost one write of the 64-channel tensor).  Here the
// 9 * CIN patch values of a pixel are the K dimension of a warp-level mma.sync.m16n8k16 (bf16 x bf16 -> fp32): K = 16 per
// step holds the whole 3x3 patch of one input channel pair, the weight fragments of all 64 output channels stay in
// registers for the whole kernel, and a warp turns 32 pixels x 64 channels into 16 * ceil(9 CIN / 16) tensor instructions
// instead of 9216 * CIN / 32 FMAs per lane.  (tcgen05 would need the patches in shared memory in UMMA layout and a TMEM
// round trip for a GEMM with K = 16: legacy mma.sync from registers is the right tool for this one layer.)
//   A (16 pixels x 16 k): built from the warp's staged bf16 window with two LDS.U16 per register — lane (g, t) needs
//     pixels g / g + 8 and k = 2t, 2t+1, 2t+8, 2t+9, whose window offsets are per-lane constants;
//   B (16 k x 8 channels) = w[o][k] (k = c * 9 + tap is contiguous in the OIHW tensor), bf16-rounded like every other
//     layer's weights;
//   C starts from the bias; ReLU is the .relu of the bf16x2 conversion; the 16 x 64 result goes through a padded
//     warp-private shared-memory tile so that 8 lanes store one pixel's 128 contiguous bytes.
constexpr int kMmaPitch = 40;                 // bf16 per staged window row: 34 used
constexpr int kMmaStageRow = 72;              // 32-bit words per staged output row: 64 used (+8: conflict-free column writes)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CIN>
__global__ void __launch_bounds__(kScThreads, 2)
smallc_mma_fwd_kernel(DView src, DView dst, const float* __restrict__ w, const float* __restrict__ bias, int relu, int pad) {
  constexpr int KR = 9 * CIN;            // real K
  constexpr int KS = (KR + 15) / 16;     // K steps
  constexpr int WROWS = 3 * CIN + 1;     // staged window rows (+1: a row of zeros for the padding k)
  __shared__ __align__(16) unsigned short xs[kScThreads / 32][2][WROWS][kMmaPitch];
  __shared__ __align__(16) uint32_t stg[kScThreads / 32][16][kMmaStageRow / 2];  // bf16x2 words: 16 pixels x 64 channels
  const int cout = dst.c;
  const int o_base = blockIdx.y * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  // weight fragments: b[ks][j][0] = {W[k0][o], W[k0+1][o]}, b[ks][j][1] = {W[k0+8][o], W[k0+9][o]}, k0 = 16 ks + 2t, o = o_base + 8j + g
  uint32_t bw[KS][8][2];
  float bz[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = o_base + j * 8 + g;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k0 = ks * 16 + 2 * t + 8 * h;
        const float w0 = (o < cout && k0 < KR) ? w[(long long)o * KR + k0] : 0.f;
        const float w1 = (o < cout && k0 + 1 < KR) ? w[(long long)o * KR + k0 + 1] : 0.f;
        bw[ks][j][h] = pack_bf16x2(w0, w1);
      }
    const int oc = o_base + j * 8 + 2 * t;  // the accumulator columns of this lane
    bz[j][0] = (bias && oc < cout) ? bias[oc] : 0.f;
    bz[j][1] = (bias && oc + 1 < cout) ? bias[oc + 1] : 0.f;
  }
  // window offsets of this lane's k values: k = c * 9 + r * 3 + s lives at row r * CIN + c, column + s; padding k -> zero row
  int koff[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = ks * 16 + 2 * t + (q & 1) + 8 * (q >> 1);
      const int c = k / 9, tap = k - c * 9, r = tap / 3, sx = tap - r * 3;
      koff[ks][q] = k < KR ? (r * CIN + c) * kMmaPitch + sx : (3 * CIN) * kMmaPitch;
    }
  for (int i = lane; i < 2 * kMmaPitch; i += 32) xs[warp][i / kMmaPitch][3 * CIN][i % kMmaPitch] = 0;  // the zero rows

  const unsigned segs = (unsigned)(dst.w + 31) / 32;
  const unsigned total = (unsigned)dst.n * dst.h * segs;
  const unsigned step = gridDim.x * (kScThreads / 32);
  unsigned short v[3 * CIN][2];
  auto fetch = [&](unsigned unit) {
    const unsigned seg = unit % segs, tt = unit / segs;
    const int oy = (int)(tt % dst.h), n = (int)(tt / dst.h);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = oy + r - pad;
      const bool yok = iy >= 0 && iy < src.h;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ix = (int)seg * 32 + lane + 32 * e - pad;
        const bool ok = yok && ix >= 0 && ix < src.w && (e == 0 || lane < 2);
        const long long so = ok ? src.off(n, iy, ix) : 0;
#pragma unroll
        for (int c = 0; c < CIN; ++c)
          v[r * CIN + c][e] = ok ? reinterpret_cast<const unsigned short*>(src.p)[so + c] : (unsigned short)0;
      }
    }
  };
  auto park = [&](int buf) {
#pragma unroll
    for (int rc = 0; rc < 3 * CIN; ++rc) {
      xs[warp][buf][rc][lane] = v[rc][0];
      if (lane < 2) xs[warp][buf][rc][32 + lane] = v[rc][1];
    }
  };

  unsigned unit = blockIdx.x * (kScThreads / 32) + warp;
  int buf = 0;
  if (unit < total) {
    fetch(unit);
    park(0);
  }
  __syncwarp();
  for (; unit < total; unit += step, buf ^= 1) {
    const bool more = unit + step < total;
    if (more) fetch(unit + step);  // in flight during the tensor instructions below
    const unsigned seg = unit % segs, tt = unit / segs;
    const int oy = (int)(tt % dst.h), n = (int)(tt / dst.h);
    const unsigned short* win = &xs[warp][buf][0][0];
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {  // two tiles of 16 pixels
      const int px0 = (int)seg * 32 + mt * 16;
      if (px0 >= dst.w) break;
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j][0] = bz[j][0];
        acc[j][1] = bz[j][1];
        acc[j][2] = bz[j][0];
        acc[j][3] = bz[j][1];
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a[4];
        const unsigned short* wp = win + mt * 16 + g;
        a[0] = (uint32_t)wp[koff[ks][0]] | ((uint32_t)wp[koff[ks][1]] << 16);          // pixel g,     k = 2t, 2t + 1
        a[1] = (uint32_t)wp[koff[ks][0] + 8] | ((uint32_t)wp[koff[ks][1] + 8] << 16);  // pixel g + 8
        a[2] = (uint32_t)wp[koff[ks][2]] | ((uint32_t)wp[koff[ks][3]] << 16);          // pixel g,     k = 2t + 8, 2t + 9
        a[3] = (uint32_t)wp[koff[ks][2] + 8] | ((uint32_t)wp[koff[ks][3] + 8] << 16);
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_bf16_16816(acc[j], a, bw[ks][j]);
      }
      // lane (g, t) holds pixels g / g + 8, channels 8j + 2t, 8j + 2t + 1: park as bf16x2 words, row pitch 36 words
      // (word index 36 row + 4j + t: the 32 lanes of one store hit 32 different banks)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        stg[warp][g][j * 4 + t] = relu ? pack_bf16x2_relu(acc[j][0], acc[j][1]) : pack_bf16x2(acc[j][0], acc[j][1]);
        stg[warp][g + 8][j * 4 + t] = relu ? pack_bf16x2_relu(acc[j][2], acc[j][3]) : pack_bf16x2(acc[j][2], acc[j][3]);
      }
      __syncwarp();
      // 8 lanes move one pixel's 128 bytes: lane -> (pixel lane / 8 + 4 i, 16-byte chunk lane % 8)
      const int ch = lane & 7;
      if (o_base + ch * 8 < cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pr = (lane >> 3) + 4 * i;
          const int ox = px0 + pr;
          if (ox < dst.w)
            *reinterpret_cast<uint4*>(dst.p + dst.off(n, oy, ox) + o_base + ch * 8) =
                *reinterpret_cast<const uint4*>(&stg[warp][pr][ch * 4]);
        }
      }
      __syncwarp();
    }
    if (more) park(buf ^ 1);  // the other window was last read one iteration ago
    __syncwarp();
  }
}

// partial layout per block: [64 couts][9*CIN] weights followed by [64] bias sums.
// Thread = CPT output channels x all 9*CIN taps; one loop iteration = a strip of 4 consecutive pixels of one row, whose
// 3 x 6 input window is loaded once per input channel (sliding window) and whose index arithmetic is paid once.
template <int CIN, int CPT>
__global__ void __launch_bounds__(kScThreads)
smallc_wgrad_kernel(DView dz, DView src, int pad, float* __restrict__ partial) {
  constexpr int LPP = 64 / CPT;       // lanes per pixel strip
  constexpr int NT = 9 * CIN;         // taps x input channels
  constexpr int ROW = 64 * NT + 64;   // floats per block partial
  extern __shared__ float red[];      // [8 warps][ROW]
  const int cout = dz.c;
  const int o_base = blockIdx.y * 64;
  const int cg = threadIdx.x % LPP;
  const int slot = threadIdx.x / LPP;
  constexpr int slots = kScThreads / LPP;
  const int o0 = o_base + cg * CPT;
  float acc[CPT][NT];
  float accb[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  }
  const int groups_per_row = (dz.w + 3) >> 2;
  const unsigned total_groups = (unsigned)dz.n * dz.h * groups_per_row;
  if (o0 < cout) {
    for (unsigned g = blockIdx.x * slots + slot; g < total_groups; g += gridDim.x * slots) {
      const unsigned xg = g % groups_per_row;
      const unsigned t2 = g / groups_per_row;
      const int oy = (int)(t2 % dz.h), n = (int)(t2 / dz.h);
      const int ox0 = (int)xg * 4;
      float z[4][CPT];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (ox0 + p < dz.w) {
          const bf16* zp = dz.p + dz.off(n, oy, ox0 + p) + o0;
          if (CPT == 16) {
            float t[8];
            unpack8(*reinterpret_cast<const bf16x8*>(zp), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[p][i] = t[i];
            unpack8(*reinterpret_cast<const bf16x8*>(zp + 8), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[p][(8 + i) % CPT] = t[i];
          } else if (CPT == 8) {
            float t[8];
            unpack8(*reinterpret_cast<const bf16x8*>(zp), t);
#pragma unroll
            for (int i = 0; i < CPT; ++i) z[p][i] = t[i % 8];
          } else {
            const uint2 u = *reinterpret_cast<const uint2*>(zp);
            const float2 a2 = bf2x_to_f2(u.x), b2 = bf2x_to_f2(u.y);
            z[p][0] = a2.x;
            z[p][1 % CPT] = a2.y;
            z[p][2 % CPT] = b2.x;
            z[p][3 % CPT] = b2.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < CPT; ++i) z[p][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) accb[i] += z[p][i];
      }
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        float xin[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int iy = oy + r - pad;
          const bool yok = iy >= 0 && iy < src.h;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int ix = ox0 + j - pad;
            xin[r][j] = (yok && ix >= 0 && ix < src.w) ? bf2f(src.p[src.off(n, iy, ix) + c]) : 0.f;
          }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int sx = 0; sx < 3; ++sx)
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
              for (int i = 0; i < CPT; ++i) acc[i][c * 9 + r * 3 + sx] += z[p][i] * xin[r][p + sx];
      }
    }
  }
  // reduce over the pixel slots: first inside the warp (lanes with equal cg), then across the 8 warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int off = LPP; off < 32; off <<= 1) accb[i] += __shfl_xor_sync(0xffffffffu, accb[i], off);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int off = LPP; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[i][j] = v;
    }
  }
  if (lane < LPP) {
    float* rw = red + warp * ROW;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      rw[64 * NT + lane * CPT + i] = accb[i];
#pragma unroll
      for (int j = 0; j < NT; ++j) rw[(lane * CPT + i) * NT + j] = acc[i][j];
    }
  }
  __syncthreads();
  float* out = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * ROW;
  for (int q = threadIdx.x; q < ROW; q += kScThreads) {
    float s = 0.f;
#pragma unroll
    for (int wdx = 0; wdx < kScThreads / 32; ++wdx) s += red[wdx * ROW + q];
    out[q] = s;
  }
}

// dw[o][c][tap] = sum_b partial[yblk][b][(o%64)*NT + c*9 + tap];  db[o] likewise
__global__ void smallc_wgrad_reduce_kernel(const float* __restrict__ partial, int blocks, int nt, int cout,
                                           float* __restrict__ dw, float* __restrict__ db) {
  const int row = 64 * nt + 64;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cout * nt + cout) return;
  int yb, q;
  if (e < cout * nt) {
    const int o = e / nt;
    yb = o >> 6;
    q = (o & 63) * nt + (e - o * nt);
  } else {
    const int o = e - cout * nt;
    yb = o >> 6;
    q = 64 * nt + (o & 63);
  }
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += (double)partial[((long long)yb * blocks + b) * row + q];
  if (e < cout * nt)
    dw[e] = (float)s;
  else if (db)
    db[e - cout * nt] = (float)s;
}

constexpr int kScWgradBlocks = 2 * kNumSMsB200;

static bool sc_aligned_out(const b200_view& v) {
  return v.c % 16 == 0 && reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0 && reinterpret_cast<uintptr_t>(v.lo) % 16 == 0 &&
         v.stride_w % 8 == 0 && v.stride_h % 8 == 0 && (v.n == 1 || v.stride_n % 8 == 0);
}

bool smallc_conv_fwd_ok(const b200_conv_fwd_params* p) {
  return p->num_src == 1 && p->taps == 9 && p->src[0].c >= 1 && p->src[0].c <= 4 && sc_aligned_out(p->dst) && p->w_f32;
}

template <int CIN>
static int sc_fwd_launch(const b200_conv_fwd_params* p, cudaStream_t st) {
  const long long units = (long long)p->dst.n * p->dst.h * ((p->dst.w + 31) / 32);  // one warp iteration each
  if (units >= (1LL << 32)) return fail(-1, "smallc_fwd: more than 2^32 row segments");
  long long gx = (units + kScThreads / 32 - 1) / (kScThreads / 32);
  if (gx > 2 * kNumSMsB200) gx = 2 * kNumSMsB200;  // persistent: 2 blocks per SM
  dim3 grid((unsigned)gx, (unsigned)((p->dst.c + 63) / 64));
  const size_t smem = (size_t)(kScThreads / 32) * 2 * 3 * CIN * kScPitch * sizeof(float2);
  auto kern = smallc_fwd_kernel<CIN>;
  static bool attr_done = false;
  if (!attr_done && smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "smallc_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  kern<<<grid, kScThreads, smem, st>>>(dview(p->src[0]), dview(p->dst), p->w_f32, p->bias, p->relu, p->pad);
  return check_launch("smallc_fwd");
}

template <int CIN>
static int sc_mma_fwd_launch(const b200_conv_fwd_params* p, cudaStream_t st) {
  const long long units = (long long)p->dst.n * p->dst.h * ((p->dst.w + 31) / 32);
  if (units >= (1LL << 32)) return fail(-1, "smallc_fwd: more than 2^32 row segments");
  long long gx = (units + kScThreads / 32 - 1) / (kScThreads / 32);
  if (gx > 2 * kNumSMsB200) gx = 2 * kNumSMsB200;
  dim3 grid((unsigned)gx, (unsigned)((p->dst.c + 63) / 64));
  smallc_mma_fwd_kernel<CIN><<<grid, kScThreads, 0, st>>>(dview(p->src[0]), dview(p->dst), p->w_f32, p->bias, p->relu, p->pad);
  return check_launch("smallc_fwd (mma)");
}

static int g_sc_mma = getenv("B200UNET_NO_FIRST_LAYER_MMA") ? 0 : 1;

int smallc_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st) {
  if (g_sc_mma && !p->src[0].lo && !p->dst.lo) {  // bf16 tier: tensor-core kernel; the split tier keeps fp32 weights / CUDA cores
    switch (p->src[0].c) {
      case 1: return sc_mma_fwd_launch<1>(p, st);
      case 2: return sc_mma_fwd_launch<2>(p, st);
      case 3: return sc_mma_fwd_launch<3>(p, st);
      default: return sc_mma_fwd_launch<4>(p, st);
    }
  }
  switch (p->src[0].c) {
    case 1: return sc_fwd_launch<1>(p, st);
    case 2: return sc_fwd_launch<2>(p, st);
    case 3: return sc_fwd_launch<3>(p, st);
    default: return sc_fwd_launch<4>(p, st);
  }
}

bool smallc_conv_wgrad_ok(const b200_conv_wgrad_params* p) {
  return p->num_src == 1 && p->taps == 9 && p->src[0].c >= 1 && p->src[0].c <= 4 && sc_aligned_out(p->dz);
}

size_t smallc_conv_wgrad_workspace(const b200_conv_wgrad_params* p) {
  const int nt = 9 * p->src[0].c;
  return (size_t)((p->dz.c + 63) / 64) * kScWgradBlocks * (64 * nt + 64) * sizeof(float);
}

template <int CIN, int CPT>
static int sc_wgrad_launch(const b200_conv_wgrad_params* p, float* ws, int blocks, cudaStream_t st) {
  constexpr int ROW = 64 * 9 * CIN + 64;
  const size_t smem = (size_t)(kScThreads / 32) * ROW * sizeof(float);
  auto kern = smallc_wgrad_kernel<CIN, CPT>;
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "smallc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid(blocks, (unsigned)((p->dz.c + 63) / 64));
  kern<<<grid, kScThreads, smem, st>>>(dview(p->dz), dview(p->src[0]), p->pad, ws);
  return check_launch("smallc_wgrad");
}

int smallc_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t need = smallc_conv_wgrad_workspace(p);
  if (!ws || ws_bytes < need) return fail(-1, "conv_wgrad (first layer): workspace too small (%zu < %zu)", ws_bytes, need);
  const long long npix = view_pixels(p->dz);
  int blocks = (int)((npix + 255) / 256);
  if (blocks > kScWgradBlocks) blocks = kScWgradBlocks;
  if (blocks < 1) blocks = 1;
  int r;
  switch (p->src[0].c) {
    case 1: r = sc_wgrad_launch<1, 16>(p, (float*)ws, blocks, st); break;
    case 2: r = sc_wgrad_launch<2, 8>(p, (float*)ws, blocks, st); break;
    case 3: r = sc_wgrad_launch<3, 4>(p, (float*)ws, blocks, st); break;
    default: r = sc_wgrad_launch<4, 4>(p, (float*)ws, blocks, st); break;
  }
  if (r) return r;
  const int nt = 9 * p->src[0].c;
  const int outs = p->dz.c * nt + p->dz.c;
  smallc_wgrad_reduce_kernel<<<(outs + 255) / 256, 256, 0, st>>>((const float*)ws, blocks, nt, p->dz.c, p->dw_f32, p->db_f32);
  return check_launch("smallc_wgrad_reduce");
}

}  // namespace b200
